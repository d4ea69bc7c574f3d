#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics, stall mix, opcode mix.
usage: scripts/ncu_summary.py gpurun_out/prof_compress.ncu-rep > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__shared_mem_config_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
print(f"# ncu summary of {rep}")
for k in KEYS:
    if k in m:
        print(f"{k:75s} {m[k][0]} {m[k][1]}")
print("\n# warp stall reasons (avg warps stalled per issue-active cycle)")
st = [(h, float(v[0])) for h, v in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
for h, v in sorted(st, key=lambda x: -x[1])[:8]:
    print(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:25s} {v:.3f}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h2 = rows[1]
ix = {h: i for i, h in enumerate(h2)}
data = [r for r in rows[2:] if len(r) == len(h2)]
tot = sum(float(r[ix["Instructions Executed"]] or 0) for r in data)
samp = sum(float(r[ix["# Samples"]] or 0) for r in data)
byop, sop = collections.Counter(), collections.Counter()
for r in data:
    parts = r[ix["Source"]].split()
    op = (parts[1] if parts and parts[0].startswith("@") else parts[0] if parts else "").split(".")[0]
    byop[op] += float(r[ix["Instructions Executed"]] or 0)
    sop[op] += float(r[ix["# Samples"]] or 0)
print(f"\n# opcode mix: {len(data)} SASS instructions, {tot:.4g} warp instructions executed, {samp:.0f} stall samples")
for op, c in byop.most_common(16):
    print(f"{op:12s} {c / tot * 100:5.1f}% of instructions  {sop[op] / max(samp, 1) * 100:5.1f}% of stall samples")
