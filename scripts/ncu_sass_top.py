#!/usr/bin/env python
"""Top SASS instructions of an .ncu-rep by stall samples inside an address window, with their dominant stall reasons.
usage: scripts/ncu_sass_top.py rep [lo_hex hi_hex] [top_n]   (addresses relative to the kernel start)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {}
for i, k in enumerate(h): ix.setdefault(k, i)
data = [r for r in rows[2:] if len(r) == len(h)]
base = int(data[0][0], 16)
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
sel = [(int(r[0], 16) - base, r) for r in data if lo <= int(r[0], 16) - base < hi]
ts = sum(float(r[ix["# Samples"]] or 0) for _, r in sel); ti = sum(float(r[ix["Instructions Executed"]] or 0) for _, r in sel)
allS = sum(float(r[ix["# Samples"]] or 0) for r in data); allI = sum(float(r[ix["Instructions Executed"]] or 0) for r in data)
print(f"window [{lo:#x}, {hi:#x}): {ts:.0f} samples ({ts/allS*100:.1f}% of kernel), {ti:.4g} warp instructions ({ti/allI*100:.1f}%)")
agg = {k: sum(float(r[ix[k]] or 0) for _, r in sel) for k in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {v/ts*100:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
sel.sort(key=lambda ar: -float(ar[1][ix["# Samples"]] or 0))
for a, r in sel[:top]:
    s = float(r[ix["# Samples"]] or 0)
    why = sorted(((float(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{a:#7x} {s/ts*100:5.1f}%  inst {float(r[ix['Instructions Executed']] or 0)/ti*100:4.1f}%  {why[0][1]:>16s} {why[0][0]/max(s,1)*100:3.0f}%  {why[1][1]:>14s} {why[1][0]/max(s,1)*100:3.0f}% | {r[1].strip()[:90]}")
