#!/usr/bin/env python
"""Per-source-line stall samples / instruction counts of an .ncu-rep (needs -lineinfo + --import-source on).
usage: scripts/ncu_lines.py rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]; ix = {}
for i, k in enumerate(h):
    ix.setdefault(k, i)
ixl = ix["stall_long_sb"]
data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0].isdigit()]
tot_s = sum(float(r[ix["# Samples"]] or 0) for r in data); tot_i = sum(float(r[ix["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.4g}")
data.sort(key=lambda r: -float(r[ix["# Samples"]] or 0))
for r in data[:top]:
    s = float(r[ix["# Samples"]] or 0); i = float(r[ix["Instructions Executed"]] or 0)
    lsb = float(r[ixl] or 0)
    print(f"{r[0]:>5s} {s / tot_s * 100:5.1f}% smp (long_sb {lsb / max(s,1) * 100:3.0f}%) {i / tot_i * 100:5.1f}% inst  thr {r[ix['Avg. Threads Executed']]:>3s} | {r[1][:110]}")
