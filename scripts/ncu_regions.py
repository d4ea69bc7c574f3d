#!/usr/bin/env python
"""Instruction / stall-sample totals per source-line region of an .ncu-rep.
usage: scripts/ncu_regions.py rep name:lo-hi [name:lo-hi ...]   (lines of snb_device.cuh; helpers are attributed by call line only when inlined code keeps its own line)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
regs = []
for a in sys.argv[2:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regs.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi_]; ix = {}
for i, k in enumerate(h): ix.setdefault(k, i)
data = [r for r in rows[hi_ + 1:] if len(r) == len(h) and r[0].isdigit()]
ti = sum(float(r[ix["Instructions Executed"]] or 0) for r in data); ts = sum(float(r[ix["# Samples"]] or 0) for r in data)
acc = {n: [0, 0] for n, _, _ in regs}; acc["other"] = [0, 0]
for r in data:
    ln = int(r[0]); i = float(r[ix["Instructions Executed"]] or 0); s = float(r[ix["# Samples"]] or 0)
    for n, lo, hi in regs:
        if lo <= ln <= hi: acc[n][0] += i; acc[n][1] += s; break
    else: acc["other"][0] += i; acc["other"][1] += s
for n, (i, s) in acc.items(): print(f"{n:22s} inst {i/ti*100:5.1f}% ({i:.3g})  samples {s/ts*100:5.1f}%")
