#!/usr/bin/env python
"""Why does compress_stream_kernel take 46 ms under ncu and 23 ms in the bench loop?  Hypothesis: what the
profiler changes is the idle gap in front of every launch.  Times the same device-resident launch (CUDA events
around one call) back to back, after host sleeps of several lengths, and after a sleep followed by a short
DRAM-heavy / ALU-heavy "wake-up" kernel; samples SM / memory clocks around each.
Usage: python scripts/gpu_idle_probe.py [MiB]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
g.build()
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import text_corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
data = text_corpus(n)
lib = _lib.default(); dc = DeviceCodec(lib)
d_in = torch.from_numpy(data).cuda()
out = torch.empty(int(lib.snb_max_compressed_len_framed(n)), dtype=torch.uint8, device="cuda")
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
big2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")

def clocks():
    try:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,pstate", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True, timeout=10).stdout.strip()
        return r
    except Exception as e:
        return str(e)

def one():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dc.compress_framed(d_in, out=out, sync=False); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

for _ in range(3): one()
print("clocks idle-ish:", clocks())
print("back to back      :", " ".join(f"{one():.2f}" for _ in range(6)), flush=True)
for gap in (0.002, 0.01, 0.05, 0.2, 1.0, 3.0):
    r = []
    for _ in range(4):
        time.sleep(gap); r.append(one())
    print(f"sleep {gap:5.3f} s before:", " ".join(f"{x:.2f}" for x in r), "| clocks now:", clocks(), flush=True)
for name, fn in (("dram copy 1 GiB", lambda: big2.copy_(big)), ("alu (u8 mul-add x32 on 64 MiB)", lambda: [big[:1 << 26].mul_(3).add_(1) for _ in range(32)])):
    r = []
    for _ in range(4):
        time.sleep(1.0); fn(); r.append(one())
    print(f"sleep 1 s + {name}:", " ".join(f"{x:.2f}" for x in r), flush=True)
print("back to back again:", " ".join(f"{one():.2f}" for _ in range(6)), flush=True)
# two launches per idle gap: is only the first one slow?
r = []
for _ in range(3):
    time.sleep(1.0); r.append((one(), one(), one()))
print("sleep 1 s then 3 launches:", " | ".join(" ".join(f"{x:.2f}" for x in t) for t in r), flush=True)
