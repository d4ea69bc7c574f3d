#!/usr/bin/env python
"""One compress launch timed three ways: CUDA events around the call, %globaltimer stamps taken INSIDE the kernel
(SNB_DEBUG_SPAN=1), and -- when this script runs under `ncu --metrics gpu__time_duration.sum` -- the profiler's number.
Used to bisect why the v5 streaming kernel ran at 46 ms under ncu and 23 ms in the bench (round-1 verdict, weak #3).
Usage: SNB_DEBUG_SPAN=1 python scripts/gpu_span_probe.py [MiB] [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SNB_DEBUG_SPAN", "1")
import torch
import __graft_entry__ as g
g.build()
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import text_corpus
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = mib << 20
lib = _lib.default(); dc = DeviceCodec(lib)
d_in = torch.from_numpy(text_corpus(n)).cuda()
out = torch.empty(int(lib.snb_max_compressed_len_framed(n)), dtype=torch.uint8, device="cuda")
tag = f"L2_TABLES={os.environ.get('SNB_TUNE_L2_TABLES', 'default')} algo={os.environ.get('SNB_ALGO', '0')}"
if os.environ.get("SNB_ALGO"):
    lib.snb_set_compress_algo(int(os.environ["SNB_ALGO"]))
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dc.compress_framed(d_in, out=out, sync=False); e1.record(); torch.cuda.synchronize()
    print(f"[{tag}] launch {i}: events {e0.elapsed_time(e1):.2f} ms, in-kernel globaltimer span {lib.snb_debug_kernel_span_ns() / 1e6:.2f} ms", flush=True)
