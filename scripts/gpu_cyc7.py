#!/usr/bin/env python
"""Diagnostic (SNB_LIB=.../libsnb200_cyc*.so, -DSNB_CYC): critical-path cycles of compress_block_v7 per segment."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import text_corpus
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
geoms = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,6").split(",")]
n = mib << 20
data = text_corpus(n)
lib = _lib.default()
lib.dll.snb_debug_cyc.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
dc = DeviceCodec(lib)
d_in = torch.from_numpy(data).cuda()
out = torch.empty(int(lib.snb_max_compressed_len_framed(n)), dtype=torch.uint8, device="cuda")
dc.set_profiling(True)
lib.snb_set_compress_algo(7)
names = ["pre-resolve", "wait for the cursor", "fast: cursor -> handed over", "fast: emit + commit", "slow turn (all of it)"]
for g in geoms:
    lib.snb_set_compress_warps(g)
    dc.compress_framed(d_in, out=out)
    v = (C.c_ulonglong * 16)()
    lib.dll.snb_debug_cyc(v, 1)
    dc.compress_framed(d_in, out=out)
    ms = dc.profile()["compress_ms"]
    lib.dll.snb_debug_cyc(v, 1)
    turns = max(1, v[12])
    chunks = n // 65536
    print(f"== {g} pairs: {ms:.2f} ms -> {n / ms / 1e6:.1f} GB/s; turns/chunk {turns / chunks:.0f}, slow turns {v[10] / chunks:.0f}, serial chain {v[13] / chunks:.0f}, serial lit {v[14] / chunks:.0f}, long {v[11] / chunks:.0f}")
    for i, nm in enumerate(names):
        print(f"   {nm:20s} {v[i] / turns:8.1f} cyc/turn")
    act = sum(v[2:5])
    print(f"   active part of a turn {act / turns:8.1f};  both warps' block time {v[15] / chunks / 2e3:.0f} kcyc/chunk = {v[15] / 2 / turns:.0f} cyc per turn", flush=True)
