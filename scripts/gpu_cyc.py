#!/usr/bin/env python
"""Diagnostic (SNB_LIB=.../libsnb200_cyc*.so, built with -DSNB_CYC): critical-path cycles of compress_block_v5's
window pass per segment, for several warp geometries.  usage: python scripts/gpu_cyc.py [MiB] [smem+global,...]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import text_corpus
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
geoms = [tuple(int(x) for x in g.split("+")) for g in (sys.argv[2] if len(sys.argv) > 2 else "4+0,4+8,4+18").split(",")]
n = mib << 20
data = text_corpus(n)
lib = _lib.default()
lib.dll.snb_debug_cyc.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
dc = DeviceCodec(lib)
d_in = torch.from_numpy(data).cuda()
out = torch.empty(int(lib.snb_max_compressed_len_framed(n)), dtype=torch.uint8, device="cuda")
dc.set_profiling(True)
names = ["->T", "->ml", "->J", "->hintTn", "->walk entry", "->walk+LK", "->emit/op", "long match", "->commit/loop end"]
for s, gw in geoms:
    lib.snb_set_compress_warps(s + 100 * (gw + 1))
    dc.compress_framed(d_in, out=out)
    v = (C.c_ulonglong * 16)()
    lib.dll.snb_debug_cyc(v, 1)
    dc.compress_framed(d_in, out=out)
    ms = dc.profile()["compress_ms"]
    lib.dll.snb_debug_cyc(v, 1)
    wins = max(1, v[12])
    chunks = n // 65536
    print(f"== {s} smem + {gw} global warps: {ms:.2f} ms -> {n / ms / 1e6:.1f} GB/s; windows/chunk {wins / chunks:.0f}, serial chain {v[9] / chunks:.0f}, serial lit {v[10] / chunks:.0f}, long {v[11] / chunks:.0f}")
    tot = 0
    for i, nm in enumerate(names):
        print(f"   {nm:20s} {v[i] / wins:8.1f} cyc/window")
        tot += v[i]
    print(f"   window-pass total     {tot / wins:8.1f} cyc/window;  whole block {v[15] / chunks / 1e3:.0f} kcyc/chunk = {v[15] / wins:.0f} cyc per window pass", flush=True)
