#!/usr/bin/env python
"""Where the time of snb_dev_uncompress_framed goes (SNB_DEBUG_DEVUNC=1 prints the library's host-side timeline)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import text_corpus
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = _lib.default(); dc = DeviceCodec(lib)
d_in = torch.from_numpy(text_corpus(mib << 20)).cuda()
out, w = dc.compress_framed(d_in)
s = out[:w].clone()
d_back = torch.empty_like(d_in)
dc.set_profiling(True)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st, rd, wr = dc.uncompress_framed(s, d_back)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"call {i}: {dt * 1e3:.3f} ms wall, kernel {dc.profile()['decompress_ms']:.3f} ms, status {st}, written {wr}", flush=True)
print("equal:", bool((d_back == d_in).all()))
