#!/usr/bin/env python
"""Quick A/B on the GPU: compress kernel variants (algo, warps per CTA) on a corpus; every
variant's output is compared with the oracle stream.
Usage: python scripts/gpu_tune.py [MiB] [text|mixed|random|proto] [algo:warps,...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import __graft_entry__ as g

g.build()
from bench import CpuReference
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import mixed_corpus, text_corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
n = mib << 20
if kind == "text":
    data = text_corpus(n)
elif kind == "mixed":
    data = mixed_corpus(n, seed=1)
elif kind == "random":
    import numpy as np
    data = np.random.default_rng(3).integers(0, 256, n, dtype=np.uint8)
else:  # "proto"
    from nim_snappy_b200.workloads import PROTO_FILES, _cat, tile
    data = tile(_cat(PROTO_FILES), n)
ref = CpuReference()
_, stream = ref.framed_stream(data)
lib = _lib.default()
dc = DeviceCodec(lib)
d_in = torch.from_numpy(data).cuda()
d_expect = torch.from_numpy(stream).cuda()
out = torch.empty(int(lib.snb_max_compressed_len_framed(n)), dtype=torch.uint8, device="cuda")
variants = [(3, 0), (2, 7), (1, 7)]
if len(sys.argv) > 3:
    variants = [tuple(int(x) for x in v.split(":")) for v in sys.argv[3].split(",")]
dc.set_profiling(True)
for algo, warps in variants:
    lib.snb_set_compress_algo(algo)
    lib.snb_set_compress_warps(warps)
    _, w = dc.compress_framed(d_in, out=out)
    ok = w == stream.size and bool((out[:w] == d_expect).all())
    ms = []
    for _ in range(3):
        _, w = dc.compress_framed(d_in, out=out)
        ms.append(dc.profile()["compress_ms"])
    best = min(ms)
    print(f"algo={algo} warps={warps} ok={ok} compress_kernel_ms={best:.3f} -> {n / best / 1e6:.2f} GB/s of input", flush=True)
chunks, walk = dc.frame_index(stream)
d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(4):
    st = dc.uncompress_chunks(d_expect, chunks, d_back)
    dms = dc.profile()["decompress_ms"]
print(f"decompress ok={not st.any() and bool((d_back == d_in).all())} kernel_ms={dms:.3f} -> {n / dms / 1e6:.2f} GB/s")
