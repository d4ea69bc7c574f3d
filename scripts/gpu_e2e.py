#!/usr/bin/env python
"""End-to-end (host buffers, pinned) throughput of snb_compress_framed / snb_uncompress_framed next to the raw
PCIe copy rates of the box.  Usage: python scripts/gpu_e2e.py [MiB] [reps]   (env SNB_TUNE_* apply)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as g

g.build()
import nim_snappy_b200 as snappy
from nim_snappy_b200 import _lib
from nim_snappy_b200.codec import maxCompressedLenFramed
from nim_snappy_b200.workloads import text_corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = mib << 20
lib = _lib.default()
api = snappy.Snappy(lib)
host_in = torch.from_numpy(text_corpus(n)).pin_memory()
host_out = torch.empty(maxCompressedLenFramed(n), dtype=torch.uint8).pin_memory()
host_back = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")


def best(fn, k=reps):
    ts = []
    for _ in range(k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts)


t = best(lambda: dev.copy_(host_in, non_blocking=True))
print(f"pcie h2d {n / t / 1e9:.2f} GB/s")
t = best(lambda: host_back.copy_(dev, non_blocking=True))
print(f"pcie d2h {n / t / 1e9:.2f} GB/s")
s2 = torch.cuda.Stream()


def both():
    dev.copy_(host_in, non_blocking=True)
    with torch.cuda.stream(s2):
        host_back.copy_(dev2, non_blocking=True)


dev2 = torch.empty(n, dtype=torch.uint8, device="cuda")
t = best(both)
print(f"pcie h2d+d2h concurrently: {n / t / 1e9:.2f} GB/s each direction")

w = api.compressFramed(host_in.numpy(), host_out.numpy())
w = api.compressFramed(host_in.numpy(), host_out.numpy())
t = best(lambda: api.compressFramed(host_in.numpy(), host_out.numpy()))
print(f"e2e compressFramed {n / t / 1e9:.2f} GB/s ({t * 1e3:.2f} ms), {w} bytes out")
stream = host_out.numpy()[:w]
r = api.uncompressFramed(stream, host_back.numpy())
r = api.uncompressFramed(stream, host_back.numpy())
t = best(lambda: api.uncompressFramed(stream, host_back.numpy()))
assert r == (w, n) and np.array_equal(host_back.numpy(), host_in.numpy())
print(f"e2e uncompressFramed {n / t / 1e9:.2f} GB/s ({t * 1e3:.2f} ms) dec_slab={os.environ.get('SNB_TUNE_DEC_SLAB', 'default')}")
