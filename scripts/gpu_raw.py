#!/usr/bin/env python
"""Raw-format (snappy.nim:27-110) streams of many blocks through the host-buffer API: block-parallel compress and
uncompress (block starts found on the device).  usage: python scripts/gpu_raw.py [MiB] [text|mixed]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as g

g.build()
from bench import CpuReference
from nim_snappy_b200 import _lib
from nim_snappy_b200.workloads import mixed_corpus, text_corpus
from oracle import oracle as o

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
n = mib << 20
data = text_corpus(n) if kind == "text" else mixed_corpus(n, seed=1)
lib = _lib.default()
enc = np.frombuffer(o.encode(data), dtype=np.uint8)
h_in = torch.from_numpy(enc.copy()).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
w = C.c_size_t(0)
for it in range(4):
    n0, l0 = lib.snb_raw_block_decodes(), lib.snb_kernel_launches()
    t0 = time.perf_counter()
    st = lib.snb_uncompress(h_in.data_ptr(), enc.size, h_out.data_ptr(), n, C.byref(w))
    dt = time.perf_counter() - t0
    print(f"uncompress {kind} {mib} MiB: status {st}, {n / dt / 1e9:.2f} GB/s ({dt * 1e3:.1f} ms), block-parallel={lib.snb_raw_block_decodes() - n0}, launches={lib.snb_kernel_launches() - l0}", flush=True)
assert w.value == n and bool((h_out.numpy() == data).all())
cap = int(lib.snb_max_compressed_len(n)) if hasattr(lib, "snb_max_compressed_len") else n + n // 6 + 32
h_src = torch.from_numpy(data.copy()).pin_memory()
h_enc = torch.empty(n + n // 6 + 64, dtype=torch.uint8).pin_memory()
for it in range(3):
    t0 = time.perf_counter()
    st = lib.snb_compress(h_src.data_ptr(), n, h_enc.data_ptr(), h_enc.numel(), C.byref(w))
    dt = time.perf_counter() - t0
    print(f"compress {kind} {mib} MiB: status {st}, {n / dt / 1e9:.2f} GB/s ({dt * 1e3:.1f} ms)", flush=True)
assert w.value == enc.size and bool((h_enc.numpy()[:w.value] == enc).all())
