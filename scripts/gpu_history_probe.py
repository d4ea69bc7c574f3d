#!/usr/bin/env python
"""What the encoder's history misses cost: 16 384 block-encode jobs of 64 KiB through snb_dev_compress_batch whose
inputs are spread over `wrap` distinct chunks of the text corpus (16 384: every job its own chunk, the normal case;
512: the histories of all jobs fit the L2 together; 1: one chunk).  The hash tables stay private per warp in every case.
Usage: python scripts/gpu_history_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import text_corpus
n = 1 << 30
jobs = n >> 16
lib = _lib.default(); dc = DeviceCodec(lib)
d_in = torch.from_numpy(text_corpus(n)).cuda()
stride = int(lib.snb_max_compressed_len(65536)) + 54
stride = (stride + 127) // 128 * 128
d_out = torch.empty(jobs * stride, dtype=torch.uint8, device="cuda")
d_out_off = (torch.arange(jobs, dtype=torch.int64, device="cuda") * stride)
d_len = torch.full((jobs,), 65536, dtype=torch.int32, device="cuda")
for wrap in (16384, 4096, 1024, 512, 64, 1):
    # jobs that run at the same time should not share a chunk's cache lines more than the wrap forces them to
    idx = (torch.arange(jobs, dtype=torch.int64, device="cuda") * 7919) % wrap
    d_off = idx * 65536
    ms = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dc.compress_batch(d_in, d_off, d_len, 65536, d_out, d_out_off); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(f"wrap {wrap:6d} distinct chunks ({wrap * 64 >> 10:5d} MiB of input): {min(ms):7.3f} ms -> {n / min(ms) / 1e6:6.2f} GB/s", flush=True)
