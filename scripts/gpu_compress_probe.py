#!/usr/bin/env python
"""Device-resident framed compress of the bench corpus, timed with CUDA events: min / median of N launches.
Usage: SNB_LIB=... python scripts/gpu_compress_probe.py [MiB] [text|mixed] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import mixed_corpus, text_corpus
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
n = mib << 20
data = text_corpus(n) if kind == "text" else mixed_corpus(n, seed=0)
lib = _lib.default(); dc = DeviceCodec(lib)
d_in = torch.from_numpy(data).cuda()
out = torch.empty(int(lib.snb_max_compressed_len_framed(n)), dtype=torch.uint8, device="cuda")
for _ in range(3):
    _, w = dc.compress_framed(d_in, out=out)
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _, w = dc.compress_framed(d_in, out=out); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
import hashlib
h = hashlib.sha256(out[:w].cpu().numpy().tobytes()).hexdigest()[:16]
print(f"{os.path.basename(os.environ.get('SNB_LIB', 'libsnb200.so'))} {kind} {mib} MiB: min {min(ms):.3f} ms ({n / min(ms) / 1e6:.2f} GB/s), median {float(np.median(ms)):.3f} ms, written {w}, sha {h}", flush=True)
