#!/usr/bin/env python
"""Device-side header walk alone: time and outcome of snb_dev_frame_index on the text and mixed corpora
(with SNB_DEBUG_INDEX=1 the library prints candidate / chain counts).  Usage: python scripts/gpu_index_probe.py [MiB]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import mixed_corpus, text_corpus
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = _lib.default(); dc = DeviceCodec(lib)
for name, gen in (("text", lambda: text_corpus(mib << 20)), ("mixed", lambda: mixed_corpus(mib << 20, seed=0))):
    d_in = torch.from_numpy(gen()).cuda()
    lens = torch.zeros(d_in.numel() // 65536, dtype=torch.int32, device="cuda")
    out, w = dc.compress_framed(d_in, frame_lens=lens)
    s = out[:w]
    for hint in (None, lens):
        f0 = lib.snb_frame_index_fallbacks(); l0 = lib.snb_kernel_launches()
        dc.dev_frame_index(s, frame_lens=hint)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            ch, walk = dc.dev_frame_index(s, frame_lens=hint)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print(f"{name} {mib} MiB stream {w} B, hint={'frame_lens' if hint is not None else 'none'}: {dt * 1e3:.3f} ms per index, "
              f"{len(ch)} chunks, walk=({walk.status},{walk.read},{walk.written}), fallbacks {lib.snb_frame_index_fallbacks() - f0}, "
              f"launches per call {(lib.snb_kernel_launches() - l0) / 6:.0f}", flush=True)
    del d_in, out
