#!/usr/bin/env python
"""BASELINE configs 4 and 5 on one B200 (device-resident, CUDA events on the launching stream):
  config 4: batched raw compress / uncompress of N messages of 1-4 KiB (default 1M), per-message parity with the
            oracle on a sample + full GPU round trip;
  config 5: framed compress / uncompress of a mixed corpus slab (text | protobuf-like | random, default 4 GiB),
            byte-compared with the oracle stream built on the host cores.
Prints one JSON line per config.  Usage: python scripts/gpu_configs.py [messages] [mixed_MiB]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as g

g.build()
from bench import CpuReference
from nim_snappy_b200 import _lib
from nim_snappy_b200.codec import maxCompressedLen
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import message_batch, mixed_corpus
from oracle import oracle as o

count = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mixed_mib = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lib = _lib.default()
dc = DeviceCodec(lib)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


# ---------------------------------------------------------------- config 4
if count:
    packed, offs, lens = message_batch(count, seed=0)
    stride = (maxCompressedLen(4096) + 15) // 16 * 16
    d_in, d_off, d_len = torch.from_numpy(packed).cuda(), torch.from_numpy(offs).cuda(), torch.from_numpy(lens).cuda()
    out_off = torch.arange(count, dtype=torch.int64, device="cuda") * stride
    d_out = torch.zeros(count * stride, dtype=torch.uint8, device="cuda")
    d_out_len = torch.empty(count, dtype=torch.int32, device="cuda")
    c_ms = timed(lambda: dc.compress_batch(d_in, d_off, d_len, 4096, d_out, out_off, d_out_len))
    out_len = d_out_len.cpu().numpy()
    rng = np.random.default_rng(1)
    sample = rng.choice(count, size=min(count, 20000), replace=False)
    ok = True
    for i in sample:
        exp = o.encode(packed[offs[i]:offs[i] + lens[i]])
        got = d_out[i * stride:i * stride + int(out_len[i])].cpu().numpy().tobytes() if i in sample[:200] else None
        ok &= out_len[i] == len(exp) and (got is None or got == exp)
    # full byte compare on a contiguous range (one D2H)
    k = min(count, 20000)
    h = d_out[:k * stride].cpu().numpy()
    for i in range(k):
        ok &= h[i * stride:i * stride + out_len[i]].tobytes() == o.encode(packed[offs[i]:offs[i] + lens[i]])
    back = torch.zeros(packed.size, dtype=torch.uint8, device="cuda")
    res = {}

    def dec():
        res["r"] = dc.uncompress_batch(d_out, out_off, d_out_len, back, d_off, d_len)
    d_ms = timed(dec)
    got_len, status = res["r"]
    rt = (not status.cpu().numpy().any()) and bool((back == d_in).all())
    C = int(out_len.sum())
    print(json.dumps({"config": 4, "messages": count, "uncompressed_bytes": int(packed.size), "compressed_bytes": C,
                      "compress_ms": round(c_ms, 3), "compress_GBps": round(packed.size / c_ms / 1e6, 2),
                      "uncompress_ms": round(d_ms, 3), "uncompress_GBps": round(packed.size / d_ms / 1e6, 2),
                      "messages_per_s_compress": round(count / c_ms * 1e3), "parity_vs_oracle": bool(ok),
                      "roundtrip_ok": bool(rt)}), flush=True)
    del d_in, d_out, back, h

# ---------------------------------------------------------------- config 5 (one shard's slab)
if mixed_mib:
    n = mixed_mib << 20
    data = mixed_corpus(n, seed=0)
    ref = CpuReference()
    _, stream = ref.framed_stream(data)
    d_in = torch.from_numpy(data).cuda()
    out = torch.empty(int(lib.snb_max_compressed_len_framed(n)), dtype=torch.uint8, device="cuda")
    w = {}

    def comp():
        w["n"] = dc.compress_framed(d_in, out=out)[1]
    c_ms = timed(comp, reps=3, warm=1)
    ok = w["n"] == stream.size and bool((out[:w["n"]] == torch.from_numpy(stream).cuda()).all())
    chunks, walk = dc.frame_index(stream)
    d_stream = out[:w["n"]]
    d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = {}

    def dec5():
        st["s"] = dc.uncompress_chunks(d_stream, chunks, d_back)
    d_ms = timed(dec5, reps=3, warm=1)
    rt = (not st["s"].any()) and bool((d_back == d_in).all())
    print(json.dumps({"config": 5, "slab_bytes": n, "compressed_bytes": int(stream.size),
                      "compress_ms": round(c_ms, 3), "compress_GBps": round(n / c_ms / 1e6, 2),
                      "uncompress_ms": round(d_ms, 3), "uncompress_GBps": round(n / d_ms / 1e6, 2),
                      "note": "wall of the whole call incl. host header walk for uncompress",
                      "parity_vs_oracle": bool(ok), "roundtrip_ok": bool(rt)}), flush=True)
