#!/usr/bin/env python
"""BASELINE config 5: the 64 GiB mixed corpus (text | protobuf-like | incompressible, repeating) framed-compressed
and -uncompressed, sharded contiguously over N GPUs (one rank per GPU; no data-path collective, only the per-shard
byte counts and the times are gathered).  Each rank works through its shard in 4 GiB device slabs; a slab is a
1 GiB pattern generated on the host (seeded by rank) and repeated on the device -- 64 GiB do not fit the host.
Timed with CUDA events per slab (kernels only, data resident), max over ranks.  Checks: full GPU round trip of
every slab; oracle byte-compare of the first slab's first 256 MiB on rank 0.

  python scripts/gpu_config5_sharded.py [total_GiB]                      # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 ... scripts/gpu_config5_sharded.py 64
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

import __graft_entry__ as g

if rank == 0:
    g.build()
if world > 1:
    dist.barrier()
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import mixed_corpus

total_gib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
shard = (total_gib << 30) // world                      # contiguous, a multiple of 64 KiB
SLAB = min(4 << 30, shard)
nslab = (shard + SLAB - 1) // SLAB
lib = _lib.default()
dc = DeviceCodec(lib)
pattern = mixed_corpus(1 << 30, seed=rank)
d_pat = torch.from_numpy(pattern).cuda()
d_in = d_pat.repeat(SLAB >> 30) if SLAB >= (1 << 30) else d_pat[:SLAB].clone()
out = torch.empty(int(lib.snb_max_compressed_len_framed(SLAB)), dtype=torch.uint8, device="cuda")
d_back = torch.empty(SLAB, dtype=torch.uint8, device="cuda")


def ev():
    return torch.cuda.Event(enable_timing=True)


c_ms = d_ms = 0.0
comp_bytes = 0
ok = True
dc.compress_framed(d_in, out=out)                       # warm-up (allocations, clocks)
for s in range(nslab):
    n = min(SLAB, shard - s * SLAB)
    src = d_in[:n]
    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
    torch.cuda.synchronize()
    e0.record()
    _, w = dc.compress_framed(src, out=out, with_header=(rank == 0 and s == 0))
    e1.record()
    torch.cuda.synchronize()
    stream = out[:w]
    h_stream = stream.cpu().numpy()
    chunks, walk = dc.frame_index(h_stream, check_header=(rank == 0 and s == 0))    # host header walk, not timed
    torch.cuda.synchronize()
    e2.record()
    st = dc.uncompress_chunks(stream, chunks, d_back[:n])
    e3.record()
    torch.cuda.synchronize()
    c_ms += e0.elapsed_time(e1)
    d_ms += e2.elapsed_time(e3)
    comp_bytes += int(w)
    ok = ok and (not st.any()) and bool((d_back[:n] == src).all())
parity = None
if rank == 0:
    from oracle import oracle as o
    piece = pattern[: 256 << 20]
    _, w = dc.compress_framed(d_pat[: 256 << 20], out=out, with_header=True)
    parity = out[:w].cpu().numpy().tobytes() == o.encode_framed(piece)
t = torch.tensor([c_ms, d_ms], dtype=torch.float64, device="cuda")
b = torch.tensor([comp_bytes, int(ok)], dtype=torch.int64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)            # device-timed, max over ranks
    tb = [torch.zeros_like(b) for _ in range(world)]
    dist.all_gather(tb, b)                              # the per-shard length table: the only thing shards exchange
    comp_total, all_ok = sum(int(x[0]) for x in tb), all(int(x[1]) for x in tb)
else:
    comp_total, all_ok = comp_bytes, ok
if rank == 0:
    tot = total_gib << 30
    print(json.dumps({"config": 5, "n_gpus": world, "total_bytes": tot, "shard_bytes": shard, "slab_bytes": SLAB,
                      "compressed_bytes": comp_total, "compress_ms_max_over_ranks": round(float(t[0]), 2),
                      "compress_GBps": round(tot / float(t[0]) / 1e6, 1), "uncompress_ms_max_over_ranks": round(float(t[1]), 2),
                      "uncompress_GBps": round(tot / float(t[1]) / 1e6, 1), "roundtrip_ok": bool(all_ok),
                      "parity_vs_oracle_256MiB_rank0": parity}), flush=True)
if world > 1:
    dist.destroy_process_group()
