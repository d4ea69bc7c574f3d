"""N > 1 host path on CPU: world_size-2 gloo run of the shard partition + length-table gather.
Each rank compresses / decodes its shard with the product sources under the SIMT emulator; the
concatenation must be byte-identical to the single-stream oracle output."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_ranges_cover_and_align():
    from nim_snappy_b200.shard import shard_ranges
    for n in [0, 1, 65535, 65536, 65537, 3 * 65536 + 17, 10 * 65536]:
        for w in [1, 2, 3, 4, 8]:
            rs = shard_ranges(n, w)
            assert len(rs) == w and sum(l for _, l in rs) == n
            pos = 0
            for s, l in rs:
                assert s == pos and (s % 65536 == 0 or s == n)
                pos += l


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import nim_snappy_b200 as s
    from nim_snappy_b200 import _lib, shard
    from nim_snappy_b200.device import CHUNK_DTYPE
    from nim_snappy_b200.workloads import mixed_corpus
    from oracle import oracle as o
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api = s.Snappy(_lib.SnbLibrary(os.path.join(ROOT, "tests", "emu", "libsnb_emu.so")))
    data = mixed_corpus(3 << 20, seed=3)[(1 << 20) - 70000:(1 << 20) + 140000 + 12345]   # text | proto boundary
    ranges = shard.shard_ranges(data.size, world)
    start, length = ranges[rank]
    mine = shard.compress_shard_host(api, data[start:start + length], rank)
    lengths = shard.gather_lengths(len(mine), world)
    offs, total = shard.stream_offsets(lengths)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    stream = b"".join(gathered)
    expect = o.encode_framed(data)
    assert total == len(expect) and stream == expect, (rank, total, len(expect))
    assert stream[offs[rank]:offs[rank] + lengths[rank]] == mine
    # decode side: split the walked chunk table by uncompressed bytes, every rank decodes its slice
    from nim_snappy_b200.device import DeviceCodec
    dc = DeviceCodec(api.lib)
    chunks, walk = dc.frame_index(np.frombuffer(stream, dtype=np.uint8))
    assert walk.status == 0 and walk.written == data.size
    first, count = shard.split_chunks(chunks, world)[rank]
    sl, ib, ie, ob, oe = shard.rebase_chunks(chunks[first:first + count])
    # host API on the slice: a header-less sub-stream of whole chunks (checkHeader = false)
    hdr_begin = ib - 8 if count else 0
    out = np.zeros(max(oe - ob, 1), dtype=np.uint8)[: oe - ob]
    r = api.uncompressFramed(stream[hdr_begin:ie], out, checkHeader=False) if count else (0, 0)
    assert r == (ie - hdr_begin, oe - ob)
    parts = [None] * world
    dist.all_gather_object(parts, out.tobytes())
    assert b"".join(parts) == data.tobytes()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_roundtrip(emu_api, tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
