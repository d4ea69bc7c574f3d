"""Sharding of the framed path over the GPUs of one box (one process per GPU).

Framed chunks are independent (fresh hash table per block encoder.nim:237-238, per-chunk CRC and
header), so a corpus is split into contiguous ranges of 64 KiB chunks, one range per rank; there
is NO data-path collective.  The only exchange is the per-shard length table (a few bytes per
rank, `torch.distributed.all_gather`), from which every rank learns where its shard sits in the
final stream.  Shard 0 carries the 10-byte stream identifier (snappy.nim:142).
"""
import numpy as np

CHUNK = 65536


def shard_ranges(nbytes, world):
    """Contiguous (start, length) byte ranges on 64 KiB boundaries, chunk counts as even as possible."""
    nchunks = (nbytes + CHUNK - 1) // CHUNK
    base, extra = divmod(nchunks, world)
    out, c0 = [], 0
    for r in range(world):
        c = base + (1 if r < extra else 0)
        start = min(c0 * CHUNK, nbytes)
        end = min((c0 + c) * CHUNK, nbytes)
        out.append((start, end - start))
        c0 += c
    return out


def gather_lengths(local_len, world, device="cpu", group=None):
    """all_gather of one int64 per rank -> list of per-shard byte counts."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [int(local_len)]
    mine = torch.tensor([int(local_len)], dtype=torch.int64, device=device)
    table = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(table, mine, group=group)
    return [int(x) for x in table.cpu().tolist()]


def stream_offsets(lengths):
    """Byte offset of every shard's output in the concatenated stream, plus the total."""
    offs = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=offs[1:])
    return [int(x) for x in offs[:-1]], int(offs[-1])


def compress_shard_host(api, shard_bytes, rank):
    """compressFramed of one shard through the host API; non-zero ranks drop the stream identifier."""
    stream = api.encodeFramed(shard_bytes)
    return stream if rank == 0 else stream[10:]


def split_chunks(chunks, world):
    """Partition a walked chunk table (device.CHUNK_DTYPE) into `world` contiguous slices of nearly
    equal uncompressed size.  Returns [(first, count)]."""
    n = len(chunks)
    if n == 0:
        return [(0, 0)] * world
    ends = np.cumsum(chunks["out_cap"].astype(np.int64))
    total = int(ends[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(ends, total * r // world, side="left")))
    cuts.append(n)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(world)]


def rebase_chunks(chunks):
    """Slice-local copy of a chunk slice: offsets relative to the slice's first input / output byte.
    Returns (rebased chunks, in_begin, in_end, out_begin, out_end)."""
    if len(chunks) == 0:
        return chunks.copy(), 0, 0, 0, 0
    c = chunks.copy()
    in_begin, out_begin = int(c["in_off"][0]), int(c["out_off"][0])
    in_end = int(c["in_off"][-1]) + int(c["in_len"][-1])
    out_end = int(c["out_off"][-1]) + int(c["out_cap"][-1])
    c["in_off"] -= in_begin
    c["out_off"] -= out_begin
    return c, in_begin, in_end, out_begin, out_end
