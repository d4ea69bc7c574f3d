"""Synthetic workloads of BASELINE.json's configs, built from the vendored Google-snappy corpus
(tests/data).  Deterministic: no seed for the tiled corpora, numpy default_rng(seed) elsewhere."""
import os

import numpy as np

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "data")
TEXT_FILES = ["alice29.txt", "asyoulik.txt", "lcet10.txt", "plrabn12.txt", "html", "urls.10K"]
PROTO_FILES = ["geo.protodata", "kppkn.gtb"]


def _cat(names):
    parts = []
    for n in names:
        with open(os.path.join(DATA, n), "rb") as f:
            parts.append(np.frombuffer(f.read(), dtype=np.uint8))
    return np.concatenate(parts)


def tile(unit, nbytes, phase=0):
    """`unit` repeated and cut to nbytes, starting `phase` bytes into the unit."""
    reps = (nbytes + phase) // unit.size + 2
    return np.tile(unit, reps)[phase:phase + nbytes].copy()


def text_corpus(nbytes, phase=0):
    """Config 2/3: T = alice29 || asyoulik || lcet10 || plrabn12 || html || urls.10K (1 994 492 B) tiled and
    truncated; the tile length is not a multiple of 64 KiB so chunk phases vary."""
    return tile(_cat(TEXT_FILES), nbytes, phase)


def mixed_corpus(nbytes, seed=0, phase_mib=0):
    """Config 5: repeating 3 MiB pattern of 1 MiB text, 1 MiB protobuf-like, 1 MiB incompressible."""
    mib = 1 << 20
    text, proto = _cat(TEXT_FILES), _cat(PROTO_FILES)
    rng = np.random.default_rng(seed)
    out = np.empty(nbytes, dtype=np.uint8)
    pos, k = 0, phase_mib
    while pos < nbytes:
        n = min(mib, nbytes - pos)
        kind = k % 3
        if kind == 0:
            out[pos:pos + n] = tile(text, n, (k * 7919) % text.size)
        elif kind == 1:
            out[pos:pos + n] = tile(proto, n, (k * 104729) % proto.size)
        else:
            out[pos:pos + n] = rng.integers(0, 256, n, dtype=np.uint8)
        pos += n
        k += 1
    return out


def message_batch(count, seed=0, lo=1024, hi=4096, pool_bytes=64 << 20):
    """Config 4: `count` messages, lengths ~U[lo, hi], each a slice of a pool (geo.protodata || T tiled).
    Returns (packed bytes, int64 offsets, int32 lengths)."""
    rng = np.random.default_rng(seed)
    pool = tile(np.concatenate([_cat(["geo.protodata"]), _cat(TEXT_FILES)]), pool_bytes)
    lens = rng.integers(lo, hi + 1, count).astype(np.int32)
    starts = rng.integers(0, pool_bytes - hi, count)
    offs = np.zeros(count, dtype=np.int64)
    np.cumsum(lens[:-1], out=offs[1:])
    packed = np.empty(int(lens.sum()), dtype=np.uint8)
    for i in range(count):
        packed[offs[i]:offs[i] + lens[i]] = pool[starts[i]:starts[i] + lens[i]]
    return packed, offs, lens
