"""the std/streams adapter (snappy/streams.nim:9-41) and a differential fuzz harness in the
spirit of tests/fuzzing/fuzz_block_format.nim:8-33 (product vs oracle vs the bundled C++ snappy)."""
import io

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import parity_cases as P
from conftest import read_data


class _ShortReads:
    """readData stand-in that hands out `piece` bytes at most per call."""

    def __init__(self, data, piece):
        self.d, self.p, self.piece, self.reads = bytes(data), 0, piece, []

    def read(self, n):
        b = self.d[self.p:self.p + min(n, self.piece)]
        self.p += len(b)
        self.reads.append(b)
        return b


def _check_streams(api, o, datas):
    from nim_snappy_b200 import streams
    s = streams.Streams(api.lib)
    for d in datas:
        out = io.BytesIO()
        s.compress(io.BytesIO(d), len(d), out)                 # full reads: the same bytes as `encode`
        assert out.getvalue() == o.encode(d) == o.st_compress([d[i:i + 65536] for i in range(0, len(d), 65536)], len(d))[0]
        src = _ShortReads(d, 40000)                            # every read is a block of its own (streams.nim:31-41)
        out = io.BytesIO()
        s.compress(src, len(d), out)
        want, _ = o.st_compress([r for r in src.reads if r], len(d))
        assert out.getvalue() == want
        assert o.decode(want) == d                             # still one valid raw stream
        out = io.BytesIO()                                     # inputLen beyond the data: header says inputLen, no error
        s.compress(io.BytesIO(d), len(d) + 5, out)
        assert out.getvalue() == o.st_compress([d[i:i + 65536] for i in range(0, len(d), 65536)], len(d) + 5)[0]
    with pytest.raises(streams.InputTooLarge):
        s.compress(io.BytesIO(b""), 1 << 32, io.BytesIO())


def test_emu_streams_adapter(emu_api, oracle):
    alice = read_data("alice29.txt")
    _check_streams(emu_api, oracle, [b"", b"abc", alice[:17], alice[:66000]])


def _differential(api, o, cpp, payload):
    """fuzz_block_format.nim:8-33 with the oracle added: accept/reject and bytes agree everywhere."""
    want = o.decode(payload, 1 << 20)
    got = api.decode(payload, 1 << 20)
    assert got == want
    if cpp is not None and len(want):
        c = cpp.decode(payload)
        assert c == want                                       # "decompression should match between libraries"
        enc = api.encode(got)
        assert enc == o.encode(got)
        assert cpp.decode(enc) == got                          # "cpp should be able to decompress our compressed data"


_valid = st.builds(lambda parts: b"".join(parts), st.lists(
    st.one_of(st.binary(min_size=1, max_size=40),
              st.builds(lambda b, k: b * k, st.binary(min_size=1, max_size=6), st.integers(2, 30))), max_size=12))


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(st.binary(max_size=200))
def test_emu_fuzz_raw_bytes(emu_api, oracle, payload):
    _differential(emu_api, oracle, oracle.ref_snappycpp(), payload)


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(_valid, st.integers(0, 1 << 30))
def test_emu_fuzz_mutated_streams(emu_api, oracle, data, seed):
    rng = np.random.default_rng(seed)
    for s in [oracle.encode(data)] + P.mutate_streams(oracle, oracle.encode(data), rng, 3):
        _differential(emu_api, oracle, oracle.ref_snappycpp(), s)


@pytest.mark.gpu
def test_gpu_streams_adapter(gpu_api, oracle):
    _check_streams(gpu_api, oracle, [b"", b"abc", read_data("alice29.txt"), read_data("geo.protodata"), read_data("house.jpg")])


@pytest.mark.gpu
@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(st.one_of(st.binary(max_size=400), _valid.map(lambda d: d)), st.integers(0, 1 << 30))
def test_gpu_fuzz(gpu_api, oracle, data, seed):
    rng = np.random.default_rng(seed)
    cpp = oracle.ref_snappycpp()
    _differential(gpu_api, oracle, cpp, data)
    enc = oracle.encode(data)
    for s in [enc] + P.mutate_streams(oracle, enc, rng, 2):
        _differential(gpu_api, oracle, cpp, s)
