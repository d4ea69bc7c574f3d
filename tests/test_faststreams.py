"""The stream entry points (snappy/faststreams.nim:20-147).

CPU: the oracle's restatement against the reference's framed fixtures, and the product host module driven
through the SIMT-emulator build of the product library (tests/emu) -- host logic + kernels' warp logic,
no GPU.  GPU (-m gpu): the same checks through the real CUDA library at multi-slab sizes."""
import io

import numpy as np
import pytest

import parity_cases as P
import vectors as V
from conftest import read_data, read_golden


def _fs(api):
    from nim_snappy_b200 import faststreams
    return faststreams.FastStreams(api.lib)


def test_oracle_faststreams_fixtures(oracle):
    """The restatement of faststreams.nim:85-147 decodes the reference's own framed fixtures
    (tests/test_framed.nim:117-119) and agrees with the in-memory oracle on what is valid."""
    for fixture, source in V.FRAMED_FIXTURES:
        assert oracle.fs_uncompress_framed(read_golden(fixture)) == (read_data(source), None)
    house = read_data("house.jpg")
    assert oracle.fs_compress_framed(house) == read_golden("house.jpg.sz")      # the one pinned encoder fixture
    assert oracle.fs_compress(b"") == (b"\x00", None)                           # tests/test_snappy.nim:163-165
    for payload, _n in V.INVALID_FRAMED:                                        # tests/test_framed.nim:198-219
        assert oracle.fs_uncompress_framed(payload)[1] is not None


def test_emu_faststreams(emu_api, oracle):
    fs = _fs(emu_api)
    alice = read_data("alice29.txt")
    pat = bytes((j * 7) & 0xFF for j in range(70000))
    P.check_faststreams(fs, oracle, [b"", b"a", alice[:17], alice[:5000], pat[:65536], pat[:65537], alice[:70000]])
    fs.close()


def test_emu_faststreams_errors(emu_api, oracle):
    fs = _fs(emu_api)
    P.check_faststreams_errors(fs, oracle)
    fs.close()


def test_input_too_large(emu_api):
    """faststreams.nim:28-33: InputTooLarge before anything is written (length taken from the stream, no data read)."""
    from nim_snappy_b200 import faststreams

    class Huge(io.RawIOBase):
        def seekable(self):
            return True

        def tell(self):
            return 0

        def seek(self, off, whence=0):
            return (1 << 32) if whence == 2 else 0

        def readinto(self, b):
            raise AssertionError("must not read")

    out = io.BytesIO()
    with pytest.raises(faststreams.InputTooLarge):
        _fs(emu_api).compress(Huge(), out)
    assert out.getvalue() == b""
    assert issubclass(faststreams.InputTooLarge, faststreams.SnappyEncodingError)       # exceptions.nim:3-12
    assert issubclass(faststreams.MalformedSnappyData, faststreams.SnappyDecodingError)
    assert issubclass(faststreams.UnexpectedEofError, faststreams.SnappyDecodingError)


@pytest.mark.gpu
def test_gpu_faststreams(gpu_api, oracle):
    fs = _fs(gpu_api)
    datas = [b"", b"a"] + [read_data(n) for n in ("alice29.txt", "html", "house.jpg", "geo.protodata")]
    P.check_faststreams(fs, oracle, datas)
    P.check_faststreams_errors(fs, oracle)
    fs.close()


@pytest.mark.gpu
def test_gpu_faststreams_multi_slab(gpu_api, oracle):
    """150 MiB + 12345 B: three 64 MiB slabs, the last one ragged; stream with unknown length and short reads."""
    from nim_snappy_b200 import workloads
    fs = _fs(gpu_api)
    n = 150 * (1 << 20) + 12345
    d = workloads.text_corpus(n).tobytes()
    want = oracle.encode_framed(d)
    out = io.BytesIO()
    fs.compressFramed(P._NoLenReader(d, 3 << 20), out)
    assert out.getvalue() == want
    out = io.BytesIO()
    fs.compress(d, out)
    assert out.getvalue() == oracle.encode(d)
    assert P._fs_decode(fs, want, source=P._NoLenReader(want, (5 << 20) + 1)) == (d, None)
    bad = bytearray(want)
    bad[len(bad) // 2] ^= 0x40                                  # a flipped byte deep in the second slab
    got = P._fs_decode(fs, bytes(bad))
    assert got == oracle.fs_uncompress_framed(bytes(bad))
    assert got[1] == "MalformedSnappyData" and 0 < len(got[0]) < n and d.startswith(got[0])
    fs.close()


def test_emu_faststreams_seekable_file_over_one_piece(emu_api, oracle):
    """A seekable file object holding more than one 32 MiB window piece: the staging buffers must be sized from
    the bytes still to come INCLUDING what the window has already buffered (round-1 advisor finding: a fresh
    FastStreams raised ValueError on a valid 33.8 MiB stream read from a BytesIO)."""
    rng = np.random.default_rng(21)
    d = rng.integers(0, 256, 33 * (1 << 20) + 777, dtype=np.uint8).tobytes()     # incompressible: stored frames
    stream = oracle.encode_framed(d)
    assert len(stream) > (32 << 20)
    fs = _fs(emu_api)                                                             # fresh: no larger pooled buffers
    out = io.BytesIO()
    fs.uncompressFramed(io.BytesIO(stream), out, checkIntegrity=False)
    assert out.getvalue() == d
    fs.close()
