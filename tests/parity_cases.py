"""Parity checks shared by the emulator suite (CPU, small inputs) and the GPU suite (full sizes).

`api` is a nim_snappy_b200.Snappy bound to either library; `o` is the oracle module.  Every check
compares the product's bytes / statuses with the oracle's on the same input -- bit-exact, no
tolerance: the whole path is integer / byte work."""
import numpy as np

import vectors as V
from conftest import read_data, read_golden

HDR = b"\xff\x06\x00\x00sNaPpY"


def check_crc(api, o, sizes, seed=3):
    rng = np.random.default_rng(seed)
    for n in sizes:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert api.maskedCrc(d) == o.masked_crc32c(d), n


def check_raw(api, o, data, cpp=None):
    """snappy.encode == oracle encode (byte-identical); decode of the oracle stream == source."""
    enc_o = o.encode(data)
    enc = api.encode(data)
    assert enc == enc_o, (len(data), len(enc), len(enc_o))
    assert api.decode(enc_o) == data
    assert api.uncompressedLen(enc_o) == len(data)
    if cpp is not None and len(data):
        assert api.decode(cpp.encode(data)) == data           # tests/test_snappy.nim:68


def check_framed(api, o, data):
    enc_o = o.encode_framed(data)
    enc = api.encodeFramed(data)
    assert enc == enc_o, (len(data), len(enc), len(enc_o))
    assert api.uncompressedLenFramed(enc_o) == len(data)
    out = np.zeros(max(len(data), 1), dtype=np.uint8)[: len(data)]
    assert api.uncompressFramed(enc_o, out) == (len(enc_o), len(data))
    assert out.tobytes() == data
    assert api.decodeFramed(enc_o) == data


def status_of(r):
    return 0 if not hasattr(r, "name") else int(r)


def check_raw_decode_parity(api, o, stream, cap=None):
    """Same status and (on success) same bytes as the oracle for an arbitrary raw stream."""
    n = o.uncompressed_len(stream)
    cap = (n if n is not None and n < (1 << 24) else 64) if cap is None else cap
    st_o, out_o = o.uncompress(stream, cap)
    buf = np.zeros(max(cap, 1), dtype=np.uint8)[:cap]
    r = api.uncompress(stream, buf)
    assert status_of(r) == st_o, (stream[:16], cap, r, st_o)
    if st_o == 0:
        assert r == len(out_o) and buf[:r].tobytes() == out_o


def check_framed_decode_parity(api, o, stream, cap, check_header=True, check_integrity=True):
    st_o, rd_o, wr_o, out_o = o.uncompress_framed(stream, cap, check_header, check_integrity)
    buf = np.zeros(max(cap, 1), dtype=np.uint8)[:cap]
    r = api.uncompressFramed(stream, buf, check_header, check_integrity)
    assert status_of(r) == st_o, (stream[:24], cap, r, st_o)
    if st_o == 0:
        assert r == (rd_o, wr_o)
        assert buf[:wr_o].tobytes() == out_o
    return r


def check_malformed(api, o):
    for enc in V.MALFORMED_RAW:                               # tests/test_snappy.nim:161-218
        assert api.decode(enc) == b""
        check_raw_decode_parity(api, o, enc)
        check_raw_decode_parity(api, o, enc, cap=64)
    for enc, dec in V.HANDBUILT_RAW:                          # :136-148
        assert api.decode(enc) == dec
    assert api.encode(b"") == b"\x00"                         # :163-165
    assert api.decode(b"\x00") == b""
    assert api.decode(b"") == b""
    for enc in (b"\x05", b"\x80\x01", b"\xff\xff\x03"):         # a length and no tags at all (snappy.nim:104-108)
        assert api.decode(enc) == b""
        check_raw_decode_parity(api, o, enc)


def check_framed_fixture(api, o, fixture, source):
    """tests/test_framed.nim:9-59 incl. the partial-output / resume contract."""
    comp, expected = read_golden(fixture), read_data(source)
    assert api.decodeFramed(comp) == expected
    assert api.uncompressedLenFramed(comp) == len(expected)
    check_framed_decode_parity(api, o, comp, len(expected))
    r = check_framed_decode_parity(api, o, comp, len(expected) - 1)
    rd, wr = r
    assert rd < len(comp) and wr < len(expected)
    buf = np.zeros(len(expected) - wr, dtype=np.uint8)
    r2 = api.uncompressFramed(comp[rd:], buf, False)
    assert r2 == (len(comp) - rd, len(expected) - wr)
    assert buf.tobytes() == expected[wr:]


def check_framed_edges(api, o, full=True):
    """tests/test_framed.nim:137-219."""
    out0 = np.zeros(1, dtype=np.uint8)[:0]
    assert api.uncompressFramed(HDR, out0) == (10, 0)         # just a header
    assert api.decodeFramed(HDR) == b""
    buf = bytes(i & 0xFF for i in range(128 * 1024))
    sizes = [0, 1, 10, 16, 17, 18, 65535, 65536, 65537, 131072] if full else [0, 1, 10, 16, 17, 18, 65537]
    for n in sizes:                                           # buffer sizes
        check_framed(api, o, buf[:n])
    data = bytes(65536)
    crc = o.masked_crc32c(data).to_bytes(4, "little")
    comp = o.encode(data)
    framed = HDR + b"\x01" + (len(data) + 4).to_bytes(3, "little") + crc + data
    framed_c = HDR + b"\x00" + (len(comp) + 4).to_bytes(3, "little") + crc + comp
    for f in (framed, framed_c):                              # full uncompressed / compressed chunk
        assert api.decodeFramed(f) == data
        check_framed_decode_parity(api, o, f, 65536)
    zero = bytes(4)
    f0 = HDR + b"\x01" + (len(data) + 4).to_bytes(3, "little") + zero + data
    f0c = HDR + b"\x00" + (len(comp) + 4).to_bytes(3, "little") + zero + comp
    for f in (f0, f0c):                                       # checkIntegrity
        check_framed_decode_parity(api, o, f, 65536)                        # crcMismatch
        check_framed_decode_parity(api, o, f, 65536, check_integrity=False)  # accepted
        assert api.decodeFramed(f, checkIntegrity=False) == data
        assert api.decodeFramed(f) == b""
    for payload, n in V.INVALID_FRAMED:                       # invalid header
        check_framed_decode_parity(api, o, payload, n)
        assert api.decodeFramed(payload) == b""
        assert api.uncompressedLenFramed(payload) is None
    big = bytes(65537)                                        # overlong frame, both chunk types
    bcrc = o.masked_crc32c(big).to_bytes(4, "little")
    bcomp = o.encode(big)
    for f in (HDR + b"\x01" + (len(big) + 4).to_bytes(3, "little") + bcrc + big,
              HDR + b"\x00" + (len(bcomp) + 4).to_bytes(3, "little") + bcrc + bcomp):
        check_framed_decode_parity(api, o, f, len(big))
        assert api.decodeFramed(f) == b""
        assert api.uncompressedLenFramed(f) is None
    # overlong stored chunk with a wrong CRC: crcMismatch outranks invalidInput (snappy.nim:244-251)
    check_framed_decode_parity(api, o, HDR + b"\x01" + (len(big) + 4).to_bytes(3, "little") + zero + big, len(big))
    # unknown / skippable / padding chunks and a repeated stream identifier (snappy.nim:259-263)
    small = o.encode_framed(b"hello hello hello hello hello")[10:]
    check_framed_decode_parity(api, o, HDR + b"\x02\x01\x00\x00x" + small, 64)
    check_framed_decode_parity(api, o, HDR + b"\x80\x01\x00\x00x" + small + b"\xfe\x02\x00\x00pp" + HDR + small, 64)
    check_framed_decode_parity(api, o, HDR + small + b"\x00\x00", 64)       # trailing < 4 bytes
    check_framed_decode_parity(api, o, HDR + b"\x00\x03\x00\x00abc", 64)    # dataLen < 4
    check_framed_decode_parity(api, o, HDR + small[:-1], 64)                # truncated chunk
    check_framed_decode_parity(api, o, small, 64)                           # no header, checkHeader = true
    check_framed_decode_parity(api, o, small, 64, check_header=False)
    check_framed_decode_parity(api, o, HDR + small, 5)                      # output too small: partial (10, 0)
    stored = HDR + b"\x01\x09\x00\x00" + o.masked_crc32c(b"abcde").to_bytes(4, "little") + b"abcde"
    check_framed_decode_parity(api, o, stored + small, 3)                   # stored chunk does not fit
    check_framed_decode_parity(api, o, stored[:-5] + b"abcdX" + small, 3)   # ... and its CRC is wrong
    check_framed_decode_parity(api, o, stored + small, 5)                   # fits, next one does not


def mutate_streams(o, base, rng, count):
    """Differential-fuzz inputs in the spirit of tests/fuzzing/fuzz_block_format.nim: valid streams
    with byte flips, truncations and insertions."""
    out = []
    for _ in range(count):
        b = bytearray(base)
        kind = rng.integers(0, 4)
        if kind == 0 and len(b) > 1:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1 and len(b) > 1:
            b = b[: int(rng.integers(1, len(b)))]
        elif kind == 2:
            i = int(rng.integers(0, len(b) + 1))
            b[i:i] = bytes(rng.integers(0, 256, int(rng.integers(1, 6)), dtype=np.uint8))
        else:
            i = int(rng.integers(0, len(b)))
            b[i] = int(rng.integers(0, 256))
        out.append(bytes(b))
    return out


# ---------------------------------------------------------------- faststreams entry points (snappy/faststreams.nim)
class _NoLenReader:
    """A non-seekable binary source handing out short reads (socket-like)."""

    def __init__(self, data, piece):
        self.d, self.p, self.piece = bytes(data), 0, piece

    def read(self, n=-1):
        k = self.piece if n < 0 else min(n, self.piece)
        b = self.d[self.p:self.p + k]
        self.p += len(b)
        return b


def _fs_decode(fs, stream, check_integrity=True, source=None):
    """-> (bytes written to the output, exception class name | None)."""
    import io
    out = io.BytesIO()
    try:
        fs.uncompressFramed(stream if source is None else source, out, checkIntegrity=check_integrity)
    except Exception as e:  # noqa: BLE001 -- the class name is the thing compared
        return out.getvalue(), type(e).__name__
    return out.getvalue(), None


def check_faststreams(fs, o, datas):
    """compress / compressFramed / uncompressFramed over streams == the oracle's restatement of
    faststreams.nim:20-147 (bytes, and for the decoder the bytes written before an exception + its type)."""
    import io
    for d in datas:
        out = io.BytesIO()
        fs.compress(d, out)                                   # bytes-like input (faststreams.nim:55-57)
        assert out.getvalue() == o.fs_compress(d)[0] == o.encode(d), len(d)
        out = io.BytesIO()
        fs.compress(io.BytesIO(d), out)                       # seekable stream input
        assert out.getvalue() == o.encode(d), len(d)
        want = o.fs_compress_framed(d)
        assert want == o.encode_framed(d)
        out = io.BytesIO()
        fs.compressFramed(d, out)
        assert out.getvalue() == want, len(d)
        out = io.BytesIO()
        fs.compressFramed(_NoLenReader(d, 50001), out)        # length unknown, short reads
        assert out.getvalue() == want, len(d)
        assert _fs_decode(fs, want) == (d, None)
        assert _fs_decode(fs, want, source=_NoLenReader(want, 70001)) == (d, None)
        assert _fs_decode(fs, want, source=io.BytesIO(want)) == (d, None)


def check_faststreams_errors(fs, o):
    """Exception parity (exceptions.nim:3-12) incl. what was written before the failure."""
    small = o.encode_framed(b"hello hello hello hello hello")[10:]
    two = o.encode_framed(bytes(range(256)) * 300)            # two chunks
    zero = bytes(4)
    data = bytes(65536)
    big = bytes(65537)
    cases = [
        b"", HDR[:9],                                         # :88-89 UnexpectedEofError
        b"\x00" + HDR[1:] + small,                            # :91-92 bad identifier
        HDR, HDR + small, HDR + small + small,
        HDR + small + b"\x00",                                # :142-143 trailing bytes
        HDR + small + b"\x00\x00\x00",
        HDR + small[:-1],                                     # :101-102 truncated frame
        HDR + small + b"\x00\xff\xff\xff",                    # :98-99 dataLen > maxCompressedFrameDataLen
        HDR + small + b"\x00\x03\x00\x00abc",                 # :105-106
        HDR + small + b"\x01\x03\x00\x00abc",                 # :120-121
        HDR + small + b"\x02\x01\x00\x00x" + small,           # :131-135 reserved unskippable
        HDR + b"\x80\x01\x00\x00x" + small + b"\xfe\x02\x00\x00pp" + HDR + small,   # skippable + repeated identifier
        HDR + small[:4] + zero + small[8:],                   # CRC mismatch, compressed
        HDR + b"\x01" + (len(data) + 4).to_bytes(3, "little") + zero + data,          # CRC mismatch, stored
        HDR + small + b"\x01" + (len(big) + 4).to_bytes(3, "little") + o.masked_crc32c(big).to_bytes(4, "little") + big,
        HDR + small + b"\x00\x08\x00\x00" + zero + b"\x05\xff\xff\xff",               # malformed body after a good chunk
        two[:-3], two + small[:5],
        two[:10 + 4] + zero + two[10 + 8:],                   # first chunk's CRC wrong
    ]
    for i, s in enumerate(cases):
        for ci in (True, False):
            want = o.fs_uncompress_framed(s, ci)
            got = _fs_decode(fs, s, ci)
            assert got == want, (i, ci, got[1], want[1], len(got[0]), len(want[0]))
        assert _fs_decode(fs, s, source=_NoLenReader(s, 7)) == o.fs_uncompress_framed(s, True), i


# ---------------------------------------------------------------- raw streams longer than one block
def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _lit(b):
    """One literal element holding b (decoder.nim:39-79 encodings)."""
    n = len(b) - 1
    if n < 60:
        return bytes([n << 2]) + b
    k = (n.bit_length() + 7) // 8
    return bytes([(59 + k) << 2]) + n.to_bytes(k, "little") + b


def check_raw_multiblock(api, o, cpp, datas, fuzz=6):
    """`uncompress` of raw streams of several blocks: the block-discovery path (raw_walk_kernel) and its fall-backs
    give the oracle's bytes and statuses."""
    rng = np.random.default_rng(5)
    for d in datas:
        enc = o.encode(d)
        n0 = api.lib.snb_raw_block_decodes()
        assert api.decode(enc) == d                                   # reference-encoded, block structured
        if cpp is not None:
            assert api.decode(cpp.encode(d)) == d                     # C++ snappy blocks (tests/test_snappy.nim:68)
        assert api.lib.snb_raw_block_decodes() - n0 == (2 if cpp is not None else 1)   # ... and decoded block-parallel
        for s in (enc[:-1], enc[:len(enc) // 2], enc + b"\x00", _varint(len(d) + 1) + enc[len(_varint(len(d))):],
                  _varint(len(d) - 1) + enc[len(_varint(len(d))):]):
            check_raw_decode_parity(api, o, s)                        # truncated / trailing / wrong announced length
        for s in mutate_streams(o, enc, rng, fuzz):
            check_raw_decode_parity(api, o, s)
    # streams no block-wise encoder would write, all valid: they must take the single-warp route and still decode
    a = bytes(rng.integers(0, 256, 65536, dtype=np.uint8))
    b = bytes(rng.integers(0, 7, 3000, dtype=np.uint8))
    straddle = _varint(65536 + 3000) + _lit(a[:65000]) + _lit(a[65000:] + b)            # an element across the 64 KiB line
    n0 = api.lib.snb_raw_block_decodes()
    assert o.decode(straddle) == a + b and api.decode(straddle) == a + b
    back = _varint(65536 + 64 + 10) + _lit(a) + bytes([(63 << 2) | 3]) + (65536).to_bytes(4, "little") + _lit(b[:10])
    assert o.decode(back) == a + a[:64] + b[:10] and api.decode(back) == a + a[:64] + b[:10]   # copy4 into the previous block
    near = _varint(65536 + 64) + _lit(a) + bytes([(63 << 2) | 2]) + (65535).to_bytes(2, "little")
    assert o.decode(near) == a + a[1:65] and api.decode(near) == a + a[1:65]                   # copy2 across the line
    far = _varint(65536 + 64) + _lit(a) + bytes([(63 << 2) | 3]) + (65537).to_bytes(4, "little")
    check_raw_decode_parity(api, o, far)                                                         # offset beyond the start
    assert api.lib.snb_raw_block_decodes() == n0


def random_element_stream(rng, target, max_lit=3000, far=True):
    """A VALID raw stream assembled from random elements (decoder.nim:39-151 encodings), not from an encoder: literals of
    any length, copy1 / copy2 / copy4 with arbitrary in-range offsets -- elements straddle 64 KiB lines and copies reach
    across them at random, which no block-wise encoder does."""
    out = bytearray()
    op = 0
    while op < target:
        kind = int(rng.integers(0, 4)) if op >= 4 else 0
        if kind == 0:
            n = int(rng.integers(1, max_lit)) if rng.random() < 0.9 else int(rng.integers(60000, 70000))
            out += _lit(bytes(rng.integers(0, 256, n, dtype=np.uint8)))
            op += n
        elif kind == 1:
            off = int(rng.integers(1, min(op, 2047) + 1)); ln = int(rng.integers(4, 12))
            out += bytes([((off >> 8) << 5) | ((ln - 4) << 2) | 1, off & 0xff]); op += ln
        elif kind == 2:
            off = int(rng.integers(1, min(op, 65535) + 1)); ln = int(rng.integers(1, 65))
            out += bytes([((ln - 1) << 2) | 2]) + off.to_bytes(2, "little"); op += ln
        else:
            off = int(rng.integers(1, op + 1)) if far else int(rng.integers(1, min(op, 65535) + 1)); ln = int(rng.integers(1, 65))
            out += bytes([((ln - 1) << 2) | 3]) + off.to_bytes(4, "little"); op += ln
    return _varint(op) + bytes(out), op


def check_random_element_streams(api, o, seeds, target):
    for seed in seeds:
        rng = np.random.default_rng(seed)
        s, n = random_element_stream(rng, target)
        want = o.decode(s, 1 << 30)
        assert len(want) == n
        assert api.decode(s, 1 << 30) == want, seed
        cut = int(rng.integers(len(s) // 2, len(s)))
        check_raw_decode_parity(api, o, s[:cut])
        b = bytearray(s); b[int(rng.integers(5, len(b)))] ^= 1 << int(rng.integers(0, 8))
        check_raw_decode_parity(api, o, bytes(b))


# ---------------------------------------------------------------- host-buffer message batches (snb_compress_batch / snb_uncompress_batch)
def check_batch(api, o, messages, extra_streams=()):
    """Every message of a batch: stream == oracle `encode` byte for byte; batch decode of the oracle's streams plus
    `extra_streams` (malformed ones) == oracle `uncompress` status and bytes per message."""
    import ctypes as C
    enc = api.encodeBatch(messages)
    exp = [o.encode(m) for m in messages]
    assert enc == exp
    streams = exp + list(extra_streams)
    dec = api.decodeBatch(streams)
    for s, d in zip(streams, dec):
        n = o.uncompressed_len(s)
        st_o, out_o = o.uncompress(s, n if n is not None and n < (1 << 24) else 0) if n is not None and n <= 0xFFFFFFFF else (2, b"")
        assert d == (out_o if st_o == 0 else b""), (s[:12], st_o)
    # explicit placement + per-message room: bufferTooSmall for exactly the short ones (snappy.nim:96-97)
    lib = api.lib
    lens = np.array([len(s) for s in exp], dtype=np.uint32)
    offs = np.zeros(len(exp), dtype=np.uint64)
    if len(exp) > 1:
        np.cumsum(lens[:-1], out=offs[1:], dtype=np.uint64)
    packed = np.frombuffer(b"".join(exp) or b"\0", dtype=np.uint8)
    ulen = np.array([len(m) for m in messages], dtype=np.uint32)
    cap = ulen.copy()
    short = np.arange(len(exp)) % 3 == 0
    cap[short & (ulen > 0)] -= 1
    stride = int(ulen.max()) + 8 if len(exp) else 8
    ooff = (np.arange(len(exp), dtype=np.uint64) * stride).astype(np.uint64)
    out = np.zeros(max(1, len(exp) * stride), dtype=np.uint8)
    out_len = np.zeros(len(exp), dtype=np.uint32)
    status = np.full(len(exp), -1, dtype=np.int32)
    st = lib.check(lib.snb_uncompress_batch(packed.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(exp), out.ctypes.data,
                                            out.size, ooff.ctypes.data, cap.ctypes.data, out_len.ctypes.data, status.ctypes.data))
    assert st == 0
    for i, m in enumerate(messages):
        if short[i] and len(m) > 0:
            assert status[i] == 1, i
        else:
            assert status[i] == 0 and out_len[i] == len(m) and out[i * stride:i * stride + len(m)].tobytes() == bytes(m), i
    # compress with explicit offsets (slots of the worst-case size)
    mlens = np.array([len(m) for m in messages], dtype=np.uint32)
    moffs = np.zeros(len(messages), dtype=np.uint64)
    if len(messages) > 1:
        np.cumsum(mlens[:-1], out=moffs[1:], dtype=np.uint64)
    mpacked = np.frombuffer(b"".join(bytes(m) for m in messages) or b"\0", dtype=np.uint8)
    cstride = 32 + int(mlens.max()) + int(mlens.max()) // 6 if len(messages) else 32
    coff = (np.arange(len(messages), dtype=np.uint64) * cstride).astype(np.uint64)
    cout = np.zeros(max(1, len(messages) * cstride), dtype=np.uint8)
    clen = np.zeros(len(messages), dtype=np.uint32)
    cst = np.full(len(messages), -1, dtype=np.int32)
    st = lib.check(lib.snb_compress_batch(mpacked.ctypes.data, moffs.ctypes.data, mlens.ctypes.data, len(messages), cout.ctypes.data,
                                          cout.size, coff.ctypes.data, clen.ctypes.data, cst.ctypes.data))
    assert st == 0 and not cst.any()
    for i in range(len(messages)):
        assert cout[i * cstride:i * cstride + clen[i]].tobytes() == exp[i], i
    # too little total room in packed mode: the whole call says bufferTooSmall (snappy.nim:44-45)
    if len(messages):
        tiny = np.zeros(8, dtype=np.uint8)
        assert lib.snb_compress_batch(mpacked.ctypes.data, moffs.ctypes.data, mlens.ctypes.data, len(messages), tiny.ctypes.data,
                                      tiny.size, None, clen.ctypes.data, cst.ctypes.data) == 1


# ---------------------------------------------------------------- device-side header walk (snb_dev_uncompress_framed)
def emu_dev_call(lib):
    """snb_dev_uncompress_framed over host arrays (the emulator's `device` pointers are host pointers)."""
    import ctypes as C

    def call(stream, cap, check_header=True, check_integrity=True, frame_lens=None):
        a = np.frombuffer(bytes(stream) + b"\0" * 16, dtype=np.uint8)
        out = np.zeros(cap + 16, dtype=np.uint8)
        rd, wr = C.c_uint64(0), C.c_uint64(0)
        fl = None if frame_lens is None else np.ascontiguousarray(frame_lens, dtype=np.uint32)
        st = lib.snb_dev_uncompress_framed(a.ctypes.data, len(stream), out.ctypes.data, cap, int(check_header),
                                           int(check_integrity), None if fl is None else fl.ctypes.data,
                                           0 if fl is None else fl.size, C.byref(rd), C.byref(wr), None)
        return lib.check(st), rd.value, wr.value, out[:wr.value].tobytes()
    return call


def gpu_dev_call(dc):
    import torch

    def call(stream, cap, check_header=True, check_integrity=True, frame_lens=None):
        d_s = torch.frombuffer(bytearray(bytes(stream) + b"\0" * 16), dtype=torch.uint8).cuda()[: len(stream)]
        d_out = torch.zeros(cap + 16, dtype=torch.uint8, device="cuda")[:cap]
        fl = None if frame_lens is None else torch.from_numpy(np.ascontiguousarray(frame_lens, dtype=np.int32)).cuda()
        st, rd, wr = dc.uncompress_framed(d_s, d_out, check_header, check_integrity, fl)
        return st, rd, wr, d_out[:wr].cpu().numpy().tobytes()
    return call


def check_dev_framed_decode_parity(call, o, stream, cap, check_header=True, check_integrity=True, frame_lens=None):
    """Same status / (read, written) / bytes as the oracle's uncompressFramed for a device-resident stream.  On a
    per-chunk failure the library additionally reports the intact prefix (see snb.h); the oracle has no counts then."""
    st_o, rd_o, wr_o, out_o = o.uncompress_framed(stream, cap, check_header, check_integrity)
    st, rd, wr, out = call(stream, cap, check_header, check_integrity, frame_lens)
    assert st == st_o, (bytes(stream[:24]), cap, st, st_o)
    if st_o == 0:
        assert (rd, wr) == (rd_o, wr_o), (rd, wr, rd_o, wr_o)
        assert out == out_o


def check_dev_framed(call, o, datas, rng, fuzz=30, lib=None):
    for d in datas:
        s = o.encode_framed(d)
        n = len(d)
        fl = []
        pos = 10
        while pos < len(s):
            ln = 4 + (s[pos + 1] | (s[pos + 2] << 8) | (s[pos + 3] << 16))
            fl.append(ln)
            pos += ln
        for cap in sorted({n, n + 5, max(n - 1, 0), 65536, 65535, 0, n // 2}):
            check_dev_framed_decode_parity(call, o, s, cap)
            check_dev_framed_decode_parity(call, o, s, cap, check_integrity=False)
        # a well-formed stream with enough room does not need the host walk -- unless its payload looks like headers all
        # over (a run of zeros compresses to `fe 01 00` copy elements: every third byte is a plausible skippable chunk),
        # which overflows the candidate slots and is handed to the host walk by design
        if lib is not None and len(s) > 10 and len(set(d)) > 1:
            f0 = lib.snb_frame_index_fallbacks()
            check_dev_framed_decode_parity(call, o, s, n)
            check_dev_framed_decode_parity(call, o, s, max(n - 1, 0))
            check_dev_framed_decode_parity(call, o, s[:10] + b"\xfe\x03\x00\x00abc" + s[10:], n)
            if fl:
                check_dev_framed_decode_parity(call, o, s, n, frame_lens=fl)
            assert lib.snb_frame_index_fallbacks() == f0, "device header walk fell back to the host on a valid stream"
        if fl:
            check_dev_framed_decode_parity(call, o, s, n, frame_lens=fl)                   # the compressor's own table
            check_dev_framed_decode_parity(call, o, s, n, frame_lens=fl[:-1] + [fl[-1] + 1])  # a table that does not fit: ignored
            check_dev_framed_decode_parity(call, o, s[10:], n, check_header=False, frame_lens=fl)
        check_dev_framed_decode_parity(call, o, s[10:], n, check_header=False)
        # skippable / padding / repeated identifier chunks in the middle, unknown chunk types, trailing garbage
        mid = 10 + (fl[0] if fl else 0)
        for ins in (b"\xfe\x03\x00\x00abc", HDR, b"\x80\x00\x00\x00", b"\x02\x01\x00\x00x", b"\x7f\x00\x00\x00"):
            check_dev_framed_decode_parity(call, o, s[:mid] + ins + s[mid:], n)
        for tail in (b"\x00", b"\xff\x06\x00", b"\x00\x10\x00\x00zz", b"\x01\x02\x00\x00ab"):
            check_dev_framed_decode_parity(call, o, s + tail, n)
        for m in mutate_streams(o, s, rng, fuzz):
            check_dev_framed_decode_parity(call, o, m, n)
            check_dev_framed_decode_parity(call, o, m, n, check_integrity=False)
    check_dev_framed_decode_parity(call, o, long_skippable_stream(o, bytes(datas[2]) if len(datas) > 2 else b"x" * 100), 1 << 20)
    for payload, n in V.INVALID_FRAMED:
        check_dev_framed_decode_parity(call, o, payload, n, check_header=False)
        check_dev_framed_decode_parity(call, o, HDR + payload, n)
    check_dev_framed_decode_parity(call, o, b"", 0)
    check_dev_framed_decode_parity(call, o, HDR, 0)
    check_dev_framed_decode_parity(call, o, HDR[:7], 0)


def long_skippable_stream(o, data):
    """A valid stream with a skippable chunk longer than the device scan considers: must still decode (host walk)."""
    s = o.encode_framed(data)
    pad = b"\xfe" + (70000).to_bytes(3, "little") + bytes(70000)
    return s[:10] + pad + s[10:]
