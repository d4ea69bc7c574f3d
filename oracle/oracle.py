"""ctypes binding of the CPU oracle (oracle/snappy_oracle.c) and of the reference-built
checkers under oracle/_ref.  TEST INFRASTRUCTURE ONLY: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never from the
product package.

Status codes: 0 ok, else 1 + ord(CodecError | FrameError) (snappy/codec.nim:56-64).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OK, BUFFER_TOO_SMALL, INVALID_INPUT, CRC_MISMATCH, UNKNOWN_CHUNK = range(5)


def build(quiet=True):
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


def _load(path):
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


_lib = None


def lib():
    global _lib
    if _lib is None:
        p = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(p):
            build()
        _lib = C.CDLL(p)
        u8p, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
        _lib.orc_masked_crc32c.restype = C.c_uint32
        _lib.orc_masked_crc32c.argtypes = [u8p, sz]
        _lib.orc_max_compressed_len.restype = u64
        _lib.orc_max_compressed_len.argtypes = [u64]
        _lib.orc_max_compressed_len_framed.restype = u64
        _lib.orc_max_compressed_len_framed.argtypes = [C.c_int64]
        _lib.orc_encode_block_ex.restype = sz
        _lib.orc_encode_block_ex.argtypes = [u8p, sz, u8p, C.c_int, C.c_int]
        _lib.orc_encode_frame.restype = sz
        _lib.orc_encode_frame.argtypes = [u8p, sz, u8p]
        for f in (_lib.orc_compress, _lib.orc_compress_framed):
            f.restype = C.c_int
            f.argtypes = [u8p, u64, u8p, u64, C.POINTER(u64)]
        _lib.orc_encode_frames_range.restype = u64
        _lib.orc_encode_frames_range.argtypes = [u8p, u64, u64, u64, u8p]
        for f in (_lib.orc_decode_all_tags, _lib.orc_uncompress):
            f.restype = C.c_int
            f.argtypes = [u8p, sz, u8p, sz, C.POINTER(sz)]
        _lib.orc_uncompress_framed.restype = C.c_int
        _lib.orc_uncompress_framed.argtypes = [u8p, sz, u8p, sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]
        _lib.orc_uncompressed_len.restype = C.c_int64
        _lib.orc_uncompressed_len.argtypes = [u8p, sz, C.POINTER(u64)]
        _lib.orc_uncompressed_len_framed.restype = C.c_int64
        _lib.orc_uncompressed_len_framed.argtypes = [u8p, sz]
        _lib.orc_uncompress_frames_range.restype = C.c_int
        _lib.orc_uncompress_frames_range.argtypes = [u8p, sz, sz, sz, u8p, sz, C.POINTER(sz), C.POINTER(sz)]
    return _lib


def _arr(b):
    """bytes-like -> contiguous uint8 numpy array (no copy for arrays)."""
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data if a.size else None


def masked_crc32c(data):
    a = _arr(data)
    return lib().orc_masked_crc32c(_ptr(a), a.size)


def max_compressed_len(n):
    return lib().orc_max_compressed_len(n)


def max_compressed_len_framed(n):
    return lib().orc_max_compressed_len_framed(n)


def encode_block(data, hash_bits=14, exit_ge=0):
    a = _arr(data)
    out = np.empty(max_compressed_len(a.size), dtype=np.uint8)
    n = lib().orc_encode_block_ex(_ptr(a), a.size, _ptr(out), hash_bits, exit_ge)
    return out[:n].tobytes()


def compress(data, cap=None):
    """snappy.nim:27 compress -> (status, bytes)."""
    a = _arr(data)
    cap = max_compressed_len(a.size) if cap is None else cap
    out = np.empty(max(cap, 1), dtype=np.uint8)
    w = C.c_uint64(0)
    st = lib().orc_compress(_ptr(a), a.size, _ptr(out), cap, C.byref(w))
    return st, out[: w.value].tobytes()


def encode(data):
    st, b = compress(data)
    return b if st == OK else b""


def compress_framed(data, cap=None, as_array=False):
    a = _arr(data)
    cap = max_compressed_len_framed(a.size) if cap is None else cap
    out = np.empty(max(cap, 1), dtype=np.uint8)
    w = C.c_uint64(0)
    st = lib().orc_compress_framed(_ptr(a), a.size, _ptr(out), cap, C.byref(w))
    res = out[: w.value]
    return st, (res if as_array else res.tobytes())


def encode_framed(data):
    st, b = compress_framed(data)
    return b if st == OK else b""


def uncompress(data, cap):
    a = _arr(data)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    w = C.c_size_t(0)
    st = lib().orc_uncompress(_ptr(a), a.size, _ptr(out), cap, C.byref(w))
    return st, out[: w.value].tobytes()


def uncompressed_len(data):
    a = _arr(data)
    v = C.c_uint64(0)
    r = lib().orc_uncompressed_len(_ptr(a), a.size, C.byref(v))
    return None if r < 0 else v.value


def decode(data, max_size=0xFFFFFFFF):
    n = uncompressed_len(data)
    if n is None or n > max_size:
        return b""
    st, b = uncompress(data, n)
    return b if st == OK else b""


def uncompress_framed(data, cap, check_header=True, check_integrity=True, as_array=False):
    """snappy.nim:169 -> (status, read, written, output[:written])."""
    a = _arr(data)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    r, w = C.c_size_t(0), C.c_size_t(0)
    st = lib().orc_uncompress_framed(_ptr(a), a.size, _ptr(out), cap, int(check_header), int(check_integrity),
                                     C.byref(r), C.byref(w))
    res = out[: w.value]
    return st, r.value, w.value, (res if as_array else res.tobytes())


def uncompressed_len_framed(data):
    a = _arr(data)
    r = lib().orc_uncompressed_len_framed(_ptr(a), a.size)
    return None if r < 0 else r


def decode_framed(data, max_size=None, check_integrity=True):
    n = uncompressed_len_framed(data)
    if n is None or (max_size is not None and n > max_size):
        return b""
    st, r, w, b = uncompress_framed(data, n, True, check_integrity)
    return b if st == OK else b""


# ------------------------------------------------------------------ reference-built checkers
def ref_crc():
    """masked_crc32c from the reference's own snappy/crc32c.c (oracle/_ref), or None."""
    l = _load(os.path.join(HERE, "_ref", "libref_crc32c.so"))
    if l is None:
        return None
    l.masked_crc32c.restype = C.c_uint32
    l.masked_crc32c.argtypes = [C.c_void_p, C.c_size_t]

    def f(data):
        a = _arr(data)
        return l.masked_crc32c(_ptr(a), a.size)

    return f


class RefSnappyCpp:
    """Bundled Google C++ snappy through its C API (tests/cpp_snappy.nim:8-45)."""

    def __init__(self, l):
        self.l = l
        l.snappy_compress.restype = C.c_int
        l.snappy_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
        l.snappy_uncompress.restype = C.c_int
        l.snappy_uncompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
        l.snappy_max_compressed_length.restype = C.c_size_t
        l.snappy_max_compressed_length.argtypes = [C.c_size_t]
        l.snappy_uncompressed_length.restype = C.c_int
        l.snappy_uncompressed_length.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]

    def encode(self, data):
        a = _arr(data)
        n = C.c_size_t(self.l.snappy_max_compressed_length(a.size))
        out = np.empty(n.value, dtype=np.uint8)
        if self.l.snappy_compress(_ptr(a), a.size, _ptr(out), C.byref(n)) != 0:
            raise ValueError("Cannot compress")
        return out[: n.value].tobytes()

    def decode(self, data):
        """Returns bytes, or None when libsnappy rejects the stream."""
        a = _arr(data)
        n = C.c_size_t(0)
        if a.size == 0 or self.l.snappy_uncompressed_length(_ptr(a), a.size, C.byref(n)) != 0:
            return None
        if n.value > (1 << 31):
            return None
        out = np.empty(max(n.value, 1), dtype=np.uint8)
        if self.l.snappy_uncompress(_ptr(a), a.size, _ptr(out), C.byref(n)) != 0:
            return None
        return out[: n.value].tobytes()


def ref_snappycpp():
    l = _load(os.path.join(HERE, "_ref", "libref_snappycpp.so"))
    return None if l is None else RefSnappyCpp(l)


# ---------------------------------------------------------------- faststreams entry points (snappy/faststreams.nim)
# Restated on top of the per-chunk oracle functions above.  Errors are reported by NAME (the reference raises
# exception types, snappy/exceptions.nim:3-12); the bytes appended to the output stream before the error are
# returned with it, because the reference writes chunk by chunk (faststreams.nim:13-18, :116-132).
MAX_UNCOMPRESSED_FRAME = 65536          # codec.nim:14-19
MAX_COMPRESSED_FRAME = 76490            # codec.nim:216-219
FRAMING_HEADER = bytes([0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59])  # codec.nim:33-34


def fs_compress(data):
    """faststreams.nim:20-51 `compress(input, output)`: Leb128(len) then encodeBlock per 64 KiB.
    -> (bytes, None) or (b"", "InputTooLarge")."""
    a = _arr(data)
    if a.size > 0xFFFFFFFF:               # :30-33 checkInputLen / maxCompressedLen
        return b"", "InputTooLarge"
    out = bytearray()
    v = a.size
    while v >= 0x80:                      # :37 toBytes(Leb128)
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    for i in range(0, a.size, MAX_UNCOMPRESSED_FRAME):   # :44-51 (full blocks, then the remainder)
        out += encode_block(a[i:i + MAX_UNCOMPRESSED_FRAME])
    return bytes(out), None


def fs_compress_framed(data):
    """faststreams.nim:64-83 `compressFramed(input, output)`: framingHeader then encodeFrame per 64 KiB."""
    a = _arr(data)
    out = bytearray(FRAMING_HEADER)
    tmp = np.empty(MAX_COMPRESSED_FRAME + 64, dtype=np.uint8)
    for i in range(0, a.size, MAX_UNCOMPRESSED_FRAME):
        blk = np.ascontiguousarray(a[i:i + MAX_UNCOMPRESSED_FRAME])
        n = lib().orc_encode_frame(_ptr(blk), blk.size, _ptr(tmp))
        out += tmp[:n].tobytes()
    return bytes(out)


def fs_uncompress_framed(data, check_integrity=True):
    """faststreams.nim:85-147 `uncompressFramed(input, output, checkIntegrity)`.
    -> (bytes appended to the output, None | "UnexpectedEofError" | "MalformedSnappyData")."""
    b = bytes(data)
    out = bytearray()
    if len(b) < len(FRAMING_HEADER):                      # :88-89
        return bytes(out), "UnexpectedEofError"
    if b[:10] != FRAMING_HEADER:                          # :91-92
        return bytes(out), "MalformedSnappyData"
    pos = 10
    while len(b) - pos >= 4:                              # :95
        cid = b[pos]
        data_len = b[pos + 1] | (b[pos + 2] << 8) | (b[pos + 3] << 16)   # codec.nim:166-172
        pos += 4
        if data_len > MAX_COMPRESSED_FRAME:               # :98-99
            return bytes(out), "MalformedSnappyData"
        if len(b) - pos < data_len:                       # :101-102
            return bytes(out), "UnexpectedEofError"
        if cid == 0x00:                                   # :104-117
            if data_len < 4:
                return bytes(out), "MalformedSnappyData"
            crc = int.from_bytes(b[pos:pos + 4], "little")
            st, dec = uncompress(b[pos + 4:pos + data_len], MAX_UNCOMPRESSED_FRAME)
            if st != OK:
                return bytes(out), "MalformedSnappyData"
            if check_integrity and masked_crc32c(dec) != crc:
                return bytes(out), "MalformedSnappyData"
            out += dec
        elif cid == 0x01:                                 # :119-129
            if data_len < 4:
                return bytes(out), "MalformedSnappyData"
            if data_len - 4 > MAX_UNCOMPRESSED_FRAME:
                return bytes(out), "MalformedSnappyData"
            crc = int.from_bytes(b[pos:pos + 4], "little")
            payload = b[pos + 4:pos + data_len]
            if check_integrity and masked_crc32c(payload) != crc:
                return bytes(out), "MalformedSnappyData"
            out += payload
        elif cid < 0x80:                                  # :131-135
            return bytes(out), "MalformedSnappyData"
        # else: skippable, including a repeated stream identifier (:137-140)
        pos += data_len
    if len(b) - pos >= 1:                                 # :142-143
        return bytes(out), "MalformedSnappyData"
    return bytes(out), None


def st_compress(reads, input_len):
    """streams.nim:9-41 `compress(input: Stream, inputLen, output: Stream)`: Leb128(inputLen), then one encodeBlock
    per readData result (`reads`: the byte strings the stream hands out for successive 64 KiB requests) until
    inputLen bytes have been seen.  -> (bytes, None) or (b"", "InputTooLarge")."""
    if input_len < 0 or input_len > 0xFFFFFFFF:
        return b"", "InputTooLarge"
    out = bytearray()
    v = input_len
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    read = 0
    for r in reads:
        if read >= input_len or not r:
            break
        out += encode_block(_arr(r))
        read += len(r)
    return bytes(out), None
