"""Host-side mirror of the reference's in-memory API (snappy.nim) over the C ABI.

Same names, argument meaning and error behaviour as the Nim procs, so the parity tests read like
tests/test_snappy.nim and tests/test_framed.nim:

  compress / uncompress / compressFramed / uncompressFramed   caller-supplied buffers, return
      `Result`-like values: an int / (read, written) tuple on success, a CodecError / FrameError
      member on failure (snappy.nim:27, :84, :130, :169)
  encode / decode / encodeFramed / decodeFramed               allocate, return b"" on error
      (snappy.nim:66, :112, :157, :269)

All per-chunk work runs in the CUDA library; nothing here compresses or decompresses bytes.
"""
import ctypes as C

import numpy as np

from . import _lib
from .codec import (CodecError, FrameError, framingHeader, maxCompressedLen, maxCompressedLenFramed,
                    maxUncompressedLen)


def _ro(buf):
    """bytes-like -> (uint8 ndarray view, address or None)."""
    a = buf if isinstance(buf, np.ndarray) else np.frombuffer(buf, dtype=np.uint8)
    if a.dtype != np.uint8 or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, (a.ctypes.data if a.size else None)


def _rw(buf):
    a = buf if isinstance(buf, np.ndarray) else np.frombuffer(buf, dtype=np.uint8)
    if a.dtype != np.uint8 or not a.flags.c_contiguous or not a.flags.writeable:
        raise TypeError("output must be a writable contiguous uint8 buffer (bytearray / ndarray)")
    return a, (a.ctypes.data if a.size else None)


class Snappy:
    """The reference API bound to one loaded library (the CUDA one unless a test passes a double)."""

    def __init__(self, lib=None):
        self._lib = lib

    @property
    def lib(self):
        if self._lib is None:
            self._lib = _lib.default()
        return self._lib

    # ---- codec.nim helpers
    def maskedCrc(self, data):
        a, p = _ro(data)
        out = C.c_uint32(0)
        self.lib.check(self.lib.snb_masked_crc32c(p, a.size, C.byref(out)))
        return out.value

    def uncompressedLen(self, data):
        a, p = _ro(data)
        v = C.c_uint64(0)
        return v.value if self.lib.snb_uncompressed_len(p, a.size, C.byref(v)) == 0 else None

    def uncompressedLenFramed(self, data):
        a, p = _ro(data)
        v = C.c_uint64(0)
        return v.value if self.lib.snb_uncompressed_len_framed(p, a.size, C.byref(v)) == 0 else None

    def isSnappyFramedStream(self, data):
        a, p = _ro(data)
        return bool(self.lib.snb_is_framed_stream(p, a.size))

    # ---- raw format
    def compress(self, input, output):
        a, pa = _ro(input)
        o, po = _rw(output)
        w = C.c_size_t(0)
        st = self.lib.check(self.lib.snb_compress(pa, a.size, po, o.size, C.byref(w)))
        return w.value if st == 0 else CodecError(st)

    def encode(self, input):
        a, _ = _ro(input)
        cap = maxCompressedLen(a.size)
        if cap is None:
            return b""
        out = np.empty(cap, dtype=np.uint8)
        r = self.compress(a, out)
        return b"" if isinstance(r, CodecError) else out[:r].tobytes()

    def uncompress(self, input, output):
        a, pa = _ro(input)
        o, po = _rw(output)
        w = C.c_size_t(0)
        st = self.lib.check(self.lib.snb_uncompress(pa, a.size, po, o.size, C.byref(w)))
        return w.value if st == 0 else CodecError(st)

    def decode(self, input, maxSize=maxUncompressedLen):
        n = self.uncompressedLen(input)
        if n is None or n > maxSize:
            return b""
        out = np.empty(max(n, 1), dtype=np.uint8)[:n]
        r = self.uncompress(input, out)
        return b"" if isinstance(r, CodecError) else out[:r].tobytes()

    # ---- framed format
    def compressFramed(self, input, output):
        a, pa = _ro(input)
        o, po = _rw(output)
        w = C.c_size_t(0)
        st = self.lib.check(self.lib.snb_compress_framed(pa, a.size, po, o.size, C.byref(w)))
        return w.value if st == 0 else FrameError(st)

    def encodeFramed(self, input):
        a, _ = _ro(input)
        out = np.empty(maxCompressedLenFramed(a.size), dtype=np.uint8)
        r = self.compressFramed(a, out)
        return b"" if isinstance(r, FrameError) else out[:r].tobytes()

    def uncompressFramed(self, input, output, checkHeader=True, checkIntegrity=True):
        a, pa = _ro(input)
        o, po = _rw(output)
        r, w = C.c_size_t(0), C.c_size_t(0)
        st = self.lib.check(self.lib.snb_uncompress_framed(pa, a.size, po, o.size, int(checkHeader),
                                                           int(checkIntegrity), C.byref(r), C.byref(w)))
        return (r.value, w.value) if st == 0 else FrameError(st)

    def decodeFramed(self, input, maxSize=None, checkIntegrity=True):
        n = self.uncompressedLenFramed(input)
        if n is None or (maxSize is not None and n > maxSize):
            return b""
        out = np.empty(max(n, 1), dtype=np.uint8)[:n]
        r = self.uncompressFramed(input, out, checkIntegrity=checkIntegrity)
        return b"" if isinstance(r, FrameError) else out.tobytes()


    # ---- batches of small messages (BASELINE config 4): `compress` / `uncompress` of every message in ONE library call
    def compressBatch(self, packed, offsets, lengths, output=None):
        """snappy.nim:27-64 for each message packed[offsets[i] : offsets[i] + lengths[i]].  Packed output mode:
        -> (output array, out_len uint32[count], status int32[count]); stream i starts at out_len[:i].sum().
        Raises nothing for reference errors: they are in `status` (1 + ord(CodecError))."""
        a, pa = _ro(packed)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        ln = np.ascontiguousarray(lengths, dtype=np.uint32)
        count = ln.size
        if output is None:
            output = np.empty(int((32 + ln.astype(np.uint64) + ln // 6).sum()), dtype=np.uint8)
        o, po = _rw(output)
        out_len = np.zeros(count, dtype=np.uint32)
        status = np.zeros(count, dtype=np.int32)
        st = self.lib.check(self.lib.snb_compress_batch(pa, off.ctypes.data, ln.ctypes.data, count, po, o.size, None,
                                                         out_len.ctypes.data, status.ctypes.data))
        if st != 0:
            return CodecError(st)
        return o, out_len, status

    def uncompressBatch(self, packed, offsets, lengths, output=None):
        """snappy.nim:84-110 for each raw stream packed[offsets[i] : +lengths[i]], outputs packed in message order
        by their announced lengths -> (output array, out_len uint32[count], status int32[count])."""
        a, pa = _ro(packed)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        ln = np.ascontiguousarray(lengths, dtype=np.uint32)
        count = ln.size
        if output is None:
            total = 0
            for i in range(count):
                n = self.uncompressedLen(a[int(off[i]):int(off[i]) + int(ln[i])])
                total += n if n is not None and n <= 0xFFFFFFFF else 0
            output = np.empty(max(total, 1), dtype=np.uint8)[:total]
        o, po = _rw(output)
        out_len = np.zeros(count, dtype=np.uint32)
        status = np.zeros(count, dtype=np.int32)
        st = self.lib.check(self.lib.snb_uncompress_batch(pa, off.ctypes.data, ln.ctypes.data, count, po, o.size, None,
                                                           None, out_len.ctypes.data, status.ctypes.data))
        if st != 0:
            return CodecError(st)
        return o, out_len, status

    def encodeBatch(self, messages):
        """`encode` (snappy.nim:66-82) of a list of bytes-like messages -> list of bytes."""
        lens = np.array([len(m) for m in messages], dtype=np.uint32)
        offs = np.zeros(len(messages), dtype=np.uint64)
        if len(messages) > 1:
            np.cumsum(lens[:-1], out=offs[1:], dtype=np.uint64)
        packed = np.frombuffer(b"".join(bytes(m) for m in messages), dtype=np.uint8)
        r = self.compressBatch(packed, offs, lens)
        if isinstance(r, CodecError):
            return [b""] * len(messages)
        out, out_len, status = r
        ends = np.cumsum(out_len, dtype=np.uint64)
        return [out[int(e) - int(n):int(e)].tobytes() if st == 0 else b"" for e, n, st in zip(ends, out_len, status)]

    def decodeBatch(self, streams):
        """`decode` (snappy.nim:112-128) of a list of raw streams -> list of bytes (b"" on error)."""
        lens = np.array([len(m) for m in streams], dtype=np.uint32)
        offs = np.zeros(len(streams), dtype=np.uint64)
        if len(streams) > 1:
            np.cumsum(lens[:-1], out=offs[1:], dtype=np.uint64)
        packed = np.frombuffer(b"".join(bytes(m) for m in streams), dtype=np.uint8)
        r = self.uncompressBatch(packed, offs, lens)
        if isinstance(r, CodecError):
            return [b""] * len(streams)
        out, out_len, status = r
        res, pos = [], 0
        for i in range(len(streams)):
            n = self.uncompressedLen(streams[i])
            room = n if n is not None and n <= 0xFFFFFFFF else 0
            res.append(out[pos:pos + int(out_len[i])].tobytes() if status[i] == 0 else b"")
            pos += room
        return res


_api = Snappy()
maskedCrc = _api.maskedCrc
uncompressedLen = _api.uncompressedLen
uncompressedLenFramed = _api.uncompressedLenFramed
isSnappyFramedStream = _api.isSnappyFramedStream
compress = _api.compress
encode = _api.encode
uncompress = _api.uncompress
decode = _api.decode
compressFramed = _api.compressFramed
encodeFramed = _api.encodeFramed
uncompressFramed = _api.uncompressFramed
decodeFramed = _api.decodeFramed
compressBatch = _api.compressBatch
uncompressBatch = _api.uncompressBatch
encodeBatch = _api.encodeBatch
decodeBatch = _api.decodeBatch
