"""Host-side mirror of the reference's stream entry points (snappy/faststreams.nim) over the C ABI.

  compress(input, output)                          faststreams.nim:20-51   raw format, length known up front
  compressFramed(input, output)                    faststreams.nim:64-83
  uncompressFramed(input, output, checkIntegrity)  faststreams.nim:85-147

`input` is a bytes-like object or a binary file object (`read` / `readinto`), `output` anything with
`write` (and optionally `flush`) -- the roles of faststreams' InputStream / OutputStream.  Errors are the
exception types of snappy/exceptions.nim:3-12.  As in the reference, output is appended chunk by chunk, so
everything decoded before a malformed chunk has been written when the exception is raised.

The reference keeps one 64 KiB `tmp` block per call and encodes a block at a time; here a call streams
slabs of 1024 chunks (64 MiB) through page-locked staging buffers (snb_host_alloc): a reader thread fills /
walks the next slab while the GPU works on the current one.  Every byte of codec work is done by the CUDA
library; this module only moves bytes and walks frame headers.
"""
import ctypes as C
import os
import threading

import numpy as np

from . import _lib
from .codec import (framingHeader, maxCompressedFrameDataLen, maxCompressedLen, maxCompressedLenFramed,
                    maxUncompressedFrameDataLen, maxUncompressedLen)

SLAB_CHUNKS = 1024
SLAB_BYTES = SLAB_CHUNKS * maxUncompressedFrameDataLen


# ---------------------------------------------------------------- snappy/exceptions.nim:3-12
class SnappyError(Exception):
    pass


class SnappyDecodingError(SnappyError):
    pass


class SnappyEncodingError(SnappyError):
    pass


class UnexpectedEofError(SnappyDecodingError):
    pass


class MalformedSnappyData(SnappyDecodingError):
    pass


class InputTooLarge(SnappyEncodingError):
    def __init__(self, msg="Input too large to be compressed with Snappy"):   # exceptions.nim:14-15
        super().__init__(msg)


# ---------------------------------------------------------------- plumbing
class _Pinned:
    """A page-locked byte buffer owned by the library (snb_host_alloc), viewed as a numpy array."""

    def __init__(self, lib, n):
        self.lib, self.n = lib, n
        self.ptr = lib.snb_host_alloc(n)
        if not self.ptr:
            raise _lib.SnbRuntimeError(f"snb_host_alloc({n}) failed: {lib.last_error()}")
        self.arr = np.ctypeslib.as_array((C.c_uint8 * n).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.arr = None
            self.lib.snb_host_free(self.ptr)
            self.ptr = None


class _Source:
    """InputStream stand-in over a bytes-like object or a binary file object."""

    def __init__(self, obj):
        self.f = None
        self.mv = None
        self.pos = 0
        if hasattr(obj, "read"):
            self.f = obj
        else:
            self.mv = memoryview(obj).cast("B")

    def known_len(self):
        """`input.len` (faststreams.nim:27): None when the object cannot tell."""
        if self.mv is not None:
            return len(self.mv) - self.pos
        try:
            if self.f.seekable():
                cur = self.f.tell()
                end = self.f.seek(0, os.SEEK_END)
                self.f.seek(cur)
                return end - cur
        except (OSError, ValueError, AttributeError):
            pass
        return None

    def read_into(self, dst):
        """Fill the uint8 array `dst` as far as the input goes -> bytes read."""
        n = dst.size
        if self.mv is not None:
            k = min(n, len(self.mv) - self.pos)
            if k:
                dst[:k] = np.frombuffer(self.mv[self.pos:self.pos + k], dtype=np.uint8)
            self.pos += k
            return k
        got = 0
        view = memoryview(dst)
        while got < n:
            if hasattr(self.f, "readinto"):
                k = self.f.readinto(view[got:])
            else:
                b = self.f.read(n - got)
                k = len(b)
                view[got:got + k] = b
            if not k:
                break
            got += k
        return got


def _flush(output):
    if hasattr(output, "flush"):
        output.flush()


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


class _Prefetch:
    """Runs `produce(slot)` for slabs 0, 1, 2, ... on a thread, at most one slab ahead of the consumer."""

    def __init__(self, produce):
        self.produce = produce
        self.items = [None, None]
        self.ready = [threading.Semaphore(0), threading.Semaphore(0)]
        self.free = [threading.Semaphore(1), threading.Semaphore(1)]
        self.stop = False
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        k = 0
        while True:
            s = k & 1
            self.free[s].acquire()
            if self.stop:
                return
            try:
                item = self.produce(s)
            except BaseException as e:  # handed to the consumer
                item = e
            self.items[s] = item
            self.ready[s].release()
            if item is None or isinstance(item, BaseException):
                return
            k += 1

    def __iter__(self):
        k = 0
        while True:
            s = k & 1
            self.ready[s].acquire()
            item = self.items[s]
            if isinstance(item, BaseException):
                raise item
            if item is None:
                return
            yield s, item
            self.free[s].release()
            k += 1

    def close(self):
        self.stop = True
        for f in self.free:
            f.release()
        self.t.join(timeout=10)


class FastStreams:
    """The stream entry points bound to one loaded library."""

    def __init__(self, lib=None):
        self._lib = lib
        self._pool = {}
        self._mu = threading.Lock()

    @property
    def lib(self):
        if self._lib is None:
            self._lib = _lib.default()
        return self._lib

    def _buf(self, name, n):
        """Staging buffers are kept between calls (page-locking 64 MiB costs more than compressing it)."""
        b = self._pool.get(name)
        if b is None or b.n < n:
            if b is not None:
                b.free()
            b = self._pool[name] = _Pinned(self.lib, n)
        return b

    def close(self):
        for b in self._pool.values():
            b.free()
        self._pool.clear()

    # ---- compress / compressFramed share the slab loop
    def _encode_slabs(self, src, output, framed, limit=None):
        with self._mu:
            self._encode_slabs_locked(src, output, framed, limit)

    def _encode_slabs_locked(self, src, output, framed, limit):
        lib = self.lib
        hint = limit if limit is not None else src.known_len()
        slab = SLAB_BYTES if hint is None else min(SLAB_BYTES, max(1, -(-hint // 65536)) * 65536)
        first = self._buf("in0", slab), self._buf("in1", slab)
        cap = (maxCompressedLenFramed(slab) if framed else maxCompressedLen(slab)) + 16
        out = self._buf("out", cap)
        left = [limit]

        def produce(slot):
            want = slab if left[0] is None else min(slab, left[0])
            if want == 0:
                return None
            n = src.read_into(first[slot].arr[:want])
            if left[0] is not None:
                left[0] -= n
            return n if n else None

        pf = _Prefetch(produce)
        try:
            w = C.c_size_t(0)
            for slot, n in pf:
                if framed:
                    st = lib.check(lib.snb_compress_framed(first[slot].ptr, n, out.ptr, cap, C.byref(w)))
                    skip = len(framingHeader)        # the stream identifier is written once, by the caller
                else:
                    st = lib.check(lib.snb_compress(first[slot].ptr, n, out.ptr, cap, C.byref(w)))
                    skip = len(_varint(n))           # the slab's own length prefix; the stream's came first
                if st != 0:
                    raise _lib.SnbRuntimeError(f"slab encode failed with status {st}")
                output.write(memoryview(out.arr)[skip:w.value])
        finally:
            pf.close()

    def compress(self, input, output):
        """faststreams.nim:20-51: Leb128(len) then the blocks; raises InputTooLarge beyond 2^32-1 bytes."""
        src = _Source(input)
        total = src.known_len()
        assert total is not None, "TODO: support actual .. streams"      # :27 doAssert input.len.isSome
        if total > maxUncompressedLen:                                    # :28-33
            raise InputTooLarge()
        output.write(_varint(total))                                      # :37
        self._encode_slabs(src, output, framed=False, limit=total)       # :44-51
        _flush(output)                                                    # :53

    def compressFramed(self, input, output):
        """faststreams.nim:64-83."""
        src = _Source(input)
        output.write(framingHeader)                                       # :67
        self._encode_slabs(src, output, framed=True)                     # :72-81
        _flush(output)                                                    # :83

    # ---- uncompressFramed
    def uncompressFramed(self, input, output, checkIntegrity=True):
        """faststreams.nim:85-147.  Raises UnexpectedEofError / MalformedSnappyData; chunks decoded before the
        offending one have been written to `output`."""
        with self._mu:
            self._uncompress_framed_locked(input, output, checkIntegrity)

    def _uncompress_framed_locked(self, input, output, checkIntegrity):
        lib = self.lib
        win = _Window(_Source(input))
        if not win.ensure(len(framingHeader)):                            # :88-89
            raise UnexpectedEofError("Failed to read stream header")
        if bytes(win.view(0, len(framingHeader))) != framingHeader:       # :91-92
            raise MalformedSnappyData("Invalid header value")
        win.consume(len(framingHeader))

        # bytes still to come = what the window already holds + what the source can still give
        left = None if win.eof else win.src.known_len()
        hint = win.avail() if win.eof else (None if left is None else left + win.avail())
        chunks = SLAB_CHUNKS if hint is None else max(1, min(SLAB_CHUNKS, hint // 8 + 1))   # a data frame is >= 8 bytes
        one_frame = maxCompressedFrameDataLen + 4
        stage_cap = SLAB_CHUNKS * one_frame if hint is None else max(one_frame, min(SLAB_CHUNKS * one_frame, hint)) + 16
        stage = self._buf("in0", stage_cap), self._buf("in1", stage_cap)
        out_cap = chunks * maxUncompressedFrameDataLen
        out = self._buf("out", out_cap)
        done = [False]

        def produce(slot):
            if done[0]:
                return None
            n, err = _walk_slab(win, stage[slot].arr, chunks)
            if err is not None or n == 0:
                done[0] = True
            if n == 0 and err is None:
                return None
            return n, err

        pf = _Prefetch(produce)
        try:
            r, w = C.c_size_t(0), C.c_size_t(0)
            for slot, (n, err) in pf:
                if n:
                    st = lib.check(lib.snb_uncompress_framed(stage[slot].ptr, n, out.ptr, out_cap, 0,
                                                             int(checkIntegrity), C.byref(r), C.byref(w)))
                    if w.value:
                        output.write(memoryview(out.arr)[:w.value])       # chunks before a failing one included
                    if st == 3:                                           # :116-117, :128-129
                        raise MalformedSnappyData("Content CRC checksum failed")
                    if st != 0:                                           # :111-112
                        raise MalformedSnappyData("Failed to decompress content")
                if err is not None:
                    raise err
        finally:
            pf.close()
        _flush(output)                                                    # :147


class _Window:
    """Sliding read window over a _Source: frame headers are walked in place, bytes are copied once (to staging)."""

    PIECE = 32 << 20

    def __init__(self, src):
        self.src = src
        if src.mv is not None:                 # in-memory input: the window is the input itself
            self.buf = np.frombuffer(src.mv, dtype=np.uint8)
            self.lo, self.hi, self.eof = src.pos, len(src.mv), True
        else:
            self.buf = np.empty(self.PIECE + 2 * (maxCompressedFrameDataLen + 4), dtype=np.uint8)
            self.lo = self.hi = 0
            self.eof = False

    def avail(self):
        return self.hi - self.lo

    def ensure(self, n):
        """At least n unconsumed bytes in the window, reading more if needed -> bool."""
        while self.hi - self.lo < n and not self.eof:
            if self.lo and self.buf.size - self.hi < self.PIECE:     # compact before refilling
                k = self.hi - self.lo
                self.buf[:k] = self.buf[self.lo:self.hi].copy()
                self.lo, self.hi = 0, k
            room = min(self.buf.size - self.hi, self.PIECE)
            assert room > 0, "window too small for the pending run"   # _walk_slab bounds its run
            got = self.src.read_into(self.buf[self.hi:self.hi + room])
            if got < room:
                self.eof = True
            self.hi += got
        return self.hi - self.lo >= n

    def view(self, off, n):
        return self.buf[self.lo + off:self.lo + off + n]

    def consume(self, n):
        self.lo += n


def _walk_slab(win, stage, max_frames=SLAB_CHUNKS):
    """Walk frame headers (faststreams.nim:95-143) and copy up to SLAB_CHUNKS data frames (header + payload)
    back to back into `stage`.  -> (staged bytes, exception or None).  An exception belongs AFTER the staged
    frames: they precede the offending frame in the stream."""
    staged = 0
    frames = 0
    run = 0          # bytes of consecutive data frames at the window front, not yet copied
    err = None

    def copy_run():
        nonlocal staged, run
        if run:
            stage[staged:staged + run] = win.view(0, run)
            staged += run
            win.consume(run)
            run = 0

    while frames < max_frames:
        if run >= (8 << 20):                                               # keep the window's pending run small
            copy_run()
        if not win.ensure(run + 4):                                        # :95 `while input.readable(4)`
            copy_run()
            if win.avail() >= 1:                                           # :142-143
                err = MalformedSnappyData("Input contains unknown trailing bytes")
            break
        h = win.view(run, 4)
        cid = int(h[0])
        data_len = int(h[1]) | (int(h[2]) << 8) | (int(h[3]) << 16)        # codec.nim:166-172
        if data_len > maxCompressedFrameDataLen:                           # :98-99
            err = MalformedSnappyData(f"Invalid frame length: {data_len}")
        elif not win.ensure(run + 4 + data_len):                           # :101-102
            err = UnexpectedEofError("Failed to read the entire snappy frame")
        elif cid in (0x00, 0x01):
            if data_len < 4:                                               # :105-106, :120-121
                err = MalformedSnappyData("Frame size too low to contain CRC checksum")
            elif cid == 0x01 and data_len - 4 > maxUncompressedFrameDataLen:   # :123-124
                err = MalformedSnappyData(f"Invalid frame length: {data_len}")
            elif staged + run + 4 + data_len > stage.size and (staged or run):
                break                                                      # the slab is full: this frame opens the next one
            else:
                run += 4 + data_len
                frames += 1
                continue
        elif cid < 0x80:                                                   # :131-135
            err = MalformedSnappyData("Invalid chunk type %02x" % cid)
        else:                                                              # :137-140 skippable: dropped here
            copy_run()
            win.consume(4 + data_len)
            continue
        break
    copy_run()
    return staged, err


_api = FastStreams()
compress = _api.compress
compressFramed = _api.compressFramed
uncompressFramed = _api.uncompressFramed
