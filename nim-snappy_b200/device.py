"""Device-resident (HBM in, HBM out) entry points over torch uint8 CUDA tensors.

torch is plumbing here: it owns device memory and the current stream; every byte of codec work is
done by the kernels in csrc/ through the C ABI (`snb_dev_*` in include/snb.h).
"""
import ctypes as C

import numpy as np

from . import _lib
from .codec import maxCompressedLen, maxCompressedLenFramed

CHUNK_DTYPE = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("in_len", "<u4"), ("out_cap", "<u4"),
                        ("crc", "<u4"), ("kind", "<u4")])
assert CHUNK_DTYPE.itemsize == C.sizeof(_lib.SnbChunk)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else None


class DeviceCodec:
    def __init__(self, lib=None):
        self.lib = lib or _lib.default()

    # ---- framed
    def compress_framed(self, d_in, out=None, with_header=True, frame_lens=None, sync=True):
        """compressFramed (snappy.nim:130-155) of a device buffer -> (out tensor, bytes written).
        With sync=False the count stays on the device (returned as a 1-element int64 tensor)."""
        import torch
        n = d_in.numel()
        cap = maxCompressedLenFramed(n)
        if out is None:
            out = torch.empty(cap, dtype=torch.uint8, device=d_in.device)
        written = C.c_uint64(0)
        d_written = None if sync else torch.zeros(1, dtype=torch.int64, device=d_in.device)
        st = self.lib.snb_dev_compress_framed(_ptr(d_in), n, _ptr(out), out.numel(), int(with_header),
                                              _ptr(frame_lens), _ptr(d_written),
                                              C.byref(written) if sync else None, _stream())
        if self.lib.check(st) != 0:
            raise ValueError(f"snb_dev_compress_framed: status {st}")
        return out, (written.value if sync else d_written)

    def frame_index(self, stream, out_cap=None, check_header=True, check_integrity=True):
        """Host header walk (snappy.nim:199-265) -> (structured chunk array, SnbWalk)."""
        a = stream if isinstance(stream, np.ndarray) else np.frombuffer(stream, dtype=np.uint8)
        if out_cap is None:
            out_cap = (1 << 62)
        n_chunks = C.c_size_t(0)
        walk = _lib.SnbWalk()
        guess = max(16, a.size // 1024)
        while True:
            chunks = np.zeros(guess, dtype=CHUNK_DTYPE)
            st = self.lib.snb_frame_index(a.ctypes.data if a.size else None, a.size, out_cap, int(check_header),
                                          int(check_integrity), chunks.ctypes.data_as(C.POINTER(_lib.SnbChunk)),
                                          guess, C.byref(n_chunks), C.byref(walk))
            if st == _lib.BUFFER_TOO_SMALL:
                guess = n_chunks.value
                continue
            return chunks[: n_chunks.value], walk

    def dev_frame_index(self, d_stream, out_cap=None, check_header=True, check_integrity=True, frame_lens=None):
        """Header walk of a stream resident in HBM (snb_dev_frame_index) -> (device chunk tensor [n, 32] uint8, SnbWalk)."""
        import torch
        n = d_stream.numel()
        if out_cap is None:
            out_cap = (1 << 62)
        guess = max(16, n // 8192)
        while True:
            d_chunks = torch.empty((guess, C.sizeof(_lib.SnbChunk)), dtype=torch.uint8, device=d_stream.device)
            n_chunks = C.c_size_t(0)
            walk = _lib.SnbWalk()
            st = self.lib.snb_dev_frame_index(_ptr(d_stream), n, out_cap, int(check_header), int(check_integrity),
                                              _ptr(frame_lens), 0 if frame_lens is None else frame_lens.numel(),
                                              _ptr(d_chunks), guess, C.byref(n_chunks), C.byref(walk), _stream())
            if self.lib.check(st) == _lib.BUFFER_TOO_SMALL:
                guess = n_chunks.value
                continue
            return d_chunks[: n_chunks.value], walk

    def uncompress_framed(self, d_stream, d_out, check_header=True, check_integrity=True, frame_lens=None):
        """uncompressFramed (snappy.nim:169-267), stream and output in HBM -> (status, read, written)."""
        rd, wr = C.c_uint64(0), C.c_uint64(0)
        st = self.lib.snb_dev_uncompress_framed(_ptr(d_stream), d_stream.numel(), _ptr(d_out), d_out.numel(),
                                                int(check_header), int(check_integrity), _ptr(frame_lens),
                                                0 if frame_lens is None else frame_lens.numel(),
                                                C.byref(rd), C.byref(wr), _stream())
        return self.lib.check(st), rd.value, wr.value

    def uncompress_chunks(self, d_stream, chunks, d_out, check_integrity=True):
        """Decode walked chunks of a device-resident stream into d_out -> int32 status per chunk."""
        chunks = np.ascontiguousarray(chunks)
        status = np.zeros(max(len(chunks), 1), dtype=np.int32)
        st = self.lib.snb_dev_uncompress_chunks(_ptr(d_stream), chunks.ctypes.data_as(C.POINTER(_lib.SnbChunk)),
                                                len(chunks), _ptr(d_out), int(check_integrity),
                                                status.ctypes.data, _stream())
        self.lib.check(st)
        if st != 0:
            raise ValueError(f"snb_dev_uncompress_chunks: status {st}")
        return status[: len(chunks)]

    def masked_crc_chunks(self, d_in):
        import torch
        n = d_in.numel()
        out = torch.empty((n + 65535) // 65536, dtype=torch.int32, device=d_in.device)
        self.lib.check(self.lib.snb_dev_masked_crc32c_chunks(_ptr(d_in), n, _ptr(out), _stream()))
        return out

    # ---- raw batches (many small messages)
    def compress_batch(self, d_in, d_in_off, d_in_len, max_len, d_out, d_out_off, d_out_len=None):
        """`compress` (snappy.nim:27-64) of every message; offsets int64, lengths int32 device tensors."""
        import torch
        count = d_in_len.numel()
        if d_out_len is None:
            d_out_len = torch.empty(count, dtype=torch.int32, device=d_in.device)
        st = self.lib.snb_dev_compress_batch(_ptr(d_in), _ptr(d_in_off), _ptr(d_in_len), count, int(max_len),
                                             _ptr(d_out), _ptr(d_out_off), _ptr(d_out_len), _stream())
        self.lib.check(st)
        if st != 0:
            raise ValueError(f"snb_dev_compress_batch: status {st}")
        return d_out_len

    def uncompress_batch(self, d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap):
        import torch
        count = d_in_len.numel()
        d_out_len = torch.empty(count, dtype=torch.int32, device=d_in.device)
        d_status = torch.empty(count, dtype=torch.int32, device=d_in.device)
        st = self.lib.snb_dev_uncompress_batch(_ptr(d_in), _ptr(d_in_off), _ptr(d_in_len), count, _ptr(d_out),
                                               _ptr(d_out_off), _ptr(d_out_cap), _ptr(d_out_len), _ptr(d_status),
                                               _stream())
        self.lib.check(st)
        if st != 0:
            raise ValueError(f"snb_dev_uncompress_batch: status {st}")
        return d_out_len, d_status

    def uncompressed_lens(self, d_in, d_in_off, d_in_len):
        import torch
        out = torch.empty(d_in_len.numel(), dtype=torch.int32, device=d_in.device)
        self.lib.check(self.lib.snb_dev_uncompressed_lens(_ptr(d_in), _ptr(d_in_off), _ptr(d_in_len),
                                                          d_in_len.numel(), _ptr(out), _stream()))
        return out

    # ---- measurement hooks
    def set_profiling(self, on):
        self.lib.snb_set_profiling(int(on))

    def profile(self):
        buf = (C.c_float * 6)()
        self.lib.snb_get_profile(buf, 6)
        return dict(zip(("crc_ms", "compress_ms", "scan_ms", "compact_ms", "decompress_ms", "launches"), list(buf)))


def max_batch_out_stride(max_len):
    """Per-message output room for compress_batch: maxCompressedLen rounded up to 16 bytes."""
    return (maxCompressedLen(max_len) + 15) // 16 * 16
