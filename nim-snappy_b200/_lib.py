"""ctypes binding of libsnb200.so (the C ABI of include/snb.h).

The product has exactly one compute backend: the CUDA library built from csrc/.  `default()`
raises if the library is missing or the process has no CUDA device -- there is no CPU path.
(tests/ may construct `SnbLibrary(path)` over the SIMT-emulator build of the same sources; the
package itself never does.)
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNB_LIB") or os.path.join(HERE, "libsnb200.so")  # SNB_LIB: A/B builds of the same CUDA library (tuning only)

OK, BUFFER_TOO_SMALL, INVALID_INPUT, CRC_MISMATCH, UNKNOWN_CHUNK = 0, 1, 2, 3, 4
CUDA_ERROR, UNSUPPORTED, TIMEOUT = 100, 101, 102


class SnbChunk(C.Structure):
    _fields_ = [("in_off", C.c_uint64), ("out_off", C.c_uint64), ("in_len", C.c_uint32), ("out_cap", C.c_uint32),
                ("crc", C.c_uint32), ("kind", C.c_uint32)]


class SnbWalk(C.Structure):
    _fields_ = [("status", C.c_int32), ("partial", C.c_int32), ("crc_first", C.c_int32), ("reserved", C.c_int32),
                ("read", C.c_uint64), ("written", C.c_uint64), ("big_crc_off", C.c_uint64),
                ("big_crc_len", C.c_uint64), ("big_crc", C.c_uint32), ("reserved2", C.c_uint32)]


class SnbRuntimeError(RuntimeError):
    """CUDA / environment failure reported by the library (status >= 100)."""


# name -> (restype, argtypes); the authoritative list of exported symbols (tests check it against include/snb.h)
_vp, _sz, _u64, _i = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
SIGNATURES = {
    "snb_init": (_i, [_i]),
    "snb_shutdown": (None, []),
    "snb_last_error": (C.c_char_p, []),
    "snb_max_compressed_len": (_u64, [_u64]),
    "snb_max_compressed_len_framed": (_u64, [C.c_int64]),
    "snb_uncompressed_len": (_i, [_vp, _sz, C.POINTER(_u64)]),
    "snb_uncompressed_len_framed": (_i, [_vp, _sz, C.POINTER(_u64)]),
    "snb_is_framed_stream": (_i, [_vp, _sz]),
    "snb_masked_crc32c": (_i, [_vp, _sz, C.POINTER(C.c_uint32)]),
    "snb_compress": (_i, [_vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "snb_uncompress": (_i, [_vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "snb_compress_framed": (_i, [_vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "snb_uncompress_framed": (_i, [_vp, _sz, _vp, _sz, _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "snb_compress_batch": (_i, [_vp, _vp, _vp, _sz, _vp, _u64, _vp, _vp, _vp]),
    "snb_uncompress_batch": (_i, [_vp, _vp, _vp, _sz, _vp, _u64, _vp, _vp, _vp, _vp]),
    "snb_frame_index": (_i, [_vp, _sz, _sz, _i, _i, C.POINTER(SnbChunk), _sz, C.POINTER(_sz), C.POINTER(SnbWalk)]),
    "snb_dev_frame_index": (_i, [_vp, _u64, _u64, _i, _i, _vp, _sz, _vp, _sz, C.POINTER(_sz), C.POINTER(SnbWalk), _vp]),
    "snb_dev_uncompress_framed": (_i, [_vp, _u64, _vp, _u64, _i, _i, _vp, _sz, C.POINTER(_u64), C.POINTER(_u64), _vp]),
    "snb_dev_compress_framed": (_i, [_vp, _u64, _vp, _u64, _i, _vp, _vp, C.POINTER(_u64), _vp]),
    "snb_dev_uncompress_chunks": (_i, [_vp, C.POINTER(SnbChunk), _sz, _vp, _i, _vp, _vp]),
    "snb_dev_compress_batch": (_i, [_vp, _vp, _vp, _sz, C.c_uint32, _vp, _vp, _vp, _vp]),
    "snb_dev_uncompress_batch": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "snb_dev_uncompressed_lens": (_i, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "snb_dev_masked_crc32c_chunks": (_i, [_vp, _u64, _vp, _vp]),
    "snb_host_alloc": (_vp, [_sz]),
    "snb_host_free": (None, [_vp]),
    "snb_set_profiling": (None, [_i]),
    "snb_get_profile": (_i, [C.POINTER(C.c_float), _i]),
    "snb_kernel_launches": (_u64, []),
    "snb_raw_block_decodes": (_u64, []),
    "snb_frame_index_fallbacks": (_u64, []),
    "snb_timeouts_recovered": (_u64, []),
    "snb_debug_kernel_span_ns": (_u64, []),
    "snb_set_compress_warps": (None, [_i]),
    "snb_set_compress_algo": (None, [_i]),
    "snb_set_arrive_timeout_ms": (None, [_i]),
}


class SnbLibrary:
    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise SnbRuntimeError(
                f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
                "This package has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def last_error(self):
        return (self.snb_last_error() or b"").decode()

    def check(self, status):
        """Raise for non-reference failures; return reference status codes untouched."""
        if status >= CUDA_ERROR:
            raise SnbRuntimeError(f"libsnb200 status {status}: {self.last_error()}")
        return status


_default = None


def default():
    """The process-wide CUDA library, bound to cuda:LOCAL_RANK (one process per GPU)."""
    global _default
    if _default is None:
        lib = SnbLibrary(LIB_PATH)
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        st = lib.snb_init(dev)
        if st != OK:
            raise SnbRuntimeError(f"snb_init({dev}) failed: {lib.last_error()} -- a CUDA device is required")
        _default = lib
    return _default
