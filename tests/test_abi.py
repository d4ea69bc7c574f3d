"""The C-ABI shared library loads and exports every symbol include/snb.h declares (no compute calls:
this runs on the CPU-only box, where nvcc cross-compiles libsnb200.so)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "snb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ["snb_compress", "snb_uncompress", "snb_compress_framed", "snb_uncompress_framed", "snb_masked_crc32c",
              "snb_dev_compress_framed", "snb_dev_uncompress_chunks", "snb_dev_compress_batch",
              "snb_dev_uncompress_batch", "snb_frame_index"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from nim_snappy_b200 import _lib
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(dll, s), f"libsnb200.so does not export {s}"
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_pure_host_helpers_match_reference_bounds():
    """codec.nim arithmetic needs no device: the library and the Python mirror agree with the oracle."""
    from nim_snappy_b200 import _lib, codec
    from oracle import oracle as o
    lib = _lib.SnbLibrary(_lib.LIB_PATH)
    for n in [0, 1, 5, 16, 17, 65535, 65536, 65537, 1 << 20, (1 << 32) - 1]:
        assert lib.snb_max_compressed_len(n) == o.max_compressed_len(n) == codec.maxCompressedLen(n)
        assert lib.snb_max_compressed_len_framed(n) == o.max_compressed_len_framed(n) == codec.maxCompressedLenFramed(n)
    assert codec.maxCompressedLen(1 << 32) is None
    assert codec.maxCompressedBlockLen == 76490


def test_product_has_no_cpu_fallback():
    """Without a CUDA device the data path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        return
    from nim_snappy_b200 import _lib
    lib = _lib.SnbLibrary(_lib.LIB_PATH)
    out = (ctypes.c_uint8 * 64)()
    w = ctypes.c_size_t(0)
    src = (ctypes.c_uint8 * 8)(*range(8))
    st = lib.snb_compress(src, 8, out, 64, ctypes.byref(w))
    assert st == _lib.CUDA_ERROR and lib.last_error()


def test_product_never_imports_the_oracle_or_the_emulator():
    pkg = os.path.join(ROOT, "nim-snappy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "libsnb_emu", "snappy_oracle"):
                    assert needle not in text, (f, needle)
