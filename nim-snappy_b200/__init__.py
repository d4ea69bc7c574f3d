"""B200-native Snappy codec: the framed / raw hot path of status-im/nim-snappy as sm_100a CUDA
kernels behind a C ABI (include/snb.h), plus this thin host mirror of the reference API."""
from . import codec  # noqa: F401
from .codec import (CodecError, FrameError, framingHeader, maxCompressedLen,  # noqa: F401
                    maxCompressedLenFramed)
from .snappy import (Snappy, compress, compressFramed, decode, decodeFramed, encode,  # noqa: F401
                     encodeFramed, isSnappyFramedStream, maskedCrc, uncompress, uncompressedLen,
                     uncompressedLenFramed, uncompressFramed)
from . import faststreams  # noqa: F401,E402  (stream entry points, snappy/faststreams.nim)
