"""Constants, error enums and size bounds of snappy/codec.nim, mirrored for the host API."""
import enum

maxUncompressedLen = 0xFFFFFFFF          # codec.nim:10
maxBlockLen = 65536                      # codec.nim:14
maxUncompressedFrameDataLen = 65536      # codec.nim:18
inputMargin = 15                         # codec.nim:26
minNonLiteralBlockSize = 17              # codec.nim:53
chunkCompressed, chunkUncompressed, chunkStream = 0x00, 0x01, 0xFF   # codec.nim:29-31
framingHeader = bytes([0xFF, 0x06, 0x00, 0x00, 0x73, 0x4E, 0x61, 0x50, 0x70, 0x59])  # codec.nim:33-34
maxCompressedBlockLen = 32 + 65536 + 65536 // 6        # codec.nim:216-219 (76490)
maxCompressedFrameDataLen = maxCompressedBlockLen


class CodecError(enum.IntEnum):          # codec.nim:56-58 (value = C status code = 1 + ord)
    bufferTooSmall = 1
    invalidInput = 2


class FrameError(enum.IntEnum):          # codec.nim:60-64
    bufferTooSmall = 1
    invalidInput = 2
    crcMismatch = 3
    unknownChunk = 4


def maxCompressedLen(n):
    """codec.nim:92-127: None (Opt.none) beyond 2^32-1."""
    if n < 0 or n > maxUncompressedLen:
        return None
    return 32 + n + n // 6


def maxCompressedLenFramed(n):
    """codec.nim:140-164."""
    if n <= 0:
        return len(framingHeader)
    frames = (n + maxUncompressedFrameDataLen - 1) // maxUncompressedFrameDataLen
    return (frames - 1) * (maxUncompressedFrameDataLen + 8) + maxCompressedBlockLen + 8 + len(framingHeader)
