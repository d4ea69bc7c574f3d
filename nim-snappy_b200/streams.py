"""Host-side mirror of the reference's std/streams adapter (snappy/streams.nim:9-41) over the C ABI.

  compress(input, inputLen, output)     streams.nim:9-41   raw format, `inputLen` announced up front

The reference reads `maxBlockLen` bytes at a time with `readData` and encodes EVERY READ as one block
(streams.nim:31-41): a short read ends a block early, and reading stops once `inputLen` bytes have been seen or
the stream is exhausted -- without an error, although the doc comment promises one (streams.nim:12-13).  That
behaviour is kept.  Full-size reads are gathered into slabs so the GPU sees up to 1024 blocks per call; the
reference leaves compressFramed / uncompressFramed of this adapter as TODOs (streams.nim:43-44), so does this.
"""
import numpy as np

from .codec import maxUncompressedLen
from .faststreams import InputTooLarge, SLAB_BYTES, _varint  # noqa: F401  (same exception types, exceptions.nim)
from .snappy import Snappy

maxBlockLen = 65536      # codec.nim:14


class Streams:
    def __init__(self, lib=None):
        self._api = Snappy(lib)

    def _blocks(self, data):
        """The concatenated encodeBlock output of `data` cut at 64 KiB (the body of snappy.nim:56-62)."""
        enc = self._api.encode(data)
        if not enc:
            raise RuntimeError("snb_compress failed")
        return memoryview(enc)[len(_varint(len(data))):]

    def compress(self, input, inputLen, output):
        if inputLen < 0 or inputLen > maxUncompressedLen:              # :19-20 checkInputLen
            raise InputTooLarge()
        output.write(_varint(inputLen))                                # :21-23
        read = 0
        batch = bytearray()
        while read < inputLen:                                         # :31
            b = input.read(maxBlockLen)                                # :33 readData(addr tmpIn[0], tmpIn.len)
            if not b:                                                  # :34-35
                break
            read += len(b)                                             # :41
            if len(b) == maxBlockLen:
                batch += b
                if len(batch) >= SLAB_BYTES:
                    output.write(self._blocks(np.frombuffer(batch, dtype=np.uint8)))
                    batch = bytearray()
            else:                                                      # a short read is a block of its own (:37-38)
                if batch:
                    output.write(self._blocks(np.frombuffer(batch, dtype=np.uint8)))
                    batch = bytearray()
                output.write(self._blocks(np.frombuffer(b, dtype=np.uint8)))
        if batch:
            output.write(self._blocks(np.frombuffer(batch, dtype=np.uint8)))


_api = Streams()
compress = _api.compress
