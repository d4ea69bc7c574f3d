## snappy_b200_faststreams.nim -- reference-side binding of the stream entry points (snappy/faststreams.nim:20-147)
## over libsnb200.so.  Same names, signatures and exception types as the reference; instead of one 64 KiB block per
## loop turn, a call moves slabs of 1024 chunks (64 MiB) through page-locked staging buffers (snb_host_alloc), so
## the GPU sees enough independent chunks per call.  nim-snappy_b200/faststreams.py is the same logic in Python (that
## one is what the tests run; there is no Nim toolchain in the build image).
{.push raises: [].}

import
  pkg/faststreams/[inputs, multisync, outputs],
  stew/[leb128, endians2],
  snappy/[codec, exceptions]

export inputs, multisync, outputs, codec, exceptions

const
  snbLib = "libsnb200.so"
  slabChunks = 1024
  slabBytes = slabChunks * int(maxUncompressedFrameDataLen)
  maxFrame = int(maxCompressedFrameDataLen) + 4          # header + payload of the largest data frame

proc snb_host_alloc(n: csize_t): pointer {.cdecl, importc, dynlib: snbLib.}
proc snb_host_free(p: pointer) {.cdecl, importc, dynlib: snbLib.}
proc snb_compress(inp: ptr byte, n: csize_t, outp: ptr byte, cap: csize_t,
                  written: ptr csize_t): cint {.cdecl, importc, dynlib: snbLib.}
proc snb_compress_framed(inp: ptr byte, n: csize_t, outp: ptr byte, cap: csize_t,
                         written: ptr csize_t): cint {.cdecl, importc, dynlib: snbLib.}
proc snb_uncompress_framed(inp: ptr byte, n: csize_t, outp: ptr byte, cap: csize_t,
                           checkHeader, checkIntegrity: cint,
                           read, written: ptr csize_t): cint {.cdecl, importc, dynlib: snbLib.}

type Pinned = object
  p: ptr UncheckedArray[byte]
  n: int

proc pinned(n: int): Pinned {.raises: [IOError].} =
  let q = snb_host_alloc(csize_t n)
  if q.isNil: raise newException(IOError, "snb_host_alloc failed (no CUDA device?)")
  Pinned(p: cast[ptr UncheckedArray[byte]](q), n: n)

proc release(b: var Pinned) =
  if not b.p.isNil: snb_host_free(b.p); b.p = nil

proc fill(input: InputStream, dst: Pinned, want: int): int {.raises: [IOError].} =
  ## Reads up to `want` bytes (short only at the end of the input).
  while result < want and input.readable:
    let k = min(want - result, input.totalUnconsumedBytes)
    if k == 0: break
    copyMem(addr dst.p[result], unsafeAddr input.read(k)[0], k)
    result += k

proc encodeSlabs(input: InputStream, output: OutputStream, framed: bool) {.raises: [IOError].} =
  var
    src = pinned(slabBytes)
    dst = pinned(int(maxCompressedLenFramed(slabBytes)) + 16)
  defer: release(src); release(dst)
  while true:
    let n = fill(input, src, slabBytes)
    if n == 0: break
    var w: csize_t
    let st =
      if framed: snb_compress_framed(addr src.p[0], csize_t n, addr dst.p[0], csize_t dst.n, addr w)
      else: snb_compress(addr src.p[0], csize_t n, addr dst.p[0], csize_t dst.n, addr w)
    if st != 0: raise newException(IOError, "libsnb200: status " & $st)
    # the stream identifier / the length prefix were written once by the caller: drop the slab's own
    let skip = if framed: framingHeader.len else: uint32(n).toBytes(Leb128).len
    output.write(dst.p.toOpenArray(skip, int(w) - 1))

proc compress*(input: InputStream, output: OutputStream) {.raises: [InputTooLarge, IOError].} =
  ## faststreams.nim:20-53
  doAssert input.len.isSome, "TODO: support actual .. streams"
  let lenU32 = checkInputLen(input.len.get).valueOr: raiseInputTooLarge()
  output.write lenU32.toBytes(Leb128).toOpenArray()
  encodeSlabs(input, output, framed = false)             # slabs are multiples of 64 KiB: block boundaries are kept
  output.flush()

proc compressFramed*(input: InputStream, output: OutputStream) {.raises: [IOError].} =
  ## faststreams.nim:64-83
  output.write(framingHeader)
  encodeSlabs(input, output, framed = true)
  output.flush()

proc uncompressFramed*(input: InputStream, output: OutputStream, checkIntegrity = true) {.
    raises: [IOError, SnappyDecodingError].} =
  ## faststreams.nim:85-147.  Frame headers are walked here (same checks, same exceptions); runs of data frames
  ## go to the GPU 1024 at a time; everything decoded before a bad frame has been written when the exception is raised.
  if not input.readable(framingHeader.len): raise newException(UnexpectedEofError, "Failed to read stream header")
  if input.read(framingHeader.len) != framingHeader: raise newException(MalformedSnappyData, "Invalid header value")
  var
    stage = pinned(slabChunks * maxFrame)
    dst = pinned(slabBytes)
  defer: release(stage); release(dst)
  var staged, frames = 0
  var pending: ref SnappyDecodingError = nil             # raised after the frames staged before it have been written
  proc flushRun() {.raises: [IOError, SnappyDecodingError].} =
    if staged > 0:
      var r, w: csize_t
      let st = snb_uncompress_framed(addr stage.p[0], csize_t staged, addr dst.p[0], csize_t dst.n, 0,
                                     cint(checkIntegrity), addr r, addr w)
      if w > 0: output.write(dst.p.toOpenArray(0, int(w) - 1))       # the intact prefix, also on a per-chunk failure
      if st == 3: raise newException(MalformedSnappyData, "Content CRC checksum failed")
      if st != 0: raise newException(MalformedSnappyData, "Failed to decompress content")
      staged = 0; frames = 0
  while pending.isNil and input.readable(4):
    let
      hdr = input.read(4)
      id = hdr[0]
      dataLen = int(uint32.fromBytesLE([hdr[1], hdr[2], hdr[3], 0'u8]))
    if dataLen > int(maxCompressedFrameDataLen):
      pending = newException(MalformedSnappyData, "Invalid frame length: " & $dataLen)
    elif not input.readable(dataLen):
      pending = newException(UnexpectedEofError, "Failed to read the entire snappy frame")
    elif id in {chunkCompressed, chunkUncompressed}:
      if dataLen < 4:
        pending = newException(MalformedSnappyData, "Frame size too low to contain CRC checksum")
      elif id == chunkUncompressed and dataLen - 4 > int(maxUncompressedFrameDataLen):
        pending = newException(MalformedSnappyData, "Invalid frame length: " & $dataLen)
      else:
        copyMem(addr stage.p[staged], unsafeAddr hdr[0], 4)
        copyMem(addr stage.p[staged + 4], unsafeAddr input.read(dataLen)[0], dataLen)
        staged += 4 + dataLen; inc frames
        if frames == slabChunks: flushRun()
    elif id < 0x80:
      pending = newException(MalformedSnappyData, "Invalid chunk type " & toHex(id))
    else:
      discard input.read(dataLen)                          # skippable, including a repeated stream identifier
  flushRun()
  if not pending.isNil: raise pending
  if input.readable(1): raise newException(MalformedSnappyData, "Input contains unknown trailing bytes")
  output.flush()
