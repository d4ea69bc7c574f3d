## snappy_b200.nim -- reference-side binding of libsnb200.so (include/snb.h).
##
## Drop-in replacements for the in-memory procs of status-im/nim-snappy (snappy.nim): same names,
## same signatures, same `Result` error semantics; the per-chunk loops (snappy.nim:56-62, :146-153,
## :199-265) run on the GPU behind the C ABI.  Not compiled in the build image (no Nim toolchain
## there); kept source-compatible with the reference at 8ab5cf0.
{.push raises: [].}

import results
import snappy/codec            # CodecError, FrameError, maxCompressedLen*, framingHeader ...
export codec, results

const snbLib = "libsnb200.so"

proc snb_init(device: cint): cint {.cdecl, importc, dynlib: snbLib.}
proc snb_masked_crc32c(buf: ptr byte, n: csize_t, crc: ptr uint32): cint {.cdecl, importc, dynlib: snbLib.}
proc snb_compress(inp: ptr byte, n: csize_t, outp: ptr byte, cap: csize_t,
                  written: ptr csize_t): cint {.cdecl, importc, dynlib: snbLib.}
proc snb_uncompress(inp: ptr byte, n: csize_t, outp: ptr byte, cap: csize_t,
                    written: ptr csize_t): cint {.cdecl, importc, dynlib: snbLib.}
proc snb_compress_framed(inp: ptr byte, n: csize_t, outp: ptr byte, cap: csize_t,
                         written: ptr csize_t): cint {.cdecl, importc, dynlib: snbLib.}
proc snb_uncompress_framed(inp: ptr byte, n: csize_t, outp: ptr byte, cap: csize_t,
                           checkHeader, checkIntegrity: cint,
                           read, written: ptr csize_t): cint {.cdecl, importc, dynlib: snbLib.}

template p(a: openArray[byte]): ptr byte =
  if a.len == 0: nil else: unsafeAddr a[0]

proc snb_last_error(): cstring {.cdecl, importc, dynlib: snbLib.}

type SnbUnavailable* = object of Defect
  ## Status >= 100: no CUDA device / CUDA runtime failure / outside the device path's limits.  This is NOT a property
  ## of the input, so it must never surface as `invalidInput` ("malformed data"): the library has no CPU path, and a
  ## host that may lack a GPU asks `snbAvailable()` once at start-up and keeps the original procs otherwise.

proc snbAvailable*(device = 0): bool =
  ## True when libsnb200 can run on `device` (binds this process to it: one process per GPU).
  snb_init(cint device) == 0

proc snbLastError*(): string = $snb_last_error()

template checkEnv(st: cint) =
  if st >= 100:
    raise (ref SnbUnavailable)(msg: "libsnb200 status " & $st & ": " & $snb_last_error())

func toCodecError(st: cint): CodecError =
  # status = 1 + ord(enum) (codec.nim:56-58); environment failures were raised by checkEnv before this point
  if st == 1: CodecError.bufferTooSmall else: CodecError.invalidInput

func toFrameError(st: cint): FrameError =
  case st
  of 1: FrameError.bufferTooSmall
  of 3: FrameError.crcMismatch
  of 4: FrameError.unknownChunk
  else: FrameError.invalidInput

proc compress*(input: openArray[byte], output: var openArray[byte]): Result[int, CodecError] =
  var written: csize_t
  let st = snb_compress(p(input), input.len.csize_t, p(output), output.len.csize_t, addr written)
  checkEnv st
  if st == 0: ok(int written) else: err(toCodecError(st))

proc uncompress*(input: openArray[byte], output: var openArray[byte]): Result[int, CodecError] =
  var written: csize_t
  let st = snb_uncompress(p(input), input.len.csize_t, p(output), output.len.csize_t, addr written)
  checkEnv st
  if st == 0: ok(int written) else: err(toCodecError(st))

proc compressFramed*(input: openArray[byte], output: var openArray[byte]): Result[int, FrameError] =
  var written: csize_t
  let st = snb_compress_framed(p(input), input.len.csize_t, p(output), output.len.csize_t, addr written)
  checkEnv st
  if st == 0: ok(int written) else: err(toFrameError(st))

proc uncompressFramed*(input: openArray[byte], output: var openArray[byte],
                       checkHeader = true, checkIntegrity = true):
                       Result[tuple[read: int, written: int], FrameError] =
  var read, written: csize_t
  let st = snb_uncompress_framed(p(input), input.len.csize_t, p(output), output.len.csize_t,
                                 cint(checkHeader), cint(checkIntegrity), addr read, addr written)
  checkEnv st
  if st == 0: ok((int read, int written)) else: err(toFrameError(st))

proc encode*(input: openArray[byte]): seq[byte] =
  let maxCompressed = maxCompressedLen(input.len).valueOr: return
  result = newSeqUninitialized[byte](maxCompressed)
  let written = compress(input, result).valueOr:
    return @[]
  result.setLen(written)

proc decode*(input: openArray[byte], maxSize = maxUncompressedLen): seq[byte] =
  let uncompressed = uncompressedLen(input).valueOr: return
  if uncompressed > maxSize or uncompressed > int.high.uint64: return
  result = newSeqUninitialized[byte](int uncompressed)
  if uncompress(input, result).isErr(): result = @[]

proc encodeFramed*(input: openArray[byte]): seq[byte] =
  let maxCompressed = maxCompressedLenFramed(input.len)
  if maxCompressed > int.high.uint64: return
  result = newSeqUninitialized[byte](int maxCompressed)
  let written = compressFramed(input, result).valueOr:
    return @[]
  result.setLen(written)

proc decodeFramed*(input: openArray[byte], maxSize = int.high, checkIntegrity = true): seq[byte] =
  let uncompressed = uncompressedLenFramed(input).valueOr: return
  if uncompressed > maxSize.uint64: return
  result = newSeqUninitialized[byte](int uncompressed)
  if uncompressFramed(input, result, checkIntegrity = checkIntegrity).isErr(): result = @[]

proc maskedCrc*(data: openArray[byte]): uint32 =
  ## codec.nim:71-75 on the GPU (only worthwhile for large buffers; codec.nim's C version stays for small ones)
  discard snb_masked_crc32c(p(data), data.len.csize_t, addr result)
