## snappy_b200_batch.nim -- many small messages in ONE library call (BASELINE config 4: gossip-sized payloads).
##
## The reference compresses one message per `compress` / `encode` call (snappy.nim:27-82) and decodes one per
## `uncompress` / `decode` call (:84-128).  On the GPU a single small message is one warp's job and a call costs a
## PCIe round trip, so the binding that makes sense for a gossip node is the batched one: hand over all pending
## messages, get all results, same per-message `Result[..., CodecError]` as the reference procs would have given.
## C ABI: snb_compress_batch / snb_uncompress_batch (include/snb.h), packed-output mode.
## Not compiled in the build image (no Nim toolchain there); nim-snappy_b200/snappy.py (compressBatch, uncompressBatch,
## encodeBatch, decodeBatch) is the same logic in Python and is what tests/ run.
{.push raises: [].}

import results
import snappy/codec
import ./snappy_b200           # SnbUnavailable, snbAvailable
export codec, results

const snbLib = "libsnb200.so"

proc snb_compress_batch(inBase: ptr byte, inOff: ptr uint64, inLen: ptr uint32, count: csize_t,
                        outBase: ptr byte, outCap: uint64, outOff: ptr uint64, outLen: ptr uint32,
                        status: ptr int32): cint {.cdecl, importc, dynlib: snbLib.}
proc snb_uncompress_batch(inBase: ptr byte, inOff: ptr uint64, inLen: ptr uint32, count: csize_t,
                          outBase: ptr byte, outCapTotal: uint64, outOff: ptr uint64, outCap: ptr uint32,
                          outLen: ptr uint32, status: ptr int32): cint {.cdecl, importc, dynlib: snbLib.}
proc snb_last_error(): cstring {.cdecl, importc, dynlib: snbLib.}

func toCodecError(st: int32): CodecError =
  if st == 1: CodecError.bufferTooSmall else: CodecError.invalidInput

proc pack(msgs: openArray[seq[byte]]): tuple[data: seq[byte], off: seq[uint64], len: seq[uint32]] =
  var total = 0
  for m in msgs: total += m.len
  result.data = newSeqUninitialized[byte](max(total, 1))
  result.off = newSeq[uint64](msgs.len)
  result.len = newSeq[uint32](msgs.len)
  var pos = 0
  for i, m in msgs:
    result.off[i] = uint64 pos
    result.len[i] = uint32 m.len
    if m.len > 0: copyMem(addr result.data[pos], unsafeAddr m[0], m.len)
    pos += m.len

proc compressBatch*(msgs: openArray[seq[byte]]): seq[Result[seq[byte], CodecError]] =
  ## `encode` (snappy.nim:66-82) of every message; messages longer than one block (65536 bytes) are not batch
  ## material -- route those through `compress` / `compressFramed`.
  result = newSeq[Result[seq[byte], CodecError]](msgs.len)
  if msgs.len == 0: return
  var (data, off, lens) = pack(msgs)
  var cap = 0'u64
  for m in msgs:
    if m.len > int(maxBlockLen):
      for r in result.mitems: r = err(CodecError.invalidInput)
      return
    cap += 32'u64 + uint64(m.len) + uint64(m.len div 6)       # codec.nim:119 maxCompressedLen
  var
    outp = newSeqUninitialized[byte](int cap)
    outLen = newSeq[uint32](msgs.len)
    status = newSeq[int32](msgs.len)
  let st = snb_compress_batch(addr data[0], addr off[0], addr lens[0], csize_t msgs.len,
                              addr outp[0], cap, nil, addr outLen[0], addr status[0])
  if st >= 100:
    raise (ref SnbUnavailable)(msg: "libsnb200 status " & $st & ": " & $snb_last_error())
  var pos = 0
  for i in 0 ..< msgs.len:
    if st != 0: result[i] = err(toCodecError(int32 st))
    elif status[i] != 0: result[i] = err(toCodecError(status[i]))
    else: result[i] = ok(outp[pos ..< pos + int(outLen[i])])
    pos += int(outLen[i])

proc uncompressBatch*(streams: openArray[seq[byte]], maxSize = maxUncompressedLen): seq[Result[seq[byte], CodecError]] =
  ## `decode` (snappy.nim:112-128) of every raw stream, with the reference's per-message errors.
  result = newSeq[Result[seq[byte], CodecError]](streams.len)
  if streams.len == 0: return
  var (data, off, lens) = pack(streams)
  var total = 0'u64
  var room = newSeq[uint64](streams.len)
  for i, s in streams:
    let n = uncompressedLen(s)                                  # codec.nim:129-138
    room[i] = if n.isSome and n.get <= maxSize and n.get <= 0xffffffff'u64: n.get else: 0'u64
    total += room[i]
  var
    outp = newSeqUninitialized[byte](int max(total, 1'u64))
    outLen = newSeq[uint32](streams.len)
    status = newSeq[int32](streams.len)
  let st = snb_uncompress_batch(addr data[0], addr off[0], addr lens[0], csize_t streams.len,
                                addr outp[0], total, nil, nil, addr outLen[0], addr status[0])
  if st >= 100:
    raise (ref SnbUnavailable)(msg: "libsnb200 status " & $st & ": " & $snb_last_error())
  var pos = 0
  for i in 0 ..< streams.len:
    if st != 0: result[i] = err(toCodecError(int32 st))
    elif status[i] != 0: result[i] = err(toCodecError(status[i]))
    else: result[i] = ok(outp[pos ..< pos + int(outLen[i])])
    pos += int(room[i])
