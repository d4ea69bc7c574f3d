/*
 * snb.h -- C ABI of the B200-native Snappy codec (libsnb200.so).
 *
 * This is the drop-in boundary for the framed / raw hot path of status-im/nim-snappy
 * (reference @ 8ab5cf0, paths relative to the reference root).  Every entry point is
 * `extern "C"`, takes plain pointers and sizes, and is what a `{.importc, dynlib.}` binding in
 * snappy.nim would call in place of the loops at snappy.nim:56-62, :146-153 and :199-265
 * (binding shown in INTEGRATION.md).  The precedents for this style of FFI in the reference are
 * `masked_crc32c` (snappy/codec.nim:66-69) and the snappy-c.h import in tests/cpp_snappy.nim:8-11.
 *
 * Status codes: 0 = ok, otherwise 1 + ord(CodecError | FrameError) (snappy/codec.nim:56-64),
 * so Nim maps them back with one `case`.  Codes >= 100 are not reference errors: they report a
 * CUDA / environment failure (there is NO CPU fallback) and snb_last_error() has the text.
 *
 * Buffers: the caller owns every buffer; the library never allocates caller-visible memory.
 * Host entry points accept pageable or pinned host memory (pinned is DMA'd directly).
 * `snb_dev_*` entry points take device pointers and run on the given CUDA stream.  They share the library's work
 * buffers: a call on one stream is ordered behind the previous snb_dev_* call's use of those buffers (an event wait,
 * no host synchronisation), so calls may be issued on different streams; they do not overlap each other on the GPU.
 */
#ifndef SNB_H_
#define SNB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNB_OK 0
#define SNB_BUFFER_TOO_SMALL 1 /* CodecError.bufferTooSmall / FrameError.bufferTooSmall */
#define SNB_INVALID_INPUT 2    /* CodecError.invalidInput  / FrameError.invalidInput  */
#define SNB_CRC_MISMATCH 3     /* FrameError.crcMismatch */
#define SNB_UNKNOWN_CHUNK 4    /* FrameError.unknownChunk */
#define SNB_CUDA_ERROR 100     /* CUDA runtime failure, no device, library not initialised */
#define SNB_UNSUPPORTED 101    /* outside the device path's limits */
#define SNB_TIMEOUT 102        /* the streaming compress kernel gave up waiting for an input piece (copy failure upstream) */

#define SNB_MAX_BLOCK_LEN 65536u            /* codec.nim:14-19 maxBlockLen / maxUncompressedFrameDataLen */
#define SNB_MAX_COMPRESSED_FRAME_LEN 76490u /* codec.nim:216-219 */

/* ---- lifecycle --------------------------------------------------------------------------- */
/* Binds the calling process to one GPU (one process per GPU).  Idempotent. */
int snb_init(int device);
void snb_shutdown(void);
const char *snb_last_error(void);

/* ---- bounds and header-only helpers (pure host arithmetic) -------------------------------- */
/* codec.nim:92-120 maxCompressedLen: 32 + n + n/6. */
uint64_t snb_max_compressed_len(uint64_t n);
/* codec.nim:140-164 maxCompressedLenFramed. */
uint64_t snb_max_compressed_len_framed(int64_t n);
/* codec.nim:129-138 uncompressedLen (u64 varint).  SNB_INVALID_INPUT when absent / malformed. */
int snb_uncompressed_len(const uint8_t *in, size_t n, uint64_t *out);
/* codec.nim:178-214 uncompressedLenFramed.  SNB_INVALID_INPUT stands for Opt.none. */
int snb_uncompressed_len_framed(const uint8_t *in, size_t n, uint64_t *out);
/* codec.nim:174-176 isSnappyFramedStream. */
int snb_is_framed_stream(const uint8_t *in, size_t n);

/* ---- host-buffer API: one call == one reference call -------------------------------------- */
/* codec.nim:66-75 masked_crc32c / maskedCrc, computed on the GPU. */
/* Page-locked host buffers for streaming callers (the faststreams entry points, faststreams.nim:20-153, keep a
 * `tmp` block buffer per call; here the staging slabs are pinned so the copies are plain DMA).  NULL on failure. */
void *snb_host_alloc(size_t n);
void snb_host_free(void *p);

int snb_masked_crc32c(const uint8_t *buf, size_t n, uint32_t *crc);
/* snappy.nim:27-64 compress. */
int snb_compress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written);
/* snappy.nim:84-110 uncompress.  Limits of the device path (SNB_UNSUPPORTED, never a reference error code): input
 * below 4 GiB; a stream longer than one block that is NOT block structured (no encoder of this family emits one)
 * is decoded by a single warp and may announce at most 256 MiB. */
int snb_uncompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written);
/* snappy.nim:130-155 compressFramed. */
int snb_compress_framed(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written);
/* snappy.nim:169-267 uncompressFramed, including the resumable (read, written) contract.  On a per-chunk failure
 * (invalidInput / crcMismatch inside chunk k) *read / *written additionally report the intact prefix: the stream
 * offset of chunk k's header and the bytes produced by chunks 0..k-1 (the reference's Result has no counts then). */
int snb_uncompress_framed(const uint8_t *in, size_t n, uint8_t *out, size_t cap, int check_header,
                          int check_integrity, size_t *read, size_t *written);

/* ---- host-buffer message batches (many small independent messages, e.g. gossip payloads) ------ */
/* `compress` (snappy.nim:27-64) of `count` messages held in host memory: message i is
 * in_base[in_off[i] .. +in_len[i]) (at most 65536 bytes, i.e. one block; SNB_UNSUPPORTED otherwise) and its raw
 * stream (varint + tags) has out_len[i] bytes.
 *   out_off == NULL  packed output: the streams are written back to back in message order starting at out_base
 *                    (stream i begins at the sum of out_len[0..i)); the call returns SNB_BUFFER_TOO_SMALL unless
 *                    out_cap >= sum of snb_max_compressed_len(in_len[i]) (the reference's per-call bound, :44-45).
 *                    Only the packed bytes cross the PCIe link.
 *   out_off != NULL  stream i goes to out_base + out_off[i]; status[i] = SNB_BUFFER_TOO_SMALL when fewer than
 *                    snb_max_compressed_len(in_len[i]) bytes remain between out_off[i] and out_cap.
 * status[i]: 0 or 1 + ord(CodecError) (codec.nim:56-58).  Pinned in / out buffers are DMA'd directly. */
int snb_compress_batch(const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len, size_t count,
                       uint8_t *out_base, uint64_t out_cap, const uint64_t *out_off, uint32_t *out_len,
                       int32_t *status);
/* `uncompress` (snappy.nim:84-110) of `count` raw streams held in host memory.
 *   out_off == NULL  packed output in message order by the lengths the streams announce (codec.nim:129-138);
 *                    SNB_BUFFER_TOO_SMALL (whole call) when their sum exceeds out_cap_total; out_cap is ignored.
 *   out_off != NULL  message i goes to out_base + out_off[i] with room out_cap[i] (required).
 * status[i]: 0 / SNB_BUFFER_TOO_SMALL / SNB_INVALID_INPUT exactly as the reference's Result; out_len[i] = bytes
 * produced (0 on error). */
int snb_uncompress_batch(const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len, size_t count,
                         uint8_t *out_base, uint64_t out_cap_total, const uint64_t *out_off, const uint32_t *out_cap,
                         uint32_t *out_len, int32_t *status);

/* ---- device-resident API (inputs and outputs already in HBM) ------------------------------ */
/* One decodable unit of a framed stream, produced by snb_frame_index (the host-side header
 * walk of snappy.nim:199-265) or by the caller. */
typedef struct snb_chunk {
  uint64_t in_off;  /* payload offset in the stream: varint+tags (kind 0) or raw bytes (kind 1, 2) */
  uint64_t out_off; /* where the chunk's bytes go in the output */
  uint32_t in_len;
  uint32_t out_cap; /* uncompressed length announced by the chunk */
  uint32_t crc;     /* stored masked CRC-32C */
  uint32_t kind;    /* 0 compressed, 1 stored, 2 verify-CRC-only (no output) */
} snb_chunk;

typedef struct snb_walk {
  int32_t status;       /* header-level outcome after the listed chunks: 0, or a FrameError code */
  int32_t partial;      /* 1: stopped early because the next chunk does not fit (snappy.nim:227,254) */
  int32_t crc_first;    /* 1: chunks[n-1] is a verify-only chunk whose CRC outranks `status`/`partial` */
  int32_t reserved;
  uint64_t read;        /* reference `read` if every listed chunk decodes */
  uint64_t written;     /* reference `written` if every listed chunk decodes */
  uint64_t big_crc_off; /* oversize stored chunk (status INVALID_INPUT) whose CRC must be checked first */
  uint64_t big_crc_len; /* 0 when unused */
  uint32_t big_crc;
  uint32_t reserved2;
} snb_walk;

/* Host header walk over a framed stream held in host memory.  Fills at most max_chunks
 * descriptors; returns SNB_BUFFER_TOO_SMALL if more are needed (n_chunks = required). */
int snb_frame_index(const uint8_t *stream, size_t n, size_t out_cap, int check_header, int check_integrity,
                    snb_chunk *chunks, size_t max_chunks, size_t *n_chunks, snb_walk *walk);

/* The header walk of snappy.nim:199-265 for a stream that is RESIDENT IN HBM (e.g. one just produced by
 * snb_dev_compress_framed): the chunk list is found on the device (candidate headers in parallel, successor links,
 * pointer jumping from the first header), so the stream never visits the host.  d_frame_lens / n_frames
 * (optional): the table snb_dev_compress_framed returned for this stream -- then the positions come from a prefix sum
 * and only the headers are validated.  Streams the device path does not model (malformed headers, oversize chunks,
 * a stream that ends inside a chunk) are walked on the host from a copy, so `walk` always has the reference's verdict.
 * d_chunks: device array of max_chunks descriptors (SNB_BUFFER_TOO_SMALL with *n_chunks = required otherwise).
 * Synchronises the stream. */
int snb_dev_frame_index(const uint8_t *d_stream, uint64_t n, uint64_t out_cap, int check_header, int check_integrity,
                        const uint32_t *d_frame_lens, size_t n_frames, snb_chunk *d_chunks, size_t max_chunks,
                        size_t *n_chunks, snb_walk *walk, void *stream);
/* snappy.nim:169-267 uncompressFramed, stream and output both in HBM: device-side header walk + chunk-parallel decode
 * + CRC verify; status, *read and *written exactly as snb_uncompress_framed.  Synchronises the stream. */
int snb_dev_uncompress_framed(const uint8_t *d_stream, uint64_t n, uint8_t *d_out, uint64_t out_cap, int check_header,
                              int check_integrity, const uint32_t *d_frame_lens, size_t n_frames, uint64_t *read,
                              uint64_t *written, void *stream);

/* compressFramed over a device buffer.  d_out needs snb_max_compressed_len_framed(n) bytes.
 * with_header: emit the 10-byte stream identifier first (shard 0 of a sharded stream).
 * d_frame_lens (optional, ceil(n/65536) u32): per-chunk frame sizes, the table shards gather.
 * written (optional, host): synchronises the stream and returns the byte count;
 * d_written (optional, device u64) receives it without a sync. */
int snb_dev_compress_framed(const uint8_t *d_in, uint64_t n, uint8_t *d_out, uint64_t cap, int with_header,
                            uint32_t *d_frame_lens, uint64_t *d_written, uint64_t *written, void *stream);

/* Decode `count` chunks of a device-resident framed stream into d_out.  h_status[i] receives
 * the per-chunk status (0 / INVALID_INPUT / CRC_MISMATCH); synchronises the stream.
 * Every descriptor's (in_off, in_len) must lie inside the one buffer d_in points to: the decoder
 * reads a chunk's stream up to 320 bytes past its end when another descriptor of the call covers
 * those bytes (never past the highest end any descriptor names). */
int snb_dev_uncompress_chunks(const uint8_t *d_in, const snb_chunk *h_chunks, size_t count, uint8_t *d_out,
                              int check_integrity, int32_t *h_status, void *stream);

/* Batched raw-format `compress` of `count` independent messages (each <= 65536 bytes, i.e. one
 * block): message i is d_in[d_in_off[i] .. +d_in_len[i]), its stream (varint + tags) goes to
 * d_out + d_out_off[i] (room snb_max_compressed_len(len)) and its size to d_out_len[i].
 * max_len: upper bound of the message lengths (0 = 65536); selects the hash-table size class. */
int snb_dev_compress_batch(const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len, size_t count,
                           uint32_t max_len, uint8_t *d_out, const uint64_t *d_out_off, uint32_t *d_out_len,
                           void *stream);
/* Batched raw-format `uncompress`.  d_status[i]: 0 / BUFFER_TOO_SMALL / INVALID_INPUT.
 * All messages lie in the one buffer d_in points to (same bounded over-read as above: up to 320 bytes past a
 * message, never past the higher end of the first and the last message of the call). */
int snb_dev_uncompress_batch(const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len, size_t count,
                             uint8_t *d_out, const uint64_t *d_out_off, const uint32_t *d_out_cap,
                             uint32_t *d_out_len, int32_t *d_status, void *stream);
/* uncompressedLen of every message of a raw batch (0xffffffff = malformed varint). */
int snb_dev_uncompressed_lens(const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len, size_t count,
                              uint32_t *d_lens, void *stream);
/* Masked CRC-32C of every 64 KiB chunk of a device buffer (ceil(n/65536) results). */
int snb_dev_masked_crc32c_chunks(const uint8_t *d_in, uint64_t n, uint32_t *d_crc, void *stream);

/* ---- measurement hooks -------------------------------------------------------------------- */
/* When enabled, snb_dev_* calls bracket each kernel with CUDA events on the launching stream. */
void snb_set_profiling(int enabled);
/* Milliseconds of the last profiled call: [0] crc, [1] compress, [2] scan, [3] compact, [4] decompress;
 * [5] = number of kernels launched by the last snb_dev_* call.  Returns entries written. */
int snb_get_profile(float *ms, int n);
/* Total kernels launched by this library since snb_init (for bench.py's gpu_launches). */
uint64_t snb_kernel_launches(void);
/* Raw streams of more than one block that snb_uncompress decoded block-parallel (block starts found on the device)
 * rather than on a single warp -- for tests and tuning. */
uint64_t snb_raw_block_decodes(void);
/* Device-side header walks (snb_dev_frame_index / snb_dev_uncompress_framed) that had to be finished by the host walk. */
uint64_t snb_frame_index_fallbacks(void);
/* Debug (env SNB_DEBUG_SPAN=1): %globaltimer nanoseconds between the first CTA start and the last warp exit of the
 * last streaming compress kernel -- a clock neither CUDA events nor a profiler own.  0 when not enabled. */
uint64_t snb_debug_kernel_span_ns(void);
/* Tuning: warps per CTA of the compress kernel for 64 KiB blocks (hash tables are 32 KiB each). */
void snb_set_compress_warps(int warps);
/* A-B testing of the block encoder for 64 KiB chunks: 0 / 5 = one warp per chunk (default), 6 = resolver warps + walker
 * warp per chunk with every hash table in shared memory (experimental: byte-identical, slower). */
void snb_set_compress_algo(int algo);
/* How long the streaming compress kernel waits for a chunk whose host-to-device copy has not landed (default 2000 ms).
 * A host-buffer call whose kernel gave up is repeated once with every copy enqueued ahead of its kernel (pageable input
 * is normally copied while the kernel already runs; another thread's cudaMalloc / cudaFree can hold that up); only if
 * the repeat times out too does the call return SNB_TIMEOUT instead of hanging. */
void snb_set_arrive_timeout_ms(int ms);
/* Host-buffer compress calls that timed out once and were repeated (see above) -- for tests. */
uint64_t snb_timeouts_recovered(void);

#ifdef __cplusplus
}
#endif
#endif /* SNB_H_ */
