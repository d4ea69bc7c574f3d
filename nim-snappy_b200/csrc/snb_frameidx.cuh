// snb_frameidx.cuh -- the framed header walk (snappy.nim:199-265, codec.nim:166-172) for a stream that is resident
// in HBM (included by snb_device.cuh).
//
// The reference walks the chunk headers as a linked list: header -> 24-bit length -> next header.  Done serially on
// the device that is one dependent DRAM load per 64 KiB chunk (~16 ms for the 1 GiB corpus, longer than decoding it);
// done on the host it costs a D2H of the whole stream.  Here the list is found in parallel:
//
//   scan      every stream position is tested for "could be a chunk header": a type the walk accepts (0x00, 0x01,
//             >= 0x80) whose length fits the stream and, for data chunks, passes the reference's own size / varint
//             checks (skippable chunks: up to 64 KiB), and that leads to three more such headers or to the end of
//             the stream.  Candidates are kept in stream order.
//   link      a candidate's successor is the candidate at position + 4 + length (binary search), the end of the
//             stream, or nothing.
//   reach     starting from the first header the reachable set doubles per round (pointer jumping): after
//             ceil(log2(candidates)) rounds exactly the headers the serial walk would visit are marked, in order.
//   describe  one thread per reached header fills its chunk descriptor; a prefix sum of the announced lengths gives
//             the output offsets and the first chunk that does not fit the caller's room (snappy.nim:219-229, :253-254).
//
// Anything the fast path does not model -- a header the walk would reject, a stream that ends inside a chunk, too
// many candidates -- is reported, and the host then runs walk_framed on a copy: statuses and the (read, written)
// contract are always the reference's.
#pragma once

namespace snb {
namespace fidx {

#ifndef SNB_FIDX_THREADS
#define SNB_FIDX_THREADS 256
#endif
#ifndef SNB_FIDX_HOPS
#define SNB_FIDX_HOPS 3
#endif
constexpr uint32_t kScanThreads = SNB_FIDX_THREADS;         // 8 CTAs per SM: while one CTA waits for its stage-2 hops the others scan (1 024 threads: 0.73 ms per index, 512: 0.56, 256: 0.53, 128: 0.55)
constexpr uint32_t kScanBytes = 32768;          // stream bytes per CTA
constexpr uint32_t kBlockCands = 1024;          // candidate slots per CTA (after the successor checks: a handful per 32 KiB)
constexpr uint32_t kStage1 = 8192;              // positions per CTA that pass the register-only test (text: ~200 of 32 768)
constexpr uint32_t SNB_MAX_FRAME_DATA = 76490u + 4u;   // codec.nim:216-219 + the CRC
constexpr uint32_t kHops = SNB_FIDX_HOPS;                   // successors a candidate must lead to (or the end of the stream)
constexpr uint32_t kEnd = 0xfffffffeu, kDead = 0xffffffffu;
constexpr uint32_t kMaxSkippable = 1u << 16;    // longest skippable chunk the candidate scan considers

struct Summary {            // device -> host, one small D2H
  uint32_t overflow;        // a CTA found more candidates than it has slots for
  uint32_t header_ok;       // the 10-byte stream identifier is present (when asked for)
  uint32_t n_cand;
  uint32_t chain_len;       // reached headers (data chunks and skippable ones)
  uint32_t reached_end;     // the chain ends exactly at the end of the stream
  uint32_t n_chunks;        // descriptors written (data chunks that fit)
  uint32_t first_unfit;     // index (in chain order) of the first data chunk that does not fit out_cap, or 0xffffffff
  uint32_t pad;
  unsigned long long written;   // output bytes of the listed chunks
  unsigned long long stop_pos;  // stream position where the listing stopped (header of the unfit chunk, or n)
};

struct Args {
  const uint8_t* s;
  unsigned long long n, start, out_cap;
  uint32_t nblocks;
  uint32_t* blk_count;             // [nblocks]
  unsigned long long* blk_off;     // [nblocks] exclusive scan of blk_count
  unsigned long long* cand_tmp;    // [nblocks * kBlockCands]
  unsigned long long* cand;        // [n_cand] sorted positions
  uint32_t* nxt[2];                // successor index (kEnd / kDead / candidate index), double buffered
  uint32_t* mark;                  // reached flag per candidate (+ 2 virtual nodes)
  uint32_t* chain;                 // candidate indices in chain order
  uint32_t* ulen;                  // announced length per chain node (0 for skippable)
  unsigned long long* ooff;        // exclusive scan of ulen
  Chunk* chunks;                   // output descriptors
  uint32_t max_chunks;
  uint32_t check_integrity;
  Summary* sum;
};

__device__ __forceinline__ uint32_t ld_u8(const uint8_t* p) { return __ldg(p); }

// Is there a chunk header at p that the serial walk would accept and step over?  Same checks as walk_framed /
// snappy.nim:199-263.  kind: 0 compressed, 1 stored, 2 skippable.
__device__ __forceinline__ bool header_ok(const uint8_t* s, unsigned long long n, unsigned long long p, uint32_t hdr,
                                          uint32_t* kind, uint32_t* ulen) {
  const uint32_t id = hdr & 0xffu, dlen = hdr >> 8;
  if (n - p < 4 || n - p - 4 < dlen) return false;                // :200-207
  if (id >= 0x80u) { *kind = 2; *ulen = 0; return true; }          // :262-263
  if (id > 1u || dlen < 4) return false;                           // :259-260, :210-211, :238-239
  if (id == 1u) {
    if (dlen - 4 > kMaxBlock) return false;                        // :244-251 (oversize stored chunk: the host decides)
    *kind = 1; *ulen = dlen - 4;
    return true;
  }
  // u32 varint at p + 8 (snappy.nim:92-94 via :216), at most min(5, dlen - 4) bytes
  uint32_t v = 0;
  const uint32_t avail = dlen - 4;
  for (uint32_t i = 0; i < 5; i++) {
    if (i >= avail) return false;
    const uint32_t b = ld_u8(s + p + 8 + i);
    if (i == 4 && (b & 0xf0u)) return false;
    v |= (b & 0x7fu) << (7 * i);
    if (!(b & 0x80u)) {
      if (v > kMaxBlock) return false;                             // :219-222
      *kind = 0; *ulen = v;
      return true;
    }
  }
  return false;
}

__device__ __forceinline__ uint32_t load_hdr(const uint8_t* s, unsigned long long p) {
  return ld_u8(s + p) | (ld_u8(s + p + 1) << 8) | (ld_u8(s + p + 2) << 16) | (ld_u8(s + p + 3) << 24);
}

// scan: candidates of each 32 KiB piece of the stream, in stream order, into the CTA's slots.  Two stages, because the
// full test needs bytes from elsewhere in the stream: done inline, the few lanes that get that far stall their whole warp
// on dependent loads in almost every turn (measured: 290 GB/s of stream; staged: 745 GB/s with a per-position stage 1 of
// ~17 integer instructions, which was 0.56 of the 0.78 ms per 582 MB; stage 1 is now a byte-parallel filter, see below).
__global__ void __launch_bounds__(kScanThreads) scan_kernel(Args a) {
  __shared__ unsigned long long list[kBlockCands];
  __shared__ uint16_t list1[kStage1];
  __shared__ uint32_t cnt, cnt1;
  if (threadIdx.x == 0) { cnt = 0; cnt1 = 0; }
  __syncthreads();
  const unsigned long long lo = a.start + (unsigned long long)blockIdx.x * kScanBytes;
  const unsigned long long hi = lo + kScanBytes < a.n ? lo + kScanBytes : a.n;
  const uint8_t* s = a.s;
  // 16-byte aligned view of the whole stream: stream position p is view byte p + smis.  A thread takes 16 positions per
  // turn (one LDG.128 plus the word after it), the tests run on 32-bit offsets relative to the CTA's first position.
  const uint32_t smis = (uint32_t)((uintptr_t)s & 15);
  const uint4* vb = (const uint4*)(s - smis);
  const unsigned long long v_lo = lo + smis, v_hi = hi + smis;
  const unsigned long long u0 = v_lo >> 4;                                  // first 16-byte unit of this CTA
  const uint32_t nunits = hi > lo ? (uint32_t)(((v_hi + 15) >> 4) - u0) : 0u;
  const unsigned long long u_last = (a.n + smis - 1) >> 4;                  // last unit holding a stream byte
  const uint32_t rel0 = (uint32_t)(v_lo & 15);                              // where `lo` sits inside the first unit
  const uint32_t span = (uint32_t)(hi - lo);                                // positions of this CTA
  const unsigned long long rem64 = a.n - lo;
  const uint32_t rem = rem64 > 0xffffffffull ? 0xffffffffu : (uint32_t)rem64;   // stream bytes from `lo` on (clamped: lengths are 24 bits)
  uint32_t u = threadIdx.x;
  uint4 nX; nX.x = nX.y = nX.z = nX.w = 0;
  uint32_t n4 = 0;
  if (u < nunits) { nX = __ldg(vb + u0 + u); n4 = (u0 + u < u_last) ? __ldg((const uint32_t*)(vb + u0 + u + 1)) : 0u; }
  for (; u < nunits; u += kScanThreads) {
    const uint32_t w[5] = {nX.x, nX.y, nX.z, nX.w, n4};
    {  // the next turn's bytes are on their way while this turn's sixteen positions are tested
      const uint32_t un = u + kScanThreads;
      if (un < nunits) { nX = __ldg(vb + u0 + un); n4 = (u0 + un < u_last) ? __ldg((const uint32_t*)(vb + u0 + un + 1)) : 0u; }
    }
    // stage 1, registers only, in two steps.  First a byte-parallel filter over the sixteen positions at once: whatever
    // header the walk accepts here has a length below 2^17 (data chunks: <= 76 494, skippable ones considered: <= 65 536),
    // so its fourth byte is 0 or 1, and its type byte is 0, 1 or >= 0x80.  "byte <= 1" is bit 7 of
    // ~(((w & 0x7e..) + 0x7f..) | w) in every byte lane; the fourth-byte mask is that shifted down by three bytes.
    // About 2.5 integer instructions per position instead of 17; under 1 % of the positions get past it.
    uint32_t le[5];
#pragma unroll
    for (uint32_t k = 0; k < 5; k++) le[k] = ~(((w[k] & 0x7e7e7e7eu) + 0x7f7f7f7fu) | w[k]) & 0x80808080u;
    uint32_t m16 = 0;                                               // bit j: position j of this unit passes the filter
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      const uint32_t first = le[k] | (w[k] & 0x80808080u);
      const uint32_t fourth = __funnelshift_r(le[k], le[k + 1], 24);
      const uint32_t c = (first & fourth) >> 7;                     // bits 0, 8, 16, 24
      m16 |= ((c * 0x01020408u) >> 24) << (4 * k);                  // ... gathered into four adjacent bits
    }
    while (m16) {   // then the exact type / length test of the few survivors
      const uint32_t j = (uint32_t)__ffs((int)m16) - 1u;
      m16 &= m16 - 1u;
      const uint32_t r = 16 * u + j - rel0;                         // position relative to `lo`; wraps for the bytes before it
      if (r >= span) continue;
      const uint32_t k = j >> 2;
      const uint32_t x0 = k == 0 ? w[0] : k == 1 ? w[1] : k == 2 ? w[2] : w[3];
      const uint32_t x1 = k == 0 ? w[1] : k == 1 ? w[2] : k == 2 ? w[3] : w[4];
      const uint32_t hdr = (j & 3) ? __funnelshift_r(x0, x1, 8 * (j & 3)) : x0;
      const uint32_t id = hdr & 0xffu, dlen = hdr >> 8;
      // A skippable chunk may announce any 24-bit length, so half of all random payload bytes would qualify in a large
      // stream.  Only short ones (identifier repeats, padding) are candidates; a stream with a longer skippable chunk
      // simply does not reach its end on the device and is walked by the host.
      const bool ok = id <= 1u ? (dlen - 4u <= (id == 0u ? SNB_MAX_FRAME_DATA - 4u : kMaxBlock))
                               : (id >= 0x80u && dlen <= kMaxSkippable);
      if (!ok) continue;
      const uint32_t left = rem - r;                                // stream bytes from this position on
      if (left < 4 || left - 4 < dlen) continue;
      const uint32_t slot = atomicAdd(&cnt1, 1u);
      if (slot < kStage1) list1[slot] = (uint16_t)r;
    }
  }
  __syncthreads();
  const uint32_t c1 = cnt1;
  if (c1 > kStage1) { if (threadIdx.x == 0) { a.sum->overflow = 1; a.blk_count[blockIdx.x] = 0; } return; }
  // stage 2, one survivor per thread (no lane waits for another lane's rare path): the reference's full checks, and the
  // next kHops headers the candidate leads to must pass them too (or the stream ends on the way).  Compressed payload is
  // full of byte patterns that pass one test (00 / 01 are the tags of one-byte literals and short copies, ff 00 is JPEG
  // stuffing): after one hop 0.7 % of the positions of a text stream still qualify, after three hardly any but the
  // real headers.  A TRUE header within kHops chunks of a malformed one is dropped by this too: the chain then stops
  // short of the end, and the host walk reports the exact error.
  for (uint32_t i = threadIdx.x; i < c1; i += blockDim.x) {
    const unsigned long long p = lo + list1[i];
    uint32_t kind, ul;
    const uint32_t hdr = load_hdr(s, p);
    if (!header_ok(s, a.n, p, hdr, &kind, &ul)) continue;
    unsigned long long t = p + 4 + (hdr >> 8);
    bool chain_ok = true;
    for (uint32_t hop = 0; hop < kHops && t != a.n; hop++) {
      if (a.n - t < 4) { chain_ok = false; break; }
      const uint32_t h2 = load_hdr(s, t);
      uint32_t k2, u2;
      if (((h2 & 0xffu) >= 0x80u && (h2 >> 8) > kMaxSkippable) || !header_ok(s, a.n, t, h2, &k2, &u2)) { chain_ok = false; break; }
      t += 4 + (h2 >> 8);
    }
    if (!chain_ok) continue;
    const uint32_t slot = atomicAdd(&cnt, 1u);
    if (slot < kBlockCands) list[slot] = p;
  }
  __syncthreads();
  const uint32_t c = cnt;
  if (c > kBlockCands) { if (threadIdx.x == 0) { a.sum->overflow = 1; a.blk_count[blockIdx.x] = 0; } return; }
  // rank sort (c is small): every entry counts the entries before it
  for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
    const unsigned long long v = list[i];
    uint32_t r = 0;
    for (uint32_t j = 0; j < c; j++) r += list[j] < v ? 1u : 0u;
    a.cand_tmp[(unsigned long long)blockIdx.x * kBlockCands + r] = v;
  }
  if (threadIdx.x == 0) {
    a.blk_count[blockIdx.x] = c;
    if (blockIdx.x == 0 && a.start == 10) {                         // the stream identifier (snappy.nim:187-194)
      const uint8_t magic[10] = {0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59};
      uint32_t ok = a.n >= 10 ? 1u : 0u;
      for (uint32_t i = 0; ok && i < 10; i++) ok = ld_u8(s + i) == magic[i] ? 1u : 0u;
      a.sum->header_ok = ok;
    }
  }
}

// gather: CTA slots -> one sorted array.
__global__ void __launch_bounds__(256) gather_kernel(Args a) {
  const uint32_t b = blockIdx.x;
  const uint32_t c = a.blk_count[b];
  const unsigned long long o = a.blk_off[b];
  for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) a.cand[o + i] = a.cand_tmp[(unsigned long long)b * kBlockCands + i];
}

// link: successor of every candidate; mark the start node.
__global__ void __launch_bounds__(256) link_kernel(Args a, uint32_t n_cand) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  const unsigned long long p = a.cand[i];
  const unsigned long long t = p + 4 + (load_hdr(a.s, p) >> 8);
  uint32_t nx = kDead;
  if (t == a.n) nx = kEnd;
  else {
    uint32_t lo = i + 1, hi = n_cand;                               // first candidate at position >= t
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (a.cand[m] < t) lo = m + 1; else hi = m; }
    if (lo < n_cand && a.cand[lo] == t) nx = lo;
  }
  a.nxt[0][i] = nx;
  a.mark[i] = (p == a.start) ? 1u : 0u;
}

// reach, one doubling round: marked nodes mark their current jump target; every jump doubles.
__global__ void __launch_bounds__(256) reach_kernel(Args a, uint32_t n_cand, uint32_t parity) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  const uint32_t* cur = a.nxt[parity];
  uint32_t* nxt = a.nxt[parity ^ 1];
  const uint32_t j = cur[i];
  if (j < n_cand) {
    if (a.mark[i]) a.mark[j] = 1u;
    nxt[i] = cur[j];
  } else {
    nxt[i] = j;
    if (j == kEnd && a.mark[i]) a.sum->reached_end = 1u;   // (only the last chain node ever keeps kEnd as its 1-step target
  }                                                         //  through all rounds; earlier ones reach it by jumping)
}

// All the doubling rounds in one launch when one CTA can sweep the candidates quickly (a 582 MB text stream has 16 600:
// seventeen launches of 3.6 us each become one).  The barrier between rounds is all the ordering the rounds need: a
// node marked during round k only has to act from round k + 1 on.  (A copy of the tables in shared memory makes the
// rounds three times faster, 0.516 -> 0.490 ms per index, but that benign mark race then shows up as racecheck
// warnings; not worth it.)
__global__ void __launch_bounds__(1024) reach_fused_kernel(Args a, uint32_t n_cand, uint32_t rounds) {
  for (uint32_t k = 0; k <= rounds; k++) {
    const uint32_t* cur = a.nxt[k & 1u];
    uint32_t* nxt = a.nxt[(k & 1u) ^ 1u];
    for (uint32_t i = threadIdx.x; i < n_cand; i += blockDim.x) {
      const uint32_t j = cur[i];
      if (j < n_cand) {
        if (a.mark[i]) a.mark[j] = 1u;
        nxt[i] = cur[j];
      } else {
        nxt[i] = j;
        if (j == kEnd && a.mark[i]) a.sum->reached_end = 1u;
      }
    }
    __syncthreads();
  }
}

// chain: reached candidates in order (they are sorted by position, and the walk only moves forward).
__global__ void __launch_bounds__(256) chain_count_kernel(Args a, uint32_t n_cand) {
  __shared__ uint32_t c;
  if (threadIdx.x == 0) c = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_cand && a.mark[i]) atomicAdd(&c, 1u);
  __syncthreads();
  if (threadIdx.x == 0) a.blk_count[blockIdx.x] = c;
}
__global__ void __launch_bounds__(256) chain_write_kernel(Args a, uint32_t n_cand) {
  __shared__ uint32_t flags[256];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  flags[threadIdx.x] = (i < n_cand && a.mark[i]) ? 1u : 0u;
  __syncthreads();
  if (!flags[threadIdx.x]) return;
  uint32_t r = 0;
  for (uint32_t j = 0; j < threadIdx.x; j++) r += flags[j];
  a.chain[a.blk_off[blockIdx.x] + r] = i;
}

// describe, step 1: announced length of every chain node.
__global__ void __launch_bounds__(256) ulen_kernel(Args a, uint32_t chain_len) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= chain_len) return;
  const unsigned long long p = a.cand[a.chain[i]];
  uint32_t kind = 2, ul = 0;
  header_ok(a.s, a.n, p, load_hdr(a.s, p), &kind, &ul);
  a.ulen[i] = ul;
}
// step 2: first data chunk that does not fit the caller's room (snappy.nim:219-229, :253-254).
__global__ void __launch_bounds__(256) fit_kernel(Args a, uint32_t chain_len) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= chain_len) return;
  if (a.ulen[i] && a.ooff[i] + a.ulen[i] > a.out_cap) atomicMin(&a.sum->first_unfit, i);
}
// step 3: descriptors of the data chunks before the first unfit one, densely packed (skippable nodes leave no entry).
// blk_count / blk_off are reused for the per-CTA counts of data chunks.
__global__ void __launch_bounds__(256) desc_count_kernel(Args a, uint32_t chain_len) {
  __shared__ uint32_t c;
  if (threadIdx.x == 0) c = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lim = a.sum->first_unfit < chain_len ? a.sum->first_unfit : chain_len;
  if (i < lim) {
    const unsigned long long p = a.cand[a.chain[i]];
    if ((ld_u8(a.s + p) & 0xfeu) == 0) atomicAdd(&c, 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) a.blk_count[blockIdx.x] = c;
}
__global__ void __launch_bounds__(256) desc_write_kernel(Args a, uint32_t chain_len) {
  __shared__ uint32_t flags[256];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lim = a.sum->first_unfit < chain_len ? a.sum->first_unfit : chain_len;
  unsigned long long p = 0;
  uint32_t id = 0xff;
  if (i < lim) { p = a.cand[a.chain[i]]; id = ld_u8(a.s + p); }
  flags[threadIdx.x] = id <= 1u ? 1u : 0u;
  __syncthreads();
  if (i == lim) {   // where the listing stops: the unfit chunk's header, or the end of the chain
    a.sum->stop_pos = i < chain_len ? a.cand[a.chain[i]] : a.n;
    a.sum->written = i < chain_len ? a.ooff[i] : (chain_len ? a.ooff[chain_len - 1] + a.ulen[chain_len - 1] : 0ull);
  }
  if (id > 1u) return;
  uint32_t r = 0;
  for (uint32_t j = 0; j < threadIdx.x; j++) r += flags[j];
  const unsigned long long slot = a.blk_off[blockIdx.x] + r;
  if (slot >= a.max_chunks) return;
  const uint32_t dlen = load_hdr(a.s, p) >> 8;
  Chunk c;
  c.in_off = p + 8; c.out_off = a.ooff[i]; c.in_len = dlen - 4; c.out_cap = a.ulen[i];
  c.crc = load_hdr(a.s, p + 4); c.kind = id == 0 ? kKindCompressed : kKindStored;
  a.chunks[slot] = c;
}

// Chain positions straight from a table of frame sizes (what snb_dev_compress_framed leaves behind): candidate i is
// the i-th frame, already in chain order.  The describe steps validate every header they touch, so a table that does
// not belong to the stream is detected (header_ok fails -> ulen 0 and the sizes do not add up) and the caller falls
// back to the scan.
__global__ void __launch_bounds__(256) hint_kernel(Args a, const unsigned long long* frame_off, const uint32_t* frame_len,
                                                   uint32_t n_frames) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frames) return;
  const unsigned long long p = frame_off[i];
  a.cand[i] = p;
  a.chain[i] = i;
  uint32_t kind = 2, ul = 0;
  const bool ok = p + 4 <= a.n && header_ok(a.s, a.n, p, load_hdr(a.s, p), &kind, &ul) &&
                  (load_hdr(a.s, p) >> 8) + 4 == frame_len[i];
  if (!ok) a.sum->overflow = 1;     // the table does not describe this stream
  if (i + 1 == n_frames && ok && p + frame_len[i] == a.n) a.sum->reached_end = 1;
  if (i == 0 && a.start == 10) {
    const uint8_t magic[10] = {0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59};
    uint32_t okh = a.n >= 10 ? 1u : 0u;
    for (uint32_t k = 0; okh && k < 10; k++) okh = ld_u8(a.s + k) == magic[k] ? 1u : 0u;
    a.sum->header_ok = okh;
  }
}

}  // namespace fidx
}  // namespace snb
