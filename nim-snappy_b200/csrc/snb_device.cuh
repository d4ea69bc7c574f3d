// snb_device.cuh -- sm_100a device code of the B200-native Snappy codec.
//
// One warp owns one block / chunk / message.  All warp collectives are issued convergently
// with the full mask; "uniform" below means "same value in all 32 lanes".
//
// What each routine reproduces (reference file:line, /root/reference @ 8ab5cf0):
//   crc32c_warp          snappy/crc32c.c:204-214,676-763 + snappy/codec.nim:71-75 (maskedCrc)
//   compress_block_v5    snappy/encoder.nim:184-383 (encodeBlock), :44-73, :81-125, :130-182
//   fidx::*              snappy.nim:199-265 header walk on the device (snb_frameidx.cuh)
//   compress_job         snappy/encoder.nim:385-426 (encodeFrame) / snappy.nim:27-64 (compress)
//   decompress_warp      snappy.nim:84-110 (uncompress) + snappy/decoder.nim:20-155 (decodeAllTags)
//
// The file is also compiled by g++ against tests/emu/cuda_emu.h (SNB_EMULATE) so that the
// warp-level logic is exercised on CPU-only CI; that build is a test double, not a fallback.
#pragma once
#include <stdint.h>
#ifdef SNB_EMULATE
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define SNB_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

namespace snb {

constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kMaxBlock = 65536;       // codec.nim:14-19
constexpr uint32_t kInputMargin = 15;       // codec.nim:26
constexpr uint32_t kMinNonLiteral = 17;     // codec.nim:53
constexpr uint32_t kMaxTableEntries = 16384;  // encoder.nim:11-12
constexpr uint32_t kCrcPoly = 0x82F63B78u;  // CRC-32C (Castagnoli), reflected
constexpr uint32_t kCrcMaskDelta = 0xa282ead8u;
constexpr uint32_t kNoAbort = 0xffffffffu;
constexpr uint32_t kAborted = 0xffffffffu;
constexpr uint32_t kMaxDeviceMessage = 1u << 28;  // per-message bound of the warp decoder

enum : int32_t { kOk = 0, kBufferTooSmall = 1, kInvalidInput = 2, kCrcMismatch = 3, kUnknownChunk = 4, kUnsupported = 101 };
enum : uint32_t { kModeFrame = 0, kModeRawBlock = 1, kModeRawMessage = 2 };
enum : uint32_t { kKindCompressed = 0, kKindStored = 1, kKindVerifyOnly = 2, kKindRawBlock = 3 };  // 3: bare tags of one block, out_cap = its exact length

// Optional event counters for tuning (tests/emu builds with -DSNB_STATS; never in the product build).
#ifdef SNB_STATS
__device__ unsigned long long g_stats[16];
#define SNB_STAT(i) do { if (lane == 0) atomicAdd(&g_stats[i], 1ull); } while (0)
#define SNB_STATN(i, n) do { if (lane == 0) atomicAdd(&g_stats[i], (unsigned long long)(n)); } while (0)
#else
#define SNB_STAT(i) do { } while (0)
#define SNB_STATN(i, n) do { } while (0)
#endif

// Optional critical-path timing of the window pass (diagnostic build -DSNB_CYC; never in the product build): cycles
// between "value v has arrived" points, summed per warp in shared memory and flushed to g_cyc at the end of a block.
#if defined(SNB_CYC) && !defined(SNB_EMULATE)
__device__ unsigned long long g_cyc[16];
__device__ __forceinline__ uint32_t cyc_after(uint32_t v) {   // %clock read that cannot issue before v is available
  uint32_t t;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0x7fffffff;\n\t@p mov.u32 %0, %%clock;\n\t@!p mov.u32 %0, 0;\n\t}" : "=r"(t) : "r"(v) : "memory");
  return t;
}
#define SNB_CYC_DECL __shared__ uint32_t cyc_acc[32][16]; const uint32_t cyc_w = threadIdx.x >> 5; uint32_t cyc_t = 0; \
  if (lane < 16) cyc_acc[cyc_w][lane] = 0; __syncwarp(); const uint32_t cyc_begin = cyc_after(lane)
#define SNB_CYC_START(v) cyc_t = cyc_after(v)
#define SNB_CYC_MARK(i, v) do { const uint32_t n_ = cyc_after(v); if (lane == 0) cyc_acc[cyc_w][i] += n_ - cyc_t; cyc_t = n_; } while (0)
#define SNB_CYC_COUNT(i) do { if (lane == 0) cyc_acc[cyc_w][i] += 1; } while (0)
#define SNB_CYC_FLUSH() do { __syncwarp(); if (lane == 0) cyc_acc[cyc_w][15] = cyc_after(lane) - cyc_begin; __syncwarp(); \
  if (lane < 16) atomicAdd(&g_cyc[lane], (unsigned long long)cyc_acc[cyc_w][lane]); } while (0)
#else
#define SNB_CYC_DECL do { } while (0)
#define SNB_CYC_START(v) do { } while (0)
#define SNB_CYC_MARK(i, v) do { } while (0)
#define SNB_CYC_COUNT(i) do { } while (0)
#define SNB_CYC_FLUSH() do { } while (0)
#endif

__device__ __forceinline__ uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }

// x << n with PTX semantics: 0 once n >= 32 (C++ leaves that shift undefined)
#ifdef SNB_EMULATE
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t n) { return n >= 32 ? 0u : x << n; }
#else
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t n) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
  return r;
}
#endif
// index of the lowest non-zero byte of x (x != 0)
__device__ __forceinline__ uint32_t first_diff_byte(uint32_t x) { return (uint32_t)(__ffs((int)x) - 1) >> 3; }
// Cache hints (no destination register, no scoreboard): bring the line holding p towards the SM.
#ifdef SNB_EMULATE
__device__ __forceinline__ void prefetch_l1(const void*) {}
__device__ __forceinline__ void prefetch_l2(const void*) {}
#else
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
#endif
// 32-bit load from a pointer the compiler only knows as generic (a select between the stream and the output)
#ifdef SNB_EMULATE
__device__ __forceinline__ uint32_t ld_global32(const uint32_t* p) { return *p; }
#else
__device__ __forceinline__ uint32_t ld_global32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
#endif
#ifndef SNB_CRC_AFTER
#define SNB_CRC_AFTER 0
#endif
#ifndef SNB_V5_WDEPTH2
#define SNB_V5_WDEPTH2 0   // 1: own words prefetched two windows ahead, so the hint never reads a load in flight (measured: 45.1 -> 44.4 GB/s)
#endif
#ifndef SNB_DEC_PREFETCH
#define SNB_DEC_PREFETCH 256   // bytes ahead of the step that the stream is prefetched into the L1 (0: off -- 6.19 ms per 1 GiB text; 128: 6.04, 256: 6.03, 512: 6.05)
#endif
#ifndef SNB_DEC_FASTFETCH
#define SNB_DEC_FASTFETCH 1   // decoder: unclamped stream loads while the step is >= 320 bytes away from the end of the stream
#endif
#ifndef SNB_DEC_LUT32
#define SNB_DEC_LUT32 1       // decoder: tag table read through a 32-bit shared address
#endif
#ifndef SNB_DEC_SHFLPRED
#define SNB_DEC_SHFLPRED 1    // decoder: prefix sum rounds use shfl.up's range predicate
#endif
#ifndef SNB_DEC_WINDOW
#define SNB_DEC_WINDOW 0   // decoder: 1 = 256 stream bytes in a register window, read through shuffles (the round-1 formulation)
#endif
#ifndef SNB_V5_SKIP_PASSED
#define SNB_V5_SKIP_PASSED 0   // 1 / 2 / 3: no table / candidate loads for the lanes behind the cursor (an eighth of a pass's random
                               // sectors).  Byte-exact and measured SLOWER, reproducibly: 23.74 ms -> 25.05 (branch), 24.77 (branch-free,
                               // the skipped lanes read hot addresses), 25.2 (the ip-1 lane still reads its slot) per 1 GiB of text
#endif
#ifndef SNB_V5_LANDING
#define SNB_V5_LANDING 0   // 1: the landing window's own words are requested as soon as a long match ends (measured neutral: 23.78 -> 23.84 ms)
#endif
#ifndef SNB_V5_HINT_AHEAD2
#define SNB_V5_HINT_AHEAD2 0
#endif
#ifndef SNB_V5_HINT
#define SNB_V5_HINT 1   // 0: none, 1: L1 hint, 2: L2 hint for the next window's candidates
#endif
// Unaligned little-endian 32-bit load built from aligned words.  Only words holding at least
// one requested byte are touched.
__device__ __forceinline__ uint32_t load32u(const uint8_t* p) {
  uintptr_t a = (uintptr_t)p;
  uint32_t sh = (uint32_t)(a & 3) * 8;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  uint32_t lo = w[0];
  uint32_t hi = sh ? w[1] : 0u;
  return __funnelshift_r(lo, hi, sh);
}
// Cooperative byte copy, any alignment, non-overlapping.  dst-aligned 32-bit stores.
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane) {
  if (n < 64) {
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
    return;
  }
  uint32_t head = (uint32_t)((0 - (uintptr_t)dst) & 3);
  if (lane < head) dst[lane] = src[lane];
  uint32_t words = (n - head) >> 2;
  uint32_t* dw = (uint32_t*)(dst + head);
  const uint8_t* s = src + head;
  if ((((uintptr_t)s) & 3) == 0) {
    const uint32_t* sw = (const uint32_t*)s;
    for (uint32_t i = lane; i < words; i += 32) dw[i] = sw[i];
  } else {
    for (uint32_t i = lane; i < words; i += 32) dw[i] = load32u(s + 4 * i);
  }
  uint32_t done = head + 4 * words;
  if (lane < n - done) dst[done + lane] = src[done + lane];
}

// ===================================================================== CRC-32C
// Tables live in global memory (built on the host once), staged into shared memory per CTA.
struct CrcTables {
  uint32_t row[16][256];  // row[j][b] = tab[b] * x^(8*(511-j)): 16 contiguous bytes then a 496-byte skip
  uint32_t byte[256];     // tab[b]
  uint32_t x8[512];       // x^(8k)
  uint32_t xn[132];       // x^(4096q)
  uint32_t ainv[32];      // x^(-128 L)
};
struct CrcSmem {
  uint32_t row[16][256];
  uint32_t byte[256];
};

__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b) {
  uint32_t p = 0;
#pragma unroll 4
  for (int i = 0; i < 32; i++) {
    p ^= (a & 0x80000000u) ? b : 0u;
    a <<= 1;
    b = (b >> 1) ^ ((b & 1u) ? kCrcPoly : 0u);
  }
  return p;
}
__device__ __forceinline__ uint32_t warp_xor(uint32_t v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v ^= __shfl_xor_sync(kFull, v, d);
  return v;
}
__device__ __forceinline__ uint32_t crc_xpow8(const CrcTables* g, uint32_t k) {
  return crc_mulmod(g->xn[k >> 9], g->x8[k & 511]);
}
// F(0, p[0..n)) for n <= 511, lanes take 16-byte pieces.  Uniform result.
__device__ __forceinline__ uint32_t crc_small(const CrcSmem& t, const CrcTables* g, const uint8_t* p, uint32_t n,
                                              uint32_t lane) {
  uint32_t b0 = 16 * lane, c = 0;
  if (b0 < n) {
    uint32_t cnt = umin32(16, n - b0);
    for (uint32_t i = 0; i < cnt; i++) c = t.byte[(c ^ p[b0 + i]) & 0xff] ^ (c >> 8);
    c = crc_mulmod(c, g->x8[n - b0 - cnt]);
  }
  return warp_xor(c);
}
// Finalised (~) CRC-32C of p[0..n), n <= 65536 + 15.  Uniform result.
__device__ __forceinline__ uint32_t crc32c_warp(const CrcSmem& t, const CrcTables* g, const uint8_t* p, uint32_t n,
                                                uint32_t lane) {
  uint32_t head = umin32(n, (uint32_t)((0 - (uintptr_t)p) & 15));
  uint32_t rows = (n - head) >> 9, tail = (n - head) & 511;
  uint32_t acc = 0;
  if (head) acc = crc_mulmod(crc_small(t, g, p, head, lane), crc_xpow8(g, n - head));
  if (rows) {
    const uint4* q = (const uint4*)(p + head) + lane;
    uint32_t c = 0;
    for (uint32_t r = 0; r < rows; r++, q += 32) {
      uint4 v = *q;
      v.x ^= c;
      c = t.row[0][v.x & 0xff] ^ t.row[1][(v.x >> 8) & 0xff] ^ t.row[2][(v.x >> 16) & 0xff] ^ t.row[3][v.x >> 24] ^
          t.row[4][v.y & 0xff] ^ t.row[5][(v.y >> 8) & 0xff] ^ t.row[6][(v.y >> 16) & 0xff] ^ t.row[7][v.y >> 24] ^
          t.row[8][v.z & 0xff] ^ t.row[9][(v.z >> 8) & 0xff] ^ t.row[10][(v.z >> 16) & 0xff] ^ t.row[11][v.z >> 24] ^
          t.row[12][v.w & 0xff] ^ t.row[13][(v.w >> 8) & 0xff] ^ t.row[14][(v.w >> 16) & 0xff] ^ t.row[15][v.w >> 24];
    }
    c = warp_xor(crc_mulmod(c, g->ainv[lane]));
    acc ^= crc_mulmod(c, g->x8[tail]);
  }
  if (tail) acc ^= crc_small(t, g, p + head + rows * 512, tail, lane);
  acc ^= crc_mulmod(0xffffffffu, crc_xpow8(g, n));
  return ~acc;
}
__device__ __forceinline__ uint32_t crc_mask(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + kCrcMaskDelta; }

__device__ __forceinline__ void crc_stage_tables(CrcSmem* s, const CrcTables* g) {
  const uint32_t* src = &g->row[0][0];
  uint32_t* dst = &s->row[0][0];
  for (uint32_t i = threadIdx.x; i < 17 * 256; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// ===================================================================== compress
__device__ __forceinline__ uint32_t hash14(uint32_t d, uint32_t mask) {  // encoder.nim:36-37
  return ((d * 0x1e35a7bdu) >> 18) & mask;
}

// Encoded size of an n-byte literal (tag + length bytes + payload), encoder.nim:57-73.
__device__ __forceinline__ uint32_t literal_size(uint32_t n) { return n + (n <= 60 ? 1 : n <= 256 ? 2 : 3); }

// encoder.nim:81-125.  Returns bytes written (uniform).
__device__ __forceinline__ uint32_t emit_copy_warp(uint8_t* o, uint32_t off, uint32_t len, uint32_t lane) {
  uint32_t w = 0;
  if (len >= 68) {  // leading length-64 copies
    uint32_t q = (len - 68) / 64 + 1;
    for (uint32_t i = lane; i < q; i += 32) {
      o[3 * i] = (63 << 2) | 2; o[3 * i + 1] = (uint8_t)off; o[3 * i + 2] = (uint8_t)(off >> 8);
    }
    w = 3 * q; len -= 64 * q;
  }
  if (len > 64) {
    if (lane == 0) { o[w] = (59 << 2) | 2; o[w + 1] = (uint8_t)off; o[w + 2] = (uint8_t)(off >> 8); }
    w += 3; len -= 60;
  }
  if (len >= 12 || off >= 2048) {
    if (lane == 0) { o[w] = (uint8_t)(((len - 1) << 2) | 2); o[w + 1] = (uint8_t)off; o[w + 2] = (uint8_t)(off >> 8); }
    return w + 3;
  }
  if (lane == 0) { o[w] = (uint8_t)(((off >> 8) << 5) | ((len - 4) << 2) | 1); o[w + 1] = (uint8_t)off; }
  return w + 2;
}

// --------------------------------------------------------------------- input view and emit helpers of the window encoders
struct InView {
  const uint32_t* w;  // 4-byte aligned base (global memory)
  uint32_t mis;       // byte offset of position 0 inside w[0]
  uint32_t last;      // index of the last word holding input bytes
};
__device__ __forceinline__ uint32_t in_ld4(const InView& v, uint32_t p) {
  uint32_t i = p + v.mis, k = i >> 2;
  uint32_t lo = __ldg(v.w + k), hi = __ldg(v.w + umin32(k + 1, v.last));
  return __funnelshift_r(lo, hi, (i & 3) * 8);
}
__device__ __forceinline__ uint32_t in_byte(const InView& v, uint32_t p) {
  return __ldg((const uint8_t*)v.w + v.mis + p);
}
// common prefix of positions s1 < s2, s2 side bounded by len (uniform result); 128 bytes per step
__device__ __forceinline__ uint32_t common_prefix_v2(const InView& v, uint32_t len, uint32_t s1, uint32_t s2,
                                                     uint32_t lane) {
  uint32_t matched = 0;
  for (;;) {
    uint32_t o = 4 * lane, rem = len - s2;
    uint32_t x;
    if (rem >= o + 4) {
      x = in_ld4(v, s1 + o) ^ in_ld4(v, s2 + o);
    } else {
      uint32_t nv = rem > o ? rem - o : 0;
      x = 0;
      for (uint32_t i = 0; i < nv; i++) x |= (in_byte(v, s1 + o + i) ^ in_byte(v, s2 + o + i)) << (8 * i);
      x |= 1u << (8 * nv);
    }
    uint32_t mism = __ballot_sync(kFull, x != 0);
    if (mism) {
      int l = __ffs((int)mism) - 1;
      uint32_t xl = __shfl_sync(kFull, x, l);
      return matched + 4 * (uint32_t)l + ((uint32_t)(__ffs((int)xl) - 1) >> 3);
    }
    matched += 128; s1 += 128; s2 += 128;
  }
}
__device__ __forceinline__ uint32_t emit_literal_v2(uint8_t* o, const InView& v, uint32_t s, uint32_t n, uint32_t lane) {
  uint32_t m = (n - 1) & 0xffff, w;
  if (m < 60) { w = 1; if (lane == 0) o[0] = (uint8_t)(m << 2); }
  else if (m < 256) { w = 2; if (lane == 0) { o[0] = 60 << 2; o[1] = (uint8_t)m; } }
  else { w = 3; if (lane == 0) { o[0] = 61 << 2; o[1] = (uint8_t)m; o[2] = (uint8_t)(m >> 8); } }
  if (n <= 32) {
    if (lane < n) o[w + lane] = (uint8_t)in_byte(v, s + lane);
  } else {
    warp_copy(o + w, (const uint8_t*)v.w + v.mis + s, n, lane);
  }
  return w + n;
}
__device__ __forceinline__ uint32_t emit_copy_v2(uint8_t* o, uint32_t off, uint32_t len, uint32_t lane) {
  if (len <= 64) {  // one tag (encoder.nim:113-125)
    bool c2 = (len >= 12) || (off >= 2048);
    uint32_t b0 = c2 ? (((len - 1) << 2) | 2) : (((off >> 8) << 5) | ((len - 4) << 2) | 1);
    if (lane == 0) {
      o[0] = (uint8_t)b0; o[1] = (uint8_t)off;
      if (c2) o[2] = (uint8_t)(off >> 8);
    }
    return c2 ? 3u : 2u;
  }
  return emit_copy_warp(o, off, len, lane);
}

// 12 bytes at p, p + 12 <= len (no clamping needed: every touched word holds an input byte)
__device__ __forceinline__ void in_ld12u(const InView& v, uint32_t p, uint32_t& b0, uint32_t& b1, uint32_t& b2) {
  uint32_t i = p + v.mis, k = i >> 2, s = (i & 3) * 8;
  uint32_t w0 = __ldg(v.w + k), w1 = __ldg(v.w + k + 1), w2 = __ldg(v.w + k + 2), w3 = __ldg(v.w + umin32(k + 3, v.last));
  b0 = __funnelshift_r(w0, w1, s); b1 = __funnelshift_r(w1, w2, s); b2 = __funnelshift_r(w2, w3, s);
}
// The same 12 bytes through 16-byte aligned loads: one LDG.128, and a second one only in the lanes whose bytes
// straddle it (1.7 load instructions per lane instead of 4).  Reads stay inside the 16-byte aligned envelope of
// [in, in + len).  A/B switch, off: see SNB_V5_CAND128.
#ifndef SNB_V5_CAND128
#define SNB_V5_CAND128 0   // measured neutral (45.25 -> 45.42 GB/s): the tag stage is at 23 %, not the limiter the sector count suggested
#endif
__device__ __forceinline__ void in_ld12a(const InView& v, uint32_t p, uint32_t& b0, uint32_t& b1, uint32_t& b2) {
  const uintptr_t addr = (uintptr_t)v.w + v.mis + p;
  const uint4* a16 = (const uint4*)(addr & ~(uintptr_t)15);
  const uint32_t q = (uint32_t)(addr & 15);
  const uint4 X = __ldg(a16);
  uint4 Y; Y.x = Y.y = Y.z = Y.w = 0;
  if (q > 4) Y = __ldg(a16 + 1);
  const bool hi2 = (q & 8) != 0, hi1 = (q & 4) != 0;
  // the five words from word (q >> 2) & 2 on, then the four from word q >> 2 on
  const uint32_t p0 = hi2 ? X.z : X.x, p1 = hi2 ? X.w : X.y, p2 = hi2 ? Y.x : X.z, p3 = hi2 ? Y.y : X.w, p4 = hi2 ? Y.z : Y.x;
  const uint32_t w0 = hi1 ? p1 : p0, w1 = hi1 ? p2 : p1, w2 = hi1 ? p3 : p2, w3 = hi1 ? p4 : p3;
  const uint32_t sh = (q & 3) * 8;
  b0 = __funnelshift_r(w0, w1, sh); b1 = __funnelshift_r(w1, w2, sh); b2 = __funnelshift_r(w2, w3, sh);
}
// --------------------------------------------------------------------- encodeBlock, v5
// Window walk on ballot masks.  One pass handles the 32 positions [b, b+32) that hold the cursor:
//
//   resolve  lane l owns position b+l: 12 own bytes (one coalesced word per lane + shuffles, the next
//            window's words prefetched), hash, table entry T, 12 candidate bytes, match length ml
//            (0..12), the copy tag a 4..11-byte match would emit, and the match.any group of the hash.
//   walk     as a chain arrival (post-copy lookup, encoder.nim:364-380) every lane knows its event --
//            the first hit lane at or after itself (HB & lanes >= me) -- and so the next arrival
//            J = event + ml.  The warp follows J from the entry arrival with one shuffle per arrival
//            and collects the arrival mask A.
//   masks    everything else is a function of A: looked-up lanes LK = union of [arrival, event],
//            inserted lanes = LK | (A >> 1) (encoder.nim:371-377, :293-294), copy events = LK & hits,
//            literal bytes = LK & ~events, literal starts = A & ~hits.  Output offsets are popcounts
//            of those masks below the lane; every lane writes its own tag / literal byte.
//   verify   a lookup is exact unless an earlier lane of its hash group was inserted in this window
//            (then that lane is the true candidate): the arrivals before the first such lane are
//            kept, the offending one is redone by a serial step straight on the table.  Matches that
//            reach the 12-byte horizon end a batch and are measured with the whole warp.
//   commit   the last inserted lane of each hash group writes the table.
//
// Chunk tails, strided literal runs (>= 32 probes, encoder.nim:311-331) and the rare conflicts take the
// serial steps (one chain lookup, or 32 strided probes, straight on the table).
__device__ __forceinline__ uint32_t compress_block_v5(const uint8_t* __restrict__ in, uint32_t len,
                                                      uint8_t* __restrict__ out, uint16_t* table,
                                                      uint32_t abort_thr, uint32_t lane) {
  InView v;
  v.mis = (uint32_t)((uintptr_t)in & 3);
  v.w = (const uint32_t*)(in - v.mis);
  v.last = (len + v.mis - 1) >> 2;
  if (len < kMinNonLiteral) return emit_literal_v2(out, v, 0, len, lane);

  uint32_t tsize = 256;
  while (tsize < kMaxTableEntries && tsize < len) tsize <<= 1;
  const uint32_t mask = tsize - 1;
  {
    uint32_t* t32 = (uint32_t*)table;
    for (uint32_t i = lane; i < tsize / 2; i += 32) t32[i] = 0;
  }
  const uint32_t ip_limit = len - kInputMargin;
  const uint32_t below = (1u << lane) - 1u;          // lanes before mine
  const uint32_t notbelow = ~below;                  // my lane and the ones after it
  const uint32_t own_q = (lane + v.mis) >> 2, own_s = ((lane + v.mis) & 3) * 8;
  const uint32_t wlane = umin32(lane, 11u);          // 12 words hold a window's 44 bytes
  uint32_t op = 0;
  uint32_t ip = 0;          // chain mode: position of the pending post-copy lookup (<= ip_limit)
  uint32_t next_emit = 0;   // literal mode: start of the pending literal
  uint32_t pp = 1;          // literal mode: next probe position
  uint32_t k = 0;           // literal mode: probes already done in this run
  bool chain = false, force_slow = false;
  bool m1_done = false;     // chain mode, ip on a window boundary: ip-1 has already been inserted
  constexpr uint32_t kNoWindow = 0xffffffffu;
  uint32_t Wn = 0, Wn_b = kNoWindow;   // prefetched own words of window Wn_b
#if SNB_V5_WDEPTH2
  uint32_t Wm = 0, Wm_b = kNoWindow;   // ... and of the window after it
#endif
  SNB_CYC_DECL;

#define SNB_FINISH(from)                                                          \
  do {                                                                            \
    uint32_t f_ = (from);                                                         \
    if (f_ < len) {                                                               \
      uint32_t n_ = len - f_;                                                     \
      if (op + literal_size(n_) > abort_thr) { SNB_CYC_FLUSH(); return kAborted; } \
      op += emit_literal_v2(out + op, v, f_, n_, lane);                           \
    }                                                                             \
    SNB_CYC_FLUSH();                                                              \
    return op;                                                                    \
  } while (0)

  for (;;) {
    __syncwarp();  // table writes (zeroing / previous commit / serial step) visible to every lane
    const uint32_t b = (chain ? ip : pp) & ~31u;
    SNB_CYC_START(b);

    SNB_STAT(8);
    if (force_slow || b + 32 >= ip_limit || (!chain && k >= 32)) {
      force_slow = false;
      m1_done = false;
      SNB_CYC_COUNT(chain ? 9 : 10);
      if (chain) {
        SNB_STAT(1);
        // ---- one serial chain step straight on the table (encoder.nim:364-380)
        const uint32_t d1 = in_ld4(v, ip - 1), d = in_ld4(v, ip);
        const uint32_t h1 = hash14(d1, mask), hh = hash14(d, mask);
        const uint32_t tv = table[hh];
        const uint32_t cand = (hh == h1) ? ip - 1 : tv;
        __syncwarp();  // every lane has read before lane 0 writes
        if (lane == 0) { table[h1] = (uint16_t)(ip - 1); table[hh] = (uint16_t)ip; }
        const uint32_t m = common_prefix_v2(v, len, cand, ip, lane);
        if (m < 4) { chain = false; next_emit = ip; pp = ip + 1; k = 0; continue; }   // :379-380
        op += emit_copy_v2(out + op, ip - cand, m, lane);
        ip += m;
        if (op > abort_thr) return kAborted;
        if (ip > ip_limit) SNB_FINISH(ip);
        continue;
      }
      // ---- serial literal scan from (pp, k): 32 strided probes per step (encoder.nim:281-331)
      SNB_STAT(2);
      uint32_t wip = pp, wskip = 32 + k;
      uint32_t hit_ip = 0, hit_cand = 0;
      bool found = false;
      for (;;) {
        uint32_t pos = wip, sk = wskip;
        for (uint32_t j = 0; j < lane; j++) { uint32_t st = sk >> 5; sk += st; pos += st; }
        uint32_t step = sk >> 5, sk_after = sk + step;
        bool valid = (pos <= ip_limit) && (pos + step <= ip_limit);
        uint32_t d = 0, h = 0x10000u + lane, tv = 0;
        if (valid) { d = in_ld4(v, pos); h = hash14(d, mask); tv = table[h]; }
        uint32_t grp = __match_any_sync(kFull, h);
        uint32_t earlier = grp & below;
        int src = earlier ? 31 - __clz((int)earlier) : (int)lane;
        uint32_t pe = __shfl_sync(kFull, pos, src);
        uint32_t c = earlier ? pe : tv;
        bool hit = valid && (in_ld4(v, c) == d);
        uint32_t hitmask = __ballot_sync(kFull, hit);
        uint32_t vmask = __ballot_sync(kFull, valid);
        if (hitmask) {
          int kh = __ffs((int)hitmask) - 1;
          uint32_t commit = (2u << kh) - 1u;
          __syncwarp();   // every lane's table read of this step precedes the writes (WAR; the collectives above already imply it)
          if (((commit >> lane) & 1u) && ((grp & commit) >> lane) == 1u) table[h] = (uint16_t)pos;
          hit_ip = __shfl_sync(kFull, pos, kh);
          hit_cand = __shfl_sync(kFull, c, kh);
          found = true;
          break;
        }
        if (vmask != kFull) break;
        __syncwarp();
        if ((grp >> lane) == 1u) table[h] = (uint16_t)pos;
        wip = __shfl_sync(kFull, pos + step, 31);
        wskip = __shfl_sync(kFull, sk_after, 31);
        __syncwarp();
      }
      if (!found) SNB_FINISH(next_emit);
      {
        uint32_t n = hit_ip - next_emit;
        if (op + literal_size(n) > abort_thr) return kAborted;
        op += emit_literal_v2(out + op, v, next_emit, n, lane);
      }
      uint32_t m = 4 + common_prefix_v2(v, len, hit_cand + 4, hit_ip + 4, lane);
      op += emit_copy_v2(out + op, hit_ip - hit_cand, m, lane);
      ip = hit_ip + m;
      if (op > abort_thr) return kAborted;
      if (ip > ip_limit) SNB_FINISH(ip);
      chain = true;
      continue;
    }

    // ---- window [b, b+32): every lane is a valid lookup position (b + 32 < ip_limit)
    const uint32_t p = b + lane;
    if (chain && ip == b && !m1_done) {  // arrival on the window boundary after a serial step: ip-1 is not a lane here
      const uint32_t d1 = in_ld4(v, ip - 1);
      if (lane == 0) table[hash14(d1, mask)] = (uint16_t)(ip - 1);
      __syncwarp();
    }
    m1_done = false;

    // ---- resolve
    SNB_STAT(0);
    const uint32_t W = (Wn_b == b) ? Wn : __ldg(v.w + (b >> 2) + wlane);
#if SNB_V5_WDEPTH2
    Wn = (Wm_b == b + 32) ? Wm : __ldg(v.w + umin32((b >> 2) + 8 + wlane, v.last));
    Wn_b = b + 32;
    Wm = __ldg(v.w + umin32((b >> 2) + 16 + wlane, v.last));
    Wm_b = b + 64;
#else
    Wn = __ldg(v.w + umin32((b >> 2) + 8 + wlane, v.last));
    Wn_b = b + 32;
#endif
    const uint32_t w0 = __shfl_sync(kFull, W, (int)own_q), w1 = __shfl_sync(kFull, W, (int)own_q + 1);
    const uint32_t w2 = __shfl_sync(kFull, W, (int)own_q + 2), w3 = __shfl_sync(kFull, W, (int)own_q + 3);
    const uint32_t a0 = __funnelshift_r(w0, w1, own_s), a1 = __funnelshift_r(w1, w2, own_s), a2 = __funnelshift_r(w2, w3, own_s);
    const uint32_t h = hash14(a0, mask);
#if SNB_V5_SKIP_PASSED
    // lanes before the pass's entry position are behind the cursor: they are never looked up in this pass (they only
    // lend their hash to the insert of ip-1 and to the group masks), so their table entry and candidate bytes are not
    // fetched -- an eighth of a pass's random sectors on text (the entry sits 4 lanes into the window on average)
    const uint32_t c_in = (chain ? ip : pp) - b;
    const bool need = lane >= c_in;
    // the lane just before the entry is inserted by this pass (ip-1): it still reads its slot, so that the commit's
    // 16-bit store lands in a sector the L2 holds completely (a partial-sector write miss is a DRAM read-modify-write)
    const bool need_t = SNB_V5_SKIP_PASSED == 3 ? lane + 1 >= c_in : need;
#else
    const bool need = true, need_t = true;
#endif
#if SNB_V5_SKIP_PASSED == 2
    // branch-free form: the skipped lanes read slot 0 and the window's own first bytes (both hot), and report no match
    const uint32_t T = need ? table[h] : table[0];
    const uint32_t Tc = need ? T : b;
#else
    const uint32_t T = need_t ? table[h] : 0u;
    const uint32_t Tc = T;
#endif
    SNB_CYC_MARK(0, T);
    uint32_t ml = 0;
    if (SNB_V5_SKIP_PASSED == 2 || need) {
      uint32_t c0, c1, c2;
#if SNB_V5_CAND128
      in_ld12a(v, Tc, c0, c1, c2);
#else
      in_ld12u(v, Tc, c0, c1, c2);
#endif
      const uint32_t x0 = a0 ^ c0, x1 = a1 ^ c1, x2 = a2 ^ c2;
      const uint32_t xs = x0 ? x0 : (x1 ? x1 : x2);
      const uint32_t base = x0 ? 0u : (x1 ? 4u : 8u);
      ml = xs ? base + first_diff_byte(xs) : 12u;
      if (!need) ml = 0;
    }
    SNB_CYC_MARK(1, ml);
    const uint32_t off = p - T;
    const bool c3 = off >= 2048;   // three-byte copy tag (ml <= 11 whenever the tag is used, encoder.nim:113-125)
    const uint32_t tb = c3 ? ((((ml - 1) << 2) | 2u) | (off << 8))
                           : ((((off >> 8) << 5) | ((ml - 4) << 2) | 1u) | ((off & 0xff) << 8));
    const uint32_t grp = __match_any_sync(kFull, h);
    const uint32_t eg = grp & below;                   // earlier lanes of my hash group
    const uint32_t HB = __ballot_sync(kFull, ml >= 4);
    const uint32_t C3 = __ballot_sync(kFull, c3);
    // every lane as a chain arrival: its event lane, the looked-up lanes, the next arrival
    const uint32_t qb = HB & notbelow;
    const uint32_t evt = qb ? (uint32_t)__ffs((int)qb) - 1u : 32u;
    const uint32_t eml = __shfl_sync(kFull, ml, (int)(evt & 31u));
    // J < 32: next arrival in this window; 32..63: beyond it; 64 + e: the match at lane e reaches the horizon;
    // 128: no hit up to the window end (the literal run goes on)
    const uint32_t J = evt >= 32 ? 128u : (eml >= 12 ? 64u + evt : evt + eml);
    const uint32_t lkmask = notbelow & ((2u << umin32(evt, 31u)) - 1u);
    SNB_CYC_MARK(2, J + lkmask + grp);
#if SNB_V5_HINT
    // hint for the next window: its (pre-commit) table entries say where its candidates most likely are
    uint32_t Tn;
    {
#if SNB_V5_HINT_AHEAD2   // (needs SNB_V5_WDEPTH2) hint two windows ahead: a whole pass for the table sector and the candidate line to arrive
      const uint32_t n0 = __shfl_sync(kFull, Wm, (int)own_q), n1 = __shfl_sync(kFull, Wm, (int)own_q + 1);
#else
      const uint32_t n0 = __shfl_sync(kFull, Wn, (int)own_q), n1 = __shfl_sync(kFull, Wn, (int)own_q + 1);
#endif
      Tn = table[hash14(__funnelshift_r(n0, n1, own_s), mask)];
    }
    SNB_CYC_MARK(3, Tn);
#endif

    uint32_t ins = 0;      // lanes inserted in this pass
    uint32_t c = 0;        // entry arrival lane
    uint32_t noprev = 0;   // arrival whose predecessor lane is not inserted again (a probe hit)
    if (!chain) {
      // ---- the literal run entering the window: stride-1 probes at lanes [l0, bend)
      const uint32_t l0 = pp - b;
      const uint32_t bend = umin32(32u, l0 + (32u - k));
      const uint32_t range = (0xffffffffu << l0) & (bend >= 32 ? 0xffffffffu : ((1u << bend) - 1u));
      const uint32_t qh = HB & range;
      const uint32_t q0 = qh ? (uint32_t)__ffs((int)qh) - 1u : 32u;
      const uint32_t run = range & (q0 >= 32 ? 0xffffffffu : ((2u << q0) - 1u));   // probes performed
      const uint32_t stale = __ballot_sync(kFull, (eg & run) != 0) & run;
      if (stale) { SNB_STAT(5); force_slow = true; continue; }
      if (q0 >= 32) {  // every probe missed: insert them and move on (k == 32 switches to the strided scan)
        SNB_STAT(6);
        __syncwarp();
        if (((run >> lane) & 1u) && ((grp & run) >> lane) == 1u) table[h] = (uint16_t)p;
        k += (uint32_t)__popc(run);
        pp = b + bend;
        continue;
      }
      {  // :336-338; the copy at q0 is the first event of the walk
        const uint32_t n = b + q0 - next_emit;
        if (op + literal_size(n) > abort_thr) return kAborted;
        op += emit_literal_v2(out + op, v, next_emit, n, lane);
      }
      ins = run;
      chain = true;
      c = q0;
      noprev = 1u << q0;
    } else {
      c = ip - b;
    }

    bool done = false;  // block finished inside this pass (ip > ip_limit after a copy)
    SNB_CYC_MARK(4, c + ins);
    for (;;) {          // one batch per turn; a match that reaches the horizon ends a batch
      SNB_STAT(3);
      uint32_t A = 0, cc = c;
      do { A |= 1u << cc; cc = __shfl_sync(kFull, J, (int)cc); } while (cc < 32);
      bool onp = (A >> lane) & 1u;
      uint32_t LK = __reduce_or_sync(kFull, onp ? lkmask : 0u);
      uint32_t insb = LK | ((A & ~noprev) >> 1);
      const uint32_t stale = __ballot_sync(kFull, ((LK >> lane) & 1u) && (eg & (ins | insb)) != 0);
      uint32_t slow_at = 32;
      if (stale) {  // keep the arrivals whose lookups all precede the first stale lane
        SNB_STAT(9);
        const uint32_t xs = (uint32_t)__ffs((int)stale) - 1u;
        slow_at = 31u - (uint32_t)__clz((int)(A & ((2u << xs) - 1u)));
        A &= (1u << slow_at) - 1u;
        onp = (A >> lane) & 1u;
        LK = __reduce_or_sync(kFull, onp ? lkmask : 0u);
        insb = LK | ((A & ~noprev) >> 1);
        cc = 32;
      }
      ins |= insb;
      noprev = 0;
      SNB_STATN(7, __popc(A));
      SNB_CYC_MARK(5, ins + stale);
      const uint32_t Ev = LK & HB;
      uint32_t EvT = Ev, tail = 0;
      if (cc >= 64 && cc < 96) EvT &= ~(1u << (cc - 64));                        // its copy is emitted below
      if (cc == 128) tail = 0xffffffffu << (31u - (uint32_t)__clz((int)A));      // the last run leaves the window
      const uint32_t LitB = LK & ~Ev & ~tail;    // literal bytes
      const uint32_t LS = A & ~HB & ~tail;       // literal starts (one tag byte each: n <= 31)
      const uint32_t E3 = EvT & C3;
      {
        const uint32_t offl = (uint32_t)(__popc(LitB & below) + __popc(LS & below) + 2 * __popc(EvT & below) + __popc(E3 & below));
        uint8_t* o = out + op + offl;
        if ((LS >> lane) & 1u) { o[0] = (uint8_t)((evt - lane - 1u) << 2); o += 1; }
        if ((LitB >> lane) & 1u) o[0] = (uint8_t)a0;
        if ((EvT >> lane) & 1u) {
          o[0] = (uint8_t)tb; o[1] = (uint8_t)(tb >> 8);
          if (c3) o[2] = (uint8_t)(tb >> 16);
        }
      }
      op += (uint32_t)(__popc(LitB) + __popc(LS) + 2 * __popc(EvT) + __popc(E3));
      SNB_CYC_MARK(6, op);
      if (op > abort_thr) return kAborted;
      if (slow_at < 32) { ip = b + slow_at; force_slow = true; break; }
      if (cc < 64) { ip = b + cc; break; }
      if (cc == 128) {
        const uint32_t cl = 31u - (uint32_t)__clz((int)A);
        chain = false; next_emit = b + cl; pp = b + 32; k = 31 - cl;
        break;
      }
      // ---- the copy at lane e is at least 12 bytes long: measured with the whole warp (encoder.nim:130-182)
      const uint32_t e = cc - 64;
      SNB_STAT(4);
      const uint32_t cand = __shfl_sync(kFull, T, (int)e);
      const uint32_t m = 12 + common_prefix_v2(v, len, cand + 12, b + e + 12, lane);
      op += emit_copy_v2(out + op, b + e - cand, m, lane);
      ip = b + e + m;
      SNB_CYC_MARK(7, ip);
      SNB_CYC_COUNT(11);
      if (op > abort_thr) return kAborted;
      if (ip > ip_limit) { done = true; break; }
      c = ip - b;
      if (c >= 32) {
#if SNB_V5_LANDING
        // a long match lands beyond the window whose words are already here: ask for the landing window's now, the
        // copy tag and the commit below run while they travel (565 of a text chunk's 2 090 passes start this way)
        if (c >= 64) { Wn_b = ip & ~31u; Wn = __ldg(v.w + umin32((Wn_b >> 2) + wlane, v.last)); }
#endif
        break;
      }
    }
    if (chain && !force_slow && ip > ip_limit) done = true;   // encoder.nim:362-363
    // the next arrival's ip-1 insert when no window pass will own it (lane c-1 of the arrival rule)
    bool pend_insert = false;
    if (chain && !force_slow && !done && (ip & 31u) == 0) {
      if (ip == b + 32) ins |= 1u << 31; else pend_insert = true;
      m1_done = true;
    }
#if SNB_V5_HINT == 1
    prefetch_l1(v.w + ((Tn + v.mis) >> 2));
#elif SNB_V5_HINT == 2
    prefetch_l2(v.w + ((Tn + v.mis) >> 2));
#endif
    // ---- commit: the last inserted lane of each hash group writes the table
    __syncwarp();   // after every lane's reads of this pass (WAR: racecheck-clean; the ballots above already order them in practice)
    if (((ins >> lane) & 1u) && ((grp & ins) >> lane) == 1u) table[h] = (uint16_t)p;
    if (done) SNB_FINISH(ip);
    if (pend_insert) {
      __syncwarp();
      const uint32_t d1 = in_ld4(v, ip - 1);
      if (lane == 0) table[hash14(d1, mask)] = (uint16_t)(ip - 1);
    }
    SNB_CYC_MARK(8, ip + pp);
    SNB_CYC_COUNT(12);
  }
#undef SNB_FINISH
}

struct CompressArgs {
  const uint8_t* in_base;
  uint8_t* out_base;
  const uint64_t* in_off;   // null: job i starts at i * 65536 (contiguous stream of total_len bytes)
  const uint32_t* in_len;   // null: min(65536, total_len - i * 65536)
  const uint64_t* out_off;  // null: i * out_stride
  uint64_t total_len;
  uint64_t out_stride;
  const uint32_t* crc;      // masked CRC per job (frame mode)
  uint32_t* out_len;        // bytes produced per job
  uint32_t* counter;        // work queue head (zeroed before launch)
  uint32_t count;
  uint32_t mode;
  uint32_t table_entries;   // per-warp shared-memory table capacity (u16 entries)
  uint32_t defer_stored;    // 1: stored frames get only their header here (compact_kernel copies the payload)
  uint16_t* gtables;        // hash tables of the warps beyond smem_warps (table_entries each)
  uint32_t smem_warps;      // warps per CTA whose table lives in shared memory
};

__device__ __forceinline__ uint32_t put_varint32(uint8_t* o, uint32_t v, uint32_t lane) {
  uint32_t n = 0;
  while (v >= 0x80) { if (lane == 0) o[n] = (uint8_t)(v | 0x80); v >>= 7; n++; }
  if (lane == 0) o[n] = (uint8_t)v;
  return n + 1;
}
__device__ __forceinline__ uint32_t varint32_len(uint32_t v) {
  return v < (1u << 7) ? 1 : v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5;
}

// One job = one encodeFrame (encoder.nim:385-426), one bare encodeBlock, or one whole
// single-block `compress` message (snappy.nim:27-64).
__device__ __forceinline__ uint32_t compress_job(const CompressArgs& a, uint32_t job, uint16_t* table, uint32_t lane,
                                                 uint32_t crc_in = 0, bool publish = true) {
  uint64_t ioff = a.in_off ? a.in_off[job] : (uint64_t)job * kMaxBlock;
  uint32_t len = a.in_len ? a.in_len[job] : (uint32_t)((a.total_len - ioff) < kMaxBlock ? (a.total_len - ioff) : kMaxBlock);
  uint64_t ooff = a.out_off ? a.out_off[job] : (uint64_t)job * a.out_stride;
  const uint8_t* in = a.in_base + ioff;
  uint8_t* out = a.out_base + ooff;
  // where the block goes inside the job's output, and whether a block is encoded at all
  const bool frame = a.mode == kModeFrame;
  const uint32_t hl = a.mode == kModeRawBlock ? 0u : varint32_len(len);
  const uint32_t pre = frame ? 8u + hl : hl;
  const uint32_t thr = frame ? len - len / 8 : kNoAbort;
  uint32_t bl = frame ? kAborted : 0u;
  if (frame ? len >= kMinNonLiteral : len > 0) bl = compress_block_v5(in, len, out + pre, table, thr, lane);  // encoder.nim:401-406
  uint32_t total;
  if (frame) {
    uint32_t crc = a.crc ? a.crc[job] : crc_in;
    uint32_t fl;
    if (bl != kAborted && bl <= thr) {  // :408-416
      put_varint32(out + 8, len, lane);
      fl = hl + bl + 4;
      if (lane == 0) out[0] = 0x00;
    } else {  // :418-426
      fl = len + 4;
      if (lane == 0) out[0] = 0x01;
      if (!a.defer_stored) warp_copy(out + 8, in, len, lane);
    }
    if (lane == 0) {
      out[1] = (uint8_t)fl; out[2] = (uint8_t)(fl >> 8); out[3] = (uint8_t)(fl >> 16);
      out[4] = (uint8_t)crc; out[5] = (uint8_t)(crc >> 8); out[6] = (uint8_t)(crc >> 16); out[7] = (uint8_t)(crc >> 24);
    }
    total = fl + 4;
  } else {
    if (a.mode == kModeRawMessage) put_varint32(out, len, lane);  // snappy.nim:49-50
    total = hl + bl;
  }
  if (publish && lane == 0) a.out_len[job] = total;
  return total;
}

// Persistent kernel: blockDim.x/32 warps per CTA pulling jobs from a global counter.  The first
// `smem_warps` warps keep their hash table in dynamic shared memory; the others use a 32 KiB
// slice of `gtables` (global memory, L2 resident) -- same code, the table pointer is generic.
// The v5 encoder on its own register budget (no staging area: the tables are the only shared memory).
#ifndef SNB_COMPRESS5_MAXT
#define SNB_COMPRESS5_MAXT 704
#endif
constexpr uint32_t kCompress5MaxThreads = SNB_COMPRESS5_MAXT;
__global__ void __launch_bounds__(kCompress5MaxThreads, 1) compress_kernel5(CompressArgs a) {
  SNB_DYN_SMEM(smem);
  uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tglobal = warp >= a.smem_warps;
  uint16_t* table = !tglobal
                        ? (uint16_t*)smem + (size_t)warp * a.table_entries
                        : a.gtables + ((size_t)blockIdx.x * (blockDim.x / 32 - a.smem_warps) + (warp - a.smem_warps)) * a.table_entries;
  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(a.counter, 1u);
    job = __shfl_sync(kFull, job, 0);
    if (job >= a.count) return;
    compress_job(a, job, table, lane);
    __syncwarp();
  }
}

// --------------------------------------------------------------------- streaming compress (one launch per slab)
// One persistent kernel does everything `compressFramed` needs for a slab of chunks (snappy.nim:130-155):
//
//   arrive   a job may start before the whole input is in HBM: the host copies the input in pieces and
//            bumps `arrived` (chunks landed) after each piece; the warp that dequeues chunk j waits for it.
//   crc      masked CRC-32C of the chunk (codec.nim:71-75) by the same warp -- this sequential pass also
//            pulls the chunk into L2 for the parse that follows.
//   encode   compress_block_v5 into the chunk's scratch slot (encodeFrame, encoder.nim:385-426).
//   pack     frames are concatenated in chunk order.  `frontier` holds (next chunk to place, its offset);
//            a warp that has just published its frame length places every consecutive chunk whose length
//            is known (CAS on the frontier), copying slot -> final position.  Nobody ever waits for
//            another warp: the publisher of the chunk the frontier is stuck at picks up the backlog.
//   signal   chunks are grouped (1 << group_shift); when every chunk of a group has been placed the
//            group's end offset is written to host-mapped memory, and the host starts the D2H of that
//            byte range while later chunks are still being encoded.
struct StreamArgs {
  CompressArgs c;                        // encoder geometry; c.out_base = scratch slots, c.out_len = frame lengths (zeroed)
  uint8_t* dst;                          // packed output
  unsigned long long* frontier;          // (offset << 24) | next chunk
  const uint32_t* arrived;               // chunks whose input has landed (null: the input is resident)
  uint32_t* group_done;                  // per group: chunks placed (zeroed)
  unsigned long long* group_end;         // per group: end offset, written by whoever places its last chunk
  unsigned long long* host_flags;        // per group, host-mapped: end offset + 1 once the group is complete (null: none)
  unsigned long long* total;             // end offset of the slab (written when the last chunk is placed)
  uint32_t* error;                       // set to 1 when a wait times out
  const CrcTables* crc_tables;
  uint32_t group_shift;
  uint32_t chunk0;                       // index of this slab's first chunk in `arrived` units
  unsigned long long arrive_timeout_ns;  // how long a warp waits for its chunk's input before it gives up (error flag, SNB_TIMEOUT)
  unsigned long long* span;              // optional: [0] = min %globaltimer at CTA start, [1] = max at CTA end (a clock the profiler does not own)
};

#ifdef SNB_EMULATE
__device__ __forceinline__ unsigned long long snb_now_ns() { return 0; }
__device__ __forceinline__ void snb_sleep(unsigned) { emu::yield(); }
__device__ __forceinline__ uint32_t ld_cg32(const uint32_t* p) { return *p; }
#else
__device__ __forceinline__ unsigned long long snb_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void snb_sleep(unsigned ns) { __nanosleep(ns); }
__device__ __forceinline__ uint32_t ld_cg32(const uint32_t* p) { return __ldcg(p); }
#endif
// Release operations at GPU scope.  Unlike __threadfence() (MEMBAR.SC.GPU + CCTL.IVALL in SASS) they do not
// invalidate the SM's L1, which the 22 encoder warps of the CTA live on (candidate loads): MEMBAR.ALL.GPU only.
#ifdef SNB_EMULATE
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { *p = v; }
__device__ __forceinline__ uint32_t atom_add_release_gpu(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
#else
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_release_gpu(uint32_t* p, uint32_t v) {
  uint32_t o;
  asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
  return o;
}
#endif

// Copy n bytes, any alignment, reading through L2 only (the source was written by another SM).
__device__ __forceinline__ void warp_copy_cg(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane) {
  uint32_t head = umin32(n, (uint32_t)((0 - (uintptr_t)dst) & 15));
  if (lane < head) dst[lane] = src[lane];
  const uint32_t vecs = (n - head) >> 4;
  const uint8_t* s = src + head;
  const uint32_t sh = (uint32_t)((uintptr_t)s & 3) * 8;
  const uint32_t* sw = (const uint32_t*)((uintptr_t)s & ~(uintptr_t)3);
  uint4* dv = (uint4*)(dst + head);
  for (uint32_t base = 0; base < vecs; base += 128) {   // uniform trip count: the shuffles below are convergent
    // four 512-byte rows per turn, every load issued before the first use (one L2 round trip per 2 KiB)
    uint32_t w[4][5];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t i = base + 32 * r + lane;
      const bool on = i < vecs;
      w[r][0] = w[r][1] = w[r][2] = w[r][3] = w[r][4] = 0;
      if (on) { w[r][0] = ld_cg32(sw + 4 * i); w[r][1] = ld_cg32(sw + 4 * i + 1); w[r][2] = ld_cg32(sw + 4 * i + 2); w[r][3] = ld_cg32(sw + 4 * i + 3); }
      if (on && sh && (lane == 31 || i + 1 == vecs)) w[r][4] = ld_cg32(sw + 4 * i + 4);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t i = base + 32 * r + lane;
      const bool on = i < vecs;
      const uint32_t nx = __shfl_down_sync(kFull, w[r][0], 1);
      const uint32_t w4 = (lane == 31 || i + 1 == vecs) ? w[r][4] : nx;
      if (on) {
        uint4 v;
        v.x = __funnelshift_r(w[r][0], w[r][1], sh); v.y = __funnelshift_r(w[r][1], w[r][2], sh);
        v.z = __funnelshift_r(w[r][2], w[r][3], sh); v.w = __funnelshift_r(w[r][3], w4, sh);
        dv[i] = v;
      }
    }
  }
  const uint32_t done = head + 16 * vecs;
  if (lane < n - done) dst[done + lane] = src[done + lane];
}

// Claims are batched: the frontier is a serial chain (read frontier -> read length -> CAS, three L2 round trips), so
// one claim per chunk would cap the whole kernel at well under a million chunks per second.  A placer looks at the
// next kPlaceRun lengths at once and moves the frontier over every consecutive chunk whose length is known.
constexpr uint32_t kPlaceRun = 8;
__device__ __forceinline__ void stream_place(const StreamArgs& a, uint32_t lane) {
  const uint32_t count = a.c.count;
  for (;;) {
    unsigned long long f = 0;
    if (lane == 0) f = *(volatile unsigned long long*)a.frontier;
    const uint32_t idx = __shfl_sync(kFull, (uint32_t)(f & 0xffffffull), 0);
    if (idx >= count) return;
    uint32_t len = 0;
    if (lane < kPlaceRun && idx + lane < count) len = *(volatile uint32_t*)(a.c.out_len + idx + lane);
    const uint32_t known = __ballot_sync(kFull, len != 0);
    const uint32_t run = (uint32_t)__ffs((int)~known) - 1u;   // consecutive known lengths from idx on (<= kPlaceRun)
    if (run == 0) return;
    uint32_t incl = lane < run ? len : 0u;
#pragma unroll
    for (int d = 1; d < (int)kPlaceRun; d <<= 1) {
      const uint32_t u = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d) incl += u;
    }
    const uint32_t total = __shfl_sync(kFull, incl, (int)run - 1);
    uint32_t won = 0;
    if (lane == 0) {
      const unsigned long long nf = (((f >> 24) + total) << 24) | (unsigned long long)(idx + run);
      won = atomicCAS(a.frontier, f, nf) == f ? 1u : 0u;
      if (won) __threadfence();   // the slots' bytes were published before their lengths
    }
    won = __shfl_sync(kFull, won, 0);
    if (!won) continue;           // lost the race: look again
    const uint32_t off_lo = __shfl_sync(kFull, (uint32_t)(f >> 24), 0), off_hi = __shfl_sync(kFull, (uint32_t)(f >> 56), 0);
    const unsigned long long off0 = ((unsigned long long)off_hi << 32) | off_lo;
    for (uint32_t r = 0; r < run; r++) {
      const uint32_t ci = idx + r;
      const uint32_t clen = __shfl_sync(kFull, len, (int)r);
      const unsigned long long off = off0 + (__shfl_sync(kFull, incl, (int)r) - clen);
      const uint8_t* slot = a.c.out_base + (unsigned long long)ci * a.c.out_stride;
      uint8_t* d = a.dst + off;
      const uint32_t first = ld_cg32((const uint32_t*)slot) & 0xff;
      if (a.c.mode == kModeFrame && a.c.defer_stored && first == 0x01) {   // stored chunk: header from the slot, payload from the input
        if (lane < 8) d[lane] = slot[lane];
        warp_copy_cg(d + 8, a.c.in_base + (unsigned long long)ci * kMaxBlock, clen - 8, lane);
      } else {
        warp_copy_cg(d, slot, clen, lane);
      }
      __syncwarp();
      if (lane == 0) {
        const uint32_t grp = ci >> a.group_shift;
        const uint32_t gfirst = grp << a.group_shift;
        const uint32_t gsize = umin32(count - gfirst, 1u << a.group_shift);
        const bool last_of_group = ci + 1 == gfirst + gsize;
        if (last_of_group) *(volatile unsigned long long*)(a.group_end + grp) = off + clen;
        if (ci + 1 == count) *(volatile unsigned long long*)a.total = off + clen;
        if (atom_add_release_gpu(a.group_done + grp, 1u) + 1 == gsize && a.host_flags) {   // the chunk's bytes before the count
          __threadfence();
          const unsigned long long end = *(volatile unsigned long long*)(a.group_end + grp);
          __threadfence_system();
          *(volatile unsigned long long*)(a.host_flags + grp) = end + 1;
        }
      }
      __syncwarp();
    }
  }
}

// 22 warps (4 shared-memory + 18 L2 tables) is the measured optimum, so the kernel is compiled for exactly that:
// with a bound of 704 threads ptxas produces a schedule (same 80 registers, no spills) that runs 8 % faster than the
// one for 768 (text 42.6 -> 46.2 GB/s, mixed 61.8 -> 66.8); 672 spills and drops to 35 GB/s.
#ifndef SNB_STREAM_MAXT
#define SNB_STREAM_MAXT 704
#endif
constexpr uint32_t kStreamMaxThreads = SNB_STREAM_MAXT;
__global__ void __launch_bounds__(kStreamMaxThreads, 1) compress_stream_kernel(StreamArgs a) {
  SNB_DYN_SMEM(smem);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tglobal = warp >= a.c.smem_warps;
  uint16_t* table = !tglobal
                        ? (uint16_t*)smem + (size_t)warp * a.c.table_entries
                        : a.c.gtables + ((size_t)blockIdx.x * (blockDim.x / 32 - a.c.smem_warps) + (warp - a.c.smem_warps)) * a.c.table_entries;
  CrcSmem* crc_t = (CrcSmem*)(smem + (size_t)a.c.smem_warps * a.c.table_entries * 2);
  const bool frame = a.c.mode == kModeFrame;
  if (a.span && threadIdx.x == 0) atomicMin(a.span, snb_now_ns());
  if (frame) crc_stage_tables(crc_t, a.crc_tables);
  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(a.c.counter, 1u);
    job = __shfl_sync(kFull, job, 0);
    if (job >= a.c.count) { if (a.span && lane == 0) atomicMax(a.span + 1, snb_now_ns()); return; }
    if (a.arrived) {  // the chunk's bytes may still be on their way
      if (lane == 0) {
        const volatile uint32_t* arr = (const volatile uint32_t*)a.arrived;
        if (*arr <= a.chunk0 + job) {
          const unsigned long long t0 = snb_now_ns();
          while (*arr <= a.chunk0 + job) {
            snb_sleep(500);
            if (snb_now_ns() - t0 > a.arrive_timeout_ns) { *(volatile uint32_t*)a.error = 1; break; }
          }
        }
        __threadfence();
      }
      __syncwarp();
    }
    uint32_t crc = 0;
#if !SNB_CRC_AFTER
    if (frame) {
      const unsigned long long ioff = (unsigned long long)job * kMaxBlock;
      const uint32_t len = (uint32_t)((a.c.total_len - ioff) < kMaxBlock ? (a.c.total_len - ioff) : kMaxBlock);
      crc = crc_mask(crc32c_warp(*crc_t, a.crc_tables, a.c.in_base + ioff, len, lane));
    }
#endif
    const uint32_t total = compress_job(a.c, job, table, lane, crc, false);
#if SNB_CRC_AFTER
    // The CRC pass reads the chunk AFTER the encoder has walked it: the bytes are still in L2 from the parse, and a
    // chunk that has only just been dequeued does not occupy 64 KiB of L2 before its first window is looked at.
    if (frame) {
      const unsigned long long ioff = (unsigned long long)job * kMaxBlock;
      const uint32_t len = (uint32_t)((a.c.total_len - ioff) < kMaxBlock ? (a.c.total_len - ioff) : kMaxBlock);
      crc = crc_mask(crc32c_warp(*crc_t, a.crc_tables, a.c.in_base + ioff, len, lane));
      uint8_t* hdr = a.c.out_base + (unsigned long long)job * a.c.out_stride;
      if (lane < 4) hdr[4 + lane] = (uint8_t)(crc >> (8 * lane));
    }
#endif
    __syncwarp();
    if (lane == 0) {
      st_release_gpu(a.c.out_len + job, total);   // the frame's bytes before its length
      __threadfence();   // ... and the length before this warp looks at the frontier (store -> load: needs the full fence)
    }
    __syncwarp();
    stream_place(a, lane);
    __syncwarp();
  }
}

// Keeps one thread busy for `ns` nanoseconds: the host times the launch CALL of this kernel to find out whether
// launches block (profilers, CUDA_LAUNCH_BLOCKING); a kernel that waits for copies enqueued after it must not be
// launched before them then.
__global__ void launch_probe_kernel(unsigned ns) {
  const unsigned long long t0 = snb_now_ns();
  while (snb_now_ns() - t0 < ns) snb_sleep(1000);
}

// frontier = (base << 24) for the first slab, or continues at *prev_total.
__global__ void stream_init_kernel(unsigned long long* frontier, const unsigned long long* prev_total,
                                   unsigned long long base, unsigned long long* total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const unsigned long long b = prev_total ? *prev_total : base;
    *frontier = b << 24;
    *total = b;
  }
}

// ===================================================================== CRC kernels
struct CrcArgs {
  const uint8_t* base;
  const uint64_t* off;   // null: i * 65536 over total_len
  const uint32_t* len;
  uint64_t total_len;
  uint32_t* out;         // per job CRC
  uint32_t count;
  uint32_t masked;       // 1: snappy-masked, 0: plain finalised CRC-32C
};
__global__ void __launch_bounds__(256) crc_kernel(CrcArgs a, const CrcTables* g) {
  __shared__ CrcSmem t;
  crc_stage_tables(&t, g);
  uint32_t lane = threadIdx.x & 31;
  uint32_t wpb = blockDim.x >> 5;
  for (uint32_t job = blockIdx.x * wpb + (threadIdx.x >> 5); job < a.count; job += gridDim.x * wpb) {
    uint64_t o = a.off ? a.off[job] : (uint64_t)job * kMaxBlock;
    uint32_t n = a.len ? a.len[job] : (uint32_t)((a.total_len - o) < kMaxBlock ? (a.total_len - o) : kMaxBlock);
    uint32_t c = crc32c_warp(t, g, a.base + o, n, lane);
    if (lane == 0) a.out[job] = a.masked ? crc_mask(c) : c;
  }
}

// ===================================================================== scan + compaction
// Exclusive prefix sum of lens[0..count) into offs (u64), starting at base_const + *base_ptr;
// end offset into *total (may alias base_ptr).  One CTA: every thread sums a contiguous run of ceil(count / threads)
// elements, one block-wide scan of the per-thread sums, then the run is written -- three barriers whatever the count
// (the tile-by-tile form this replaces took 1.5 us per 1 024 elements: 29 us for the 17 759 block counts of a frame
// index, 1.4 ms for a batch of a million messages).
__global__ void __launch_bounds__(1024) scan_kernel(const uint32_t* lens, uint64_t* offs, uint32_t count,
                                                    const uint64_t* base_ptr, uint64_t base_const, uint64_t* total) {
  __shared__ unsigned long long warp_sums[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t per = (count + blockDim.x - 1) / blockDim.x;
  const uint64_t lo64 = (uint64_t)threadIdx.x * per;
  const uint32_t lo = lo64 < count ? (uint32_t)lo64 : count;
  const uint32_t hi = count - lo < per ? count : lo + per;
  unsigned long long sum = 0;
#pragma unroll 8
  for (uint32_t i = lo; i < hi; i++) sum += lens[i];
  unsigned long long incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned long long u = __shfl_up_sync(kFull, incl, d);
    if (lane >= (uint32_t)d) incl += u;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {   // inclusive scan of the warp totals
    unsigned long long w = lane < nw ? warp_sums[lane] : 0ull;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long u = __shfl_up_sync(kFull, w, d);
      if (lane >= (uint32_t)d) w += u;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  const unsigned long long base = base_const + (base_ptr ? *base_ptr : 0ull);
  unsigned long long run = base + (warp ? warp_sums[warp - 1] : 0ull) + incl - sum;
  __syncthreads();   // every thread has read *base_ptr before thread 0 may overwrite it through `total`
#pragma unroll 8
  for (uint32_t i = lo; i < hi; i++) { const unsigned long long v = lens[i]; offs[i] = run; run += v; }
  if (threadIdx.x == blockDim.x - 1) *total = base + warp_sums[31 < nw ? 31 : nw - 1];
}

// ===================================================================== decompress
// One chunk / message descriptor.  For framed chunks the host's header walk fills it
// (snappy.nim:199-265); for raw batches the wrapper kernels build it from offset tables.
struct Chunk {
  uint64_t in_off;   // payload: varint+tags (kind 0) or raw bytes (kind 1/2)
  uint64_t out_off;
  uint32_t in_len;
  uint32_t out_cap;  // room the reference would hand to `uncompress` for this chunk
  uint32_t crc;      // stored masked CRC (framed)
  uint32_t kind;
};

// uncompress (snappy.nim:84-110) + decodeAllTags (decoder.nim:20-155) for one message.
// Tags are parsed 32 stream bytes at a time: every lane decodes "a tag starting at my byte",
// pointer jumping over next-tag links marks the real tag starts, a prefix sum over their
// lengths gives output positions, then literals / back-references that do not read bytes
// produced inside this step copy in parallel and the (few) dependent ones follow in order.
// Accept/reject decisions equal the reference's for every input; only bytes past `written`
// (scratch in the reference) are left untouched here.
// Tag byte -> what decoder.nim:39-109 derives from it, as one shared-memory lookup (the branchy derivation is all
// ALU-pipe work, the pipe this kernel saturates): x = L | size << 8 | offset high bits << 16, y = mask of the
// offset bytes that follow the tag (0 for a literal).  Literals with a length field (tag >> 2 >= 60) have L = 0.
__device__ __forceinline__ uint2 tag_entry(uint32_t t) {
  const uint32_t kind = t & 3;
  uint32_t L = (t >> 2) + 1, size, offhi = 0, mask = 0;
  if (kind == 0) { size = 1 + L; if (L >= 61) { L = 0; size = 0; } }
  else if (kind == 1) { L = 4 + ((t >> 2) & 7); size = 2; offhi = (t & 0xe0) << 3; mask = 0xffu; }
  else if (kind == 2) { size = 3; mask = 0xffffu; }
  else { size = 5; mask = 0xffffffffu; }
  uint2 e;
  e.x = L | (size << 8) | (offhi << 16);
  e.y = mask;
  return e;
}
__device__ __forceinline__ void stage_tag_lut(uint2* lut) {
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = tag_entry(i);
  __syncthreads();
}

// kOverRead: the caller guarantees that the kDecOverRead bytes after in[n) are readable (they belong to a later
// chunk of the same buffer), so the stream loads need no clamping.
constexpr uint32_t kDecOverRead = SNB_DEC_PREFETCH + 64;
template <bool kOverRead>
__device__ __forceinline__ int32_t decompress_warp(const uint8_t* __restrict__ in, uint32_t n, uint8_t* out,
                                                   uint32_t cap, uint32_t* written, uint32_t lane, const uint2* lut,
                                                   bool bare = false) {
  *written = 0;
  // ---- u32 varint (stew/leb128 via snappy.nim:92-94); a bare block (one 64 KiB block of a longer raw stream,
  // found by the discovery pass below) has no length prefix: it must produce exactly `cap` bytes
  uint32_t ulen = bare ? cap : 0, br = 0;
  if (!bare) {
    bool done = false, bad = false;
    for (uint32_t i = 0; i < 5 && !done && !bad; i++) {
      if (i >= n) { bad = true; break; }
      uint32_t b = in[i];
      if (i == 4 && (b & 0xf0)) { bad = true; break; }
      ulen |= (b & 0x7f) << (7 * i);
      if (!(b & 0x80)) { done = true; br = i + 1; }
    }
    if (!done) return kInvalidInput;
  }
  if (cap < ulen) return kBufferTooSmall;                      // :96-97
  if (ulen == 0) return br != n ? kInvalidInput : kOk;         // :99-102
  if (ulen > kMaxDeviceMessage || n - br > kMaxDeviceMessage) return kUnsupported;

  const uint8_t* tags = in + br;
  const uint32_t tn = n - br;
  if (tn == 0) return kInvalidInput;   // a length but no tags: decodeAllTags writes nothing, snappy.nim:104-108 (the window below needs tn >= 1)
  const uint32_t lim = ulen;  // decisions are identical for any cap >= ulen
  uint32_t ip = 0, op = 0;
  // aligned word view of the tag stream: word 0 holds tags[0] at byte `tmis`
  const uint32_t tmis = (uint32_t)((uintptr_t)tags & 3);
  const uint32_t* const tw = (const uint32_t*)(tags - tmis);
#if SNB_DEC_WINDOW
  const uint32_t tw_last = (tn + tmis - 1) >> 2;   // tn >= 1 here
#endif
  // 256 stream bytes held in registers: lane l has words kb + l (Wa) and kb + 32 + l (Wb).  The row after them
  // (Wc) is already on its way: the shuffles below read Wa and Wb in every step, so a row that has only just been
  // requested when the window slides would stall the warp for a DRAM round trip every 128 stream bytes
  // (25 % of the stall samples in profiles/r2_decompress_kernel_1GiB.txt)
#if SNB_DEC_WINDOW
  uint32_t kb = 0;
  uint32_t Wa = __ldg(tw + umin32(lane, tw_last)), Wb = __ldg(tw + umin32(32 + lane, tw_last));
  uint32_t Wc = __ldg(tw + umin32(64 + lane, tw_last));
#endif
  // short elements (<= 16 bytes) of the previous step: sources already loaded, stores still to do.  They are
  // flushed after the next step's tag parse, so the source loads never stall the warp.
  uint32_t pL = 0, pdst = 0, pssh = 0, px0 = 0, px1 = 0, px2 = 0, px3 = 0, px4 = 0;
#if SNB_DEC_LUT32 && !defined(SNB_EMULATE)
  uint32_t lut_s;
  // through a shuffle, so that ptxas cannot rematerialise it in the loop (it sees through a plain mov, and does so when
  // registers are tight: four uniform instructions per step again in the 48-register build)
  lut_s = __shfl_sync(kFull, (uint32_t)__cvta_generic_to_shared(lut), 0);
#endif

#define SNB_FLUSH_SHORT()                                                                         \
  do {                                                                                            \
    if (pL) {                                                                                     \
      const uint32_t b_[4] = {__funnelshift_r(px0, px1, pssh), __funnelshift_r(px1, px2, pssh),   \
                              __funnelshift_r(px2, px3, pssh), __funnelshift_r(px3, px4, pssh)};  \
      uint8_t* d_ = out + pdst;                                                                   \
      const uint32_t m_ = 0xffffffffu >> (32u - pL);   /* bit tests compile to one R2P per 7 predicates */ \
      _Pragma("unroll") for (uint32_t i_ = 0; i_ < 8; i_++)                                       \
        if ((m_ >> i_) & 1u) d_[i_] = (uint8_t)(b_[i_ >> 2] >> (8 * (i_ & 3)));                   \
      if (pL > 8) {                                                                               \
        _Pragma("unroll") for (uint32_t i_ = 8; i_ < 16; i_++)                                    \
          if ((m_ >> i_) & 1u) d_[i_] = (uint8_t)(b_[i_ >> 2] >> (8 * (i_ & 3)));                 \
      }                                                                                           \
      pL = 0;                                                                                     \
    }                                                                                             \
  } while (0)

  while (ip < tn) {
#if SNB_DEC_WINDOW
    // ---- the 36 stream bytes of this step, assembled from the register window through shuffles
    const uint32_t i0 = ip + tmis;
    {
      const uint32_t k0 = i0 >> 2;
      if (k0 - kb >= 32) {
        if (k0 - kb < 64) { Wa = Wb; Wb = Wc; kb += 32; }
        else { kb = k0; Wa = __ldg(tw + umin32(kb + lane, tw_last)); Wb = __ldg(tw + umin32(kb + 32 + lane, tw_last)); }
        Wc = __ldg(tw + umin32(kb + 64 + lane, tw_last));
      }
    }
    const uint32_t j = ((i0 >> 2) - kb) + (((i0 & 3) + lane) >> 2), sh = (((i0 & 3) + lane) & 3) * 8;
    uint32_t w0, w1, w2;
    {
      const uint32_t a0 = __shfl_sync(kFull, Wa, (int)j), b0 = __shfl_sync(kFull, Wb, (int)j);
      const uint32_t a1 = __shfl_sync(kFull, Wa, (int)j + 1), b1 = __shfl_sync(kFull, Wb, (int)j + 1);
      const uint32_t a2 = __shfl_sync(kFull, Wa, (int)j + 2), b2 = __shfl_sync(kFull, Wb, (int)j + 2);
      w0 = j < 32 ? a0 : b0; w1 = j + 1 < 32 ? a1 : b1; w2 = j + 2 < 32 ? a2 : b2;
    }
#else
    // ---- the 36 stream bytes of this step straight from the L1 (three word loads per lane, 44 contiguous bytes per
    // warp): 7 instructions instead of the 17 of the register window, and 36 warps per SM cover the L1 latency
    const uint32_t i1 = ip + tmis + lane, k1 = i1 >> 2, sh = (i1 & 3) * 8;
    uint32_t w0, w1, w2;
    if constexpr (kOverRead && SNB_DEC_FASTFETCH) {
      // nothing needs clamping: one address register, three loads with immediate offsets and the prefetch -- 9
      // instructions instead of 21 (a per-step test for "far from the end" is if-converted: both forms would issue)
      const uint32_t* const q = tw + k1;
      w0 = __ldg(q); w1 = __ldg(q + 1); w2 = __ldg(q + 2);
#if SNB_DEC_PREFETCH
      prefetch_l1(q + SNB_DEC_PREFETCH / 4);
#endif
    } else {
      const uint32_t tw_last = (tn + tmis - 1) >> 2;   // tn >= 1 here
      w0 = __ldg(tw + umin32(k1, tw_last)); w1 = __ldg(tw + umin32(k1 + 1, tw_last)); w2 = __ldg(tw + umin32(k1 + 2, tw_last));
#if SNB_DEC_PREFETCH
      // the stream is read once: without this every fourth step waits for a new 128-byte line to come from DRAM
      prefetch_l1(tw + umin32(k1 + SNB_DEC_PREFETCH / 4, tw_last));
#endif
    }
#endif
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    const uint32_t pos = ip + lane;
    const bool inb = pos < tn;
    const uint32_t avail = inb ? tn - pos : 0;
    const uint32_t t = lo & 0xff, v4 = __funnelshift_r(lo, hi, 8);   // bytes past `avail` are never used when !bad
    // ---- decode "a tag starting at my byte" (decoder.nim:39-109): L output bytes, `size` stream bytes
#if SNB_DEC_LUT32 && !defined(SNB_EMULATE)
    uint2 te;   // 32-bit shared address kept in a register: the generic form rebuilds the window base in every step (S2UR + 3 uniform ops)
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(te.x), "=r"(te.y) : "r"(lut_s + t * 8u));
#else
    const uint2 te = lut[t];
#endif
    const bool lit = te.y == 0;
    const uint32_t kind = lit ? 0u : 1u;                         // literal / copy is all the rest of the step needs
    uint32_t L = te.x & 0xffu, size = (te.x >> 8) & 0xffu, hdr = 1;
    const uint32_t off = (te.x >> 16) | (v4 & te.y);             // :87-109
    bool bad = avail < size;                                     // :77-79 (input side)
    if (L == 0) {                                                // literal with a length field, :54-75
      const uint32_t ll = (t >> 2) - 59;
      bad = avail - 1 < 61;
      const uint32_t m = ll == 4 ? 0xffffffffu : ((1u << (8 * ll)) - 1u);
      L = (v4 & m) + 1;
      if (L == 0) bad = true;
      hdr = 1 + ll;
      if (!bad && avail - hdr < L) bad = true;
      size = hdr + L;
    }
    if (bad) size = 64;
    // J: the lane my tag's successor starts at; bit 5 set = beyond this step (sticky through the doubling below)
    uint32_t J = inb ? umin32(lane + size, 32) : 32;
    uint32_t R = 1u;
#pragma unroll
    for (int r = 0; r < 4; r++) {  // 2^4 = 16 >= tags in 32 bytes (every element is >= 2 bytes)
      R |= __reduce_or_sync(kFull, shl_clamp((R >> lane) & 1u, J));   // reached lanes mark their successor (none when J >= 32)
      if (r < 3) J = __shfl_sync(kFull, J, (int)J) | (J & 32u);       // shfl takes the source lane modulo 32
    }
    const bool is_tag = inb && ((R >> lane) & 1u);
    const uint32_t val = (is_tag && !bad) ? L : 0;
    uint32_t incl = val;
#if SNB_DEC_SHFLPRED && !defined(SNB_EMULATE)
    // shfl.up reports "source lane in range" as a predicate: one predicated add per round instead of a compare, a
    // select and an add (the five lane >= d predicates do not stay in predicate registers across the step)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
      asm volatile("{ .reg .pred p; .reg .u32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, t; }" : "+r"(incl) : "r"(d));
#else
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t u = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d) incl += u;
    }
#endif
    const uint32_t my_op = op + (incl - val);
    // branch-free (plain & and |: predicate logic instead of a divergent region); for a bad tag L and off are never used
    const bool e = is_tag & (bad | ((kind != 0) & ((off == 0) | (off > my_op)))    // :112-113
                             | (my_op > lim) | (lim - my_op < L));                // :77, :127
    // every vote / shuffle of the step follows at once (the malformed-stream exit after them): each divergent region
    // between two collectives costs a convergence test (UMOV + BRA.DIV)
    const uint32_t emask = __ballot_sync(kFull, e);
    const bool indep = kind == 0 || (my_op - off + L <= op);
    uint32_t lmask = __ballot_sync(kFull, is_tag && indep && L > 16);
    uint32_t dmask = __ballot_sync(kFull, is_tag && !indep);
    const uint32_t tagmask = __ballot_sync(kFull, is_tag);
    const uint32_t ip_next = __shfl_sync(kFull, pos + size, 31 - __clz((int)tagmask));
    const uint32_t op_next = op + __shfl_sync(kFull, incl, 31);
    if (emask) return kInvalidInput;

    // ---- the previous step's short elements reach memory now; after the barrier every byte before `op` is visible
    // (before the votes it measured slower: 5.59 -> 5.68 ms)
    SNB_FLUSH_SHORT();
    __syncwarp();

    // ---- copies.  An element is independent when its whole source lies before this step's
    // output (literals always are); independent short elements are copied by their own lane with
    // wide loads (stored one step later), long ones by the whole warp, dependent ones in order.
    if (is_tag && indep && L <= 16) {
      const uint8_t* s = (kind == 0 ? tags : out) + (kind == 0 ? pos + hdr : my_op - off);
      const uint32_t sa = (uint32_t)((uintptr_t)s & 3);
      const uint32_t* sw = (const uint32_t*)((uintptr_t)s & ~(uintptr_t)3);
      const uint32_t nw = (L + sa + 3) >> 2;                     // 1..5 words hold the L source bytes
      // words past the element keep whatever they held: the flush never stores a byte that comes from them
      const uint32_t wm = 0xffffffffu >> (32u - nw);   // bit tests: predicated loads instead of a ladder of branches
      px0 = ld_global32(sw);
      if (wm & 2u) px1 = ld_global32(sw + 1);
      if (wm & 4u) px2 = ld_global32(sw + 2);
      if (wm & 8u) px3 = ld_global32(sw + 3);
      if (wm & 16u) px4 = ld_global32(sw + 4);
      pssh = sa * 8; pL = L; pdst = my_op;
    }
    while (lmask) {
      int s = __ffs((int)lmask) - 1;
      lmask &= lmask - 1;
      uint32_t so = __shfl_sync(kFull, my_op, s), sl = __shfl_sync(kFull, L, s);
      uint32_t sk = __shfl_sync(kFull, kind, s);
      uint32_t sx = __shfl_sync(kFull, kind == 0 ? pos + hdr : off, s);
      const uint8_t* srcp = sk == 0 ? tags + sx : out + so - sx;
      warp_copy(out + so, srcp, sl, lane);
    }
    if (dmask) {  // a source inside this step's output: everything before it has to be in memory first
      SNB_FLUSH_SHORT();
      while (dmask) {
        __syncwarp();
        int s = __ffs((int)dmask) - 1;
        dmask &= dmask - 1;
        uint32_t so = __shfl_sync(kFull, my_op, s), sl = __shfl_sync(kFull, L, s), sx = __shfl_sync(kFull, off, s);
        const uint8_t* srcp = out + so - sx;
        if (sx >= sl) { for (uint32_t i = lane; i < sl; i += 32) out[so + i] = srcp[i]; }
        else { for (uint32_t i = lane; i < sl; i += 32) out[so + i] = srcp[i % sx]; }
      }
    }
    ip = ip_next;
    op = op_next;
  }
  SNB_FLUSH_SHORT();
#undef SNB_FLUSH_SHORT
  __syncwarp();
  if (op != ulen) return kInvalidInput;                        // snappy.nim:107-108
  *written = op;
  return kOk;
}

struct DecompressArgs {
  const uint8_t* in_base;
  uint8_t* out_base;
  const Chunk* chunks;
  int32_t* status;     // per chunk
  uint32_t* out_len;   // per chunk bytes produced
  uint32_t* counter;
  uint32_t count;
  uint32_t verify;     // 1: compare masked CRC of the produced bytes with chunk.crc
  const uint32_t* order;   // optional: job j decodes chunk order[j] (largest compressed chunks first, decode_order_kernel)
};

// Longest job first: a chunk's decode time goes with its number of tags, i.e. with its compressed size, and a launch of
// 16 384 chunks on 5 328 resident warps is three rounds and a fourth that is 7 % full -- the chunks that start last
// should be the cheap ones.  Counting sort of the chunk indices by in_len / 256, descending; one CTA.
__global__ void __launch_bounds__(1024) decode_order_kernel(const Chunk* chunks, uint32_t count, uint32_t* order) {
  __shared__ uint32_t hist[260];
  for (uint32_t i = threadIdx.x; i < 260; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) atomicAdd(&hist[257u - umin32(chunks[i].in_len >> 8, 257u)], 1u);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t k = 0; k < 258; k++) { const uint32_t c = hist[k]; hist[k] = run; run += c; }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) order[atomicAdd(&hist[257u - umin32(chunks[i].in_len >> 8, 257u)], 1u)] = i;
}

// Launch geometry, measured on B200 (1 GiB text).  Round 1 (244 instructions per step, issue slots 72 % busy): 36 warps per SM
// at 56 registers were as good as anything (256x4 154 GB/s, 128x9 154, 192x6 157, 128x10 148, 256x5 155).  After the round-2f
// instruction diet (200 per step, 67 % busy) occupancy pays again: 256x5 = 40 warps per SM at 48 registers (16 bytes of
// spill) 5.23 ms, 192x6 5.38, 160x8 5.30, 128x10 5.44 (ten CTAs' tables squeeze the L1), 320x4 5.42, 384x3 5.36, 640x2 5.54,
// 448x3 / 672x2 (40 registers, 124 bytes of spill) 6.06 / 6.13.
#ifndef SNB_DEC_MINB
#define SNB_DEC_MINB 5
#endif
#ifndef SNB_DEC_THREADS
#define SNB_DEC_THREADS 256
#endif
constexpr uint32_t kDecThreads = SNB_DEC_THREADS, kDecBlocksPerSm = SNB_DEC_MINB;
__global__ void __launch_bounds__(SNB_DEC_THREADS, SNB_DEC_MINB) decompress_kernel(DecompressArgs a, const CrcTables* g) {
  __shared__ CrcSmem t;
  __shared__ uint2 lut[256];
  if (a.verify) crc_stage_tables(&t, g);
  stage_tag_lut(lut);
  uint32_t lane = threadIdx.x & 31;
  uint64_t in_hi;   // a readable bound of the input buffer: the end of its first or last chunk, whichever is higher
  {
    const Chunk c0 = a.chunks[0], cl = a.chunks[a.count - 1];
    const uint64_t e0 = c0.in_off + c0.in_len, el = cl.in_off + cl.in_len;
    in_hi = e0 > el ? e0 : el;
  }
  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(a.counter, 1u);
    job = __shfl_sync(kFull, job, 0);
    if (job >= a.count) return;
    if (a.order) job = a.order[job];
    Chunk c = a.chunks[job];
    const uint8_t* in = a.in_base + c.in_off;
    uint8_t* out = a.out_base + c.out_off;
    int32_t st = kOk;
    uint32_t w = 0;
    const uint8_t* produced = out;
    if (c.kind == kKindCompressed || c.kind == kKindRawBlock) {
      // every chunk's payload lies in one buffer: what the highest chunk covers is readable, so all chunks but the
      // ones within kDecOverRead bytes of that end may over-read their stream
      const bool bare = c.kind == kKindRawBlock;
      if (c.in_off + c.in_len + kDecOverRead <= in_hi) st = decompress_warp<true>(in, c.in_len, out, c.out_cap, &w, lane, lut, bare);
      else st = decompress_warp<false>(in, c.in_len, out, c.out_cap, &w, lane, lut, bare);
    } else if (c.kind == kKindStored) {
      warp_copy(out, in, c.in_len, lane);
      w = c.in_len;
    } else {
      produced = in; w = c.in_len;
    }
    if (st == kOk && a.verify) {
      __syncwarp();
      uint32_t crc = crc_mask(crc32c_warp(t, g, produced, w, lane));
      if (crc != c.crc) st = kCrcMismatch;
    }
    if (lane == 0) { a.status[job] = st; a.out_len[job] = st == kOk ? w : 0; }
    __syncwarp();
  }
}

// Message batches: gather slot i (at i*stride, lens[i] bytes) to dst + offs[i], one WARP per message (messages are a
// few KiB: a CTA per message, as compact_kernel does for 64 KiB frames, would idle 7 of its 8 warps).
__global__ void __launch_bounds__(256) compact_warp_kernel(const uint8_t* scratch, uint64_t stride, const uint32_t* lens,
                                                           const uint64_t* offs, uint8_t* dst, uint32_t count) {
  const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (uint32_t job = blockIdx.x * wpb + (threadIdx.x >> 5); job < count; job += gridDim.x * wpb)
    warp_copy(dst + offs[job], scratch + (uint64_t)job * stride, lens[job], lane);
}

// Build Chunk descriptors for a raw batch from offset/length tables (device-side helper).
__global__ void make_chunks_kernel(const uint64_t* in_off, const uint32_t* in_len, const uint64_t* out_off,
                                   const uint32_t* out_cap, Chunk* chunks, uint32_t count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Chunk c;
  c.in_off = in_off[i]; c.out_off = out_off[i]; c.in_len = in_len[i]; c.out_cap = out_cap[i];
  c.crc = 0; c.kind = kKindCompressed;
  chunks[i] = c;
}

// Uncompressed lengths of a raw batch (codec.nim:129-138 semantics, u32 range): 0xffffffff on a bad varint.
__global__ void varint_lens_kernel(const uint8_t* base, const uint64_t* in_off, const uint32_t* in_len, uint32_t* lens,
                                   uint32_t count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint8_t* p = base + in_off[i];
  uint32_t n = in_len[i], v = 0;
  bool done = false;
  for (uint32_t k = 0; k < 5 && k < n; k++) {
    uint32_t b = p[k];
    if (k == 4 && (b & 0xf0)) break;
    v |= (b & 0x7f) << (7 * k);
    if (!(b & 0x80)) { done = true; break; }
  }
  lens[i] = done ? v : 0xffffffffu;
}

// ===================================================================== raw streams longer than one block
// The raw format (snappy.nim:27-64) is Leb128(len) followed by the tags of consecutive 64 KiB blocks, with nothing
// marking where a block's tags start.  Encoders of this family (the reference, C++ snappy) compress every block
// on its own, so once the block starts are known the blocks decode independently, one warp each.  Finding them is
// a walk over the tag chain; it is made parallel by speculation:
//
//   speculate  one warp per segment of the stream walks tags from the segment's first byte as if a tag started
//              there, marking every position it visits in a bitmap and noting where it leaves the segment.  A
//              tag that cannot be (it overruns the stream) makes the walk start over at the next byte; a chain
//              that meets a position visited BEFORE the last such restart runs into that impossible tag.
//   merge      a tag chain started at a wrong byte falls into step with the true one after a few tags.  Every
//              segment is walked from the previous segment's speculative exit -- its true entry if that
//              segment was itself on the chain -- until it meets a position its own speculative walk visited:
//              from there on the two coincide, so the speculative exit is the exit for that entry.
//   validate   one warp follows the true chain across segments, taking each segment's exit from the merge
//              pass when its true entry is the assumed one (nearly always) and re-walking it otherwise
//              (after a literal that jumped over whole segments).
//   count      every segment is walked again from its true entry, summing the output bytes of its tags.
//   emit       with the prefix sums of those counts, the same walk writes down the stream position of every
//              tag whose output offset is a multiple of 65536, and flags elements that straddle a boundary.
//
// The walk applies exactly the decoder's rules for tag sizes and truncation (decoder.nim:39-109).  Anything
// unusual -- a malformed chain, a total that is not the announced length, an element crossing a block boundary,
// a copy reaching back into an earlier block (caught by the block decoder) -- sends the whole stream to the
// single-warp decoder, whose verdict is the reference's.
constexpr uint32_t kWalkNone = 0xffffffffu, kWalkBad = 0xfffffffeu;
struct RawWalkArgs {
  const uint8_t* in;                  // the whole raw stream, length prefix included
  uint32_t n, br;                     // its size; bytes of the length prefix
  uint32_t seg_shift, nseg;           // segment s = stream bytes [s << seg_shift, (s + 1) << seg_shift)
  uint32_t* bitmap;                   // speculate: visited positions (zeroed, n / 32 + 2 words)
  uint32_t* exitp;                    // speculate: per segment, first visited position past the segment, or kWalkBad
  uint32_t* restart;                  // speculate: per segment, where the walk last had to start over (0: never)
  uint32_t* exit2;                    // merge: per segment, the exit when entered at the previous segment's exitp
  uint32_t* entry;                    // validate: per segment, true first tag in it, or kWalkNone
  unsigned long long* outbytes;       // count: per segment
  const unsigned long long* outbase;  // emit: exclusive prefix sums of outbytes
  uint32_t* block_start;              // emit: per block (preset to kWalkNone)
  uint32_t* flags;                    // [0]: 1 = chain malformed / does not end at n, 2 = not block structured
};

// kMode 0 speculate, 1 count, 2 emit, 3 / 4 walk from `start` until the speculative chain is met (validate's repair /
// the merge pass).  Uniform return value: the first chain position at or past the segment end, or kWalkBad.
template <int kMode>
__device__ __forceinline__ uint32_t raw_walk_warp(const RawWalkArgs& a, uint32_t s, uint32_t start, const uint2* lut,
                                                  uint32_t lane) {
  const uint32_t n = a.n;
  const uint32_t seg_hi = umin32(n, (s + 1) << a.seg_shift);
  const uint32_t mis = (uint32_t)((uintptr_t)a.in & 3);
  const uint32_t* const w = (const uint32_t*)(a.in - mis);
  const uint32_t wlast = (n + mis - 1) >> 2;
  uint32_t ip = start;
  uint32_t rs = 0;   // speculate: start of the current chain fragment
  unsigned long long o = kMode == 2 ? a.outbase[s] : 0ull;
  while (ip < seg_hi) {
    const uint32_t pos = ip + lane;
    const bool inb = pos < n;
    const uint32_t avail = inb ? n - pos : 0;
    const uint32_t i = pos + mis, k = i >> 2, sh = (i & 3) * 8;
    const uint32_t w0 = __ldg(w + umin32(k, wlast)), w1 = __ldg(w + umin32(k + 1, wlast)), w2 = __ldg(w + umin32(k + 2, wlast));
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    const uint32_t t = lo & 0xff, v4 = __funnelshift_r(lo, hi, 8);
    const uint2 te = lut[t];
    uint32_t L = te.x & 0xffu, size = (te.x >> 8) & 0xffu;
    bool bad = avail < size;
    if (L == 0) {  // literal with a length field, decoder.nim:54-75
      const uint32_t ll = (t >> 2) - 59;
      bad = avail - 1 < 61;
      const uint32_t m = ll == 4 ? 0xffffffffu : ((1u << (8 * ll)) - 1u);
      L = (v4 & m) + 1;
      if (L == 0) bad = true;
      if (!bad && avail - (1 + ll) < L) bad = true;
      size = 1 + ll + L;
    }
    if (bad) size = 64;
    uint32_t J = inb ? umin32(lane + size, 32) : 32;
    uint32_t R = 1u;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      R |= __reduce_or_sync(kFull, shl_clamp((R >> lane) & 1u, J));
      if (r < 3) J = __shfl_sync(kFull, J, (int)J) | (J & 32u);
    }
    const bool is_tag = inb && ((R >> lane) & 1u);
    const bool mine = is_tag && pos < seg_hi;         // tags at or past the segment end belong to the next segment
    const uint32_t badm = __ballot_sync(kFull, mine && bad);
    if (kMode != 0 && badm) return kWalkBad;
    if (kMode == 3 || kMode == 4) {  // met a position the speculative walk of this segment visited: same chain from here on
      const uint32_t m = __ballot_sync(kFull, mine);
      const uint32_t bw = __funnelshift_r(a.bitmap[ip >> 5], a.bitmap[(ip >> 5) + 1], ip & 31u);
      if (m & bw) return ip + (uint32_t)__ffs((int)(m & bw)) - 1u >= a.restart[s] ? a.exitp[s] : kWalkBad;
    }
    if (kMode == 0) {
      const uint32_t m = __ballot_sync(kFull, mine);
      if (lane == 0) {
        atomicOr(a.bitmap + (ip >> 5), m << (ip & 31u));
        if (ip & 31u) atomicOr(a.bitmap + (ip >> 5) + 1, m >> (32u - (ip & 31u)));
      }
      if (badm) {  // not a tag chain after all: any byte is as good a guess as the first one was
        ip = ip + (uint32_t)__ffs((int)badm);
        rs = ip;
        continue;
      }
    }
    if (kMode == 1 || kMode == 2) {
      const uint32_t val = mine ? L : 0u;
      uint32_t incl = val;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(kFull, incl, d);
        if (lane >= (uint32_t)d) incl += u;
      }
      if (kMode == 2 && mine) {
        const unsigned long long my_o = o + (incl - val);
        if ((my_o & 0xffffull) == 0) a.block_start[my_o >> 16] = pos;
        if ((my_o >> 16) != ((my_o + L - 1) >> 16)) atomicOr(a.flags, 2u);
      }
      o += __shfl_sync(kFull, incl, 31);
    }
    const uint32_t over = __ballot_sync(kFull, is_tag && pos >= seg_hi);
    if (over) { ip = ip + (uint32_t)__ffs((int)over) - 1u; break; }
    const uint32_t tagmask = __ballot_sync(kFull, is_tag);
    ip = __shfl_sync(kFull, pos + size, 31 - __clz((int)tagmask));   // the last tag of the step is not bad (checked above)
  }
  if (kMode == 1 && lane == 0) a.outbytes[s] = o;
  if (kMode == 0 && lane == 0) a.restart[s] = rs;
  return ip;
}

template <int kMode>
__global__ void __launch_bounds__(256) raw_walk_kernel(RawWalkArgs a) {
  __shared__ uint2 lut[256];
  stage_tag_lut(lut);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= a.nseg) return;
  uint32_t start;
  if (kMode == 0) start = s == 0 ? a.br : (s << a.seg_shift);
  else if (kMode == 4) {
    start = s == 0 ? a.br : a.exitp[s - 1];
    if (start >= umin32(a.n, (s + 1) << a.seg_shift)) start = kWalkNone;   // bad, or beyond this segment
  } else start = a.entry[s];
  if (start == kWalkNone) {  // the chain jumps over this segment (a long literal)
    if (kMode == 1 && lane == 0) a.outbytes[s] = 0;
    if (kMode == 4 && lane == 0) a.exit2[s] = kWalkNone;
    return;
  }
  const uint32_t e = raw_walk_warp<kMode>(a, s, start, lut, lane);
  if (kMode == 0 && lane == 0) a.exitp[s] = e;
  if (kMode == 4 && lane == 0) a.exit2[s] = e;
  if ((kMode == 1 || kMode == 2) && e == kWalkBad && lane == 0) atomicOr(a.flags, 1u);
}

// One warp: the true chain across segments (see above).  32 segments per turn: the merge pass's exits are fetched
// by all lanes at once, the dependent part runs on shuffles.
__global__ void __launch_bounds__(32) raw_validate_kernel(RawWalkArgs a) {
  __shared__ uint2 lut[256];
  stage_tag_lut(lut);
  const uint32_t lane = threadIdx.x;
  uint32_t X = a.br;   // the next true tag position
  bool dead = false;
  uint32_t st_n = 0, st_ns = 0, st_long = 0;
  for (uint32_t s0 = 0; s0 < a.nseg; s0 += 32) {
    const uint32_t sj = s0 + lane;
    const uint32_t e2 = sj < a.nseg ? a.exit2[sj] : kWalkNone;
    const uint32_t as = sj == 0 ? a.br : (sj < a.nseg ? a.exitp[sj - 1] : kWalkNone);   // the entry the merge pass assumed
    {  // the regular case, all 32 at once: every segment is entered where the merge pass assumed and left inside the next one
      uint32_t px = __shfl_up_sync(kFull, e2, 1);
      if (lane == 0) px = X;
      const bool in_range = sj < a.nseg;
      const bool reg = !in_range || (px == as && px < umin32(a.n, (sj + 1) << a.seg_shift) && e2 < kWalkBad);
      if (!dead && __all_sync(kFull, reg)) {
        if (in_range) a.entry[sj] = px;
        X = __shfl_sync(kFull, e2, (int)(umin32(a.nseg - s0, 32u) - 1u));
        continue;
      }
    }
    uint32_t my_entry = kWalkNone;
#pragma unroll 1   // one copy of the re-walk: a lone warp running through 32 unrolled copies lives on instruction-cache misses
    for (uint32_t j = 0; j < 32 && s0 + j < a.nseg; j++) {
      const uint32_t s = s0 + j;
      const uint32_t seg_hi = umin32(a.n, (s + 1) << a.seg_shift);
      if (dead || X >= seg_hi) continue;
      uint32_t ej = __shfl_sync(kFull, e2, (int)j);
      const uint32_t aj = __shfl_sync(kFull, as, (int)j);
      if (aj != X) {   // entered somewhere else than assumed (after a jump)
        const unsigned long long t0 = snb_now_ns();
        // The usual reason is a run of stored blocks: the chain lands on the tag of the next 64 KiB literal, which
        // leaves this segment again.  One (uniform) look at that tag settles it -- same rules as raw_walk_warp
        // (decoder.nim:54-75) -- without the 32-lane walk; everything else takes the walk.
        uint32_t fast = kWalkNone;
        {
          const uint32_t avail = a.n - X;   // X < seg_hi <= n
          // tag and the four bytes after it, requested together (one round trip); bytes past the stream read as 0
          const uint32_t t = __ldg(a.in + X);
          const uint32_t b1 = avail > 1 ? __ldg(a.in + X + 1) : 0u, b2 = avail > 2 ? __ldg(a.in + X + 2) : 0u;
          const uint32_t b3 = avail > 3 ? __ldg(a.in + X + 3) : 0u, b4 = avail > 4 ? __ldg(a.in + X + 4) : 0u;
          if ((t & 3u) == 0 && (t >> 2) >= 60 && avail - 1 >= 61) {
            const uint32_t ll = (t >> 2) - 59;
            const uint32_t v4 = b1 | (b2 << 8) | (b3 << 16) | (b4 << 24);
            const uint32_t L = (ll == 4 ? v4 : (v4 & ((1u << (8 * ll)) - 1u))) + 1;
            if (L != 0 && avail - (1 + ll) >= L && X + 1 + ll + L >= seg_hi) fast = X + 1 + ll + L;
          }
        }
        ej = fast != kWalkNone ? fast : raw_walk_warp<3>(a, s, X, lut, lane);
        {   // statistics (registers; written once at the end): re-walks, their time, the long ones
          const uint32_t dt = (uint32_t)(snb_now_ns() - t0);
          st_n++; st_ns += dt;
          if (dt > 100000u) st_long++;
        }
      }
      if (lane == j) my_entry = X;
      if (ej == kWalkBad) { dead = true; continue; }
      X = ej;
    }
    if (sj < a.nseg) a.entry[sj] = my_entry;
  }
  if (lane == 0) { a.flags[1] = st_n; a.flags[2] = st_ns; a.flags[3] = st_long; }
  if (lane == 0 && (dead || X != a.n)) atomicOr(a.flags, 1u);
}

}  // namespace snb

#include "snb_frameidx.cuh"
