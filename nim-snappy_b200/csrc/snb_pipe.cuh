// snb_pipe.cuh -- encodeBlock v6: one warp GROUP per 64 KiB chunk (included by snb_device.cuh).
//
// The serial part of encodeBlock (encoder.nim:184-383) is the walk of the match cursor; everything a
// lookup needs -- hash, table entry, candidate bytes, match length, copy tag -- is a pure function of
// (input, position, table entry) and can run ahead of the cursor.  v5 did both on one warp, so a chunk
// advanced one window per ~5 000 cycles and 3 256 chunks had to be in flight to fill the GPU: 300 MB of
// randomly accessed chunk history + tables against a 126 MB L2, i.e. the kernel was bound by random DRAM
// sectors (33 x the algorithmic traffic).  v6 splits the warp's job:
//
//   resolver warps (kResolvers per chunk) run AHEAD of the match cursor: window after window they hash
//            the 32 positions, read the (possibly not yet final) table entry T, fetch 12 candidate
//            bytes, and leave (h, T, match length, copy tag, literal byte) per position in a ring of
//            records in shared memory.  They never write the table.
//   walker warp    owns the table and the output.  Per window it takes the record, re-reads table[h]
//            (shared memory) and re-measures only the lanes whose entry changed since the resolver
//            looked (a record is trusted only if T is still the table's value, so the result never
//            depends on timing), then does exactly v5's walk / masks / commit, i.e. the reference's
//            literal-scan, emitLiteral / emitCopy and table-insert sequence.
//
// A chunk now advances a window in well under 1 000 cycles, so kGroups chunks per SM fill it: every hash
// table lives in shared memory and the chunk histories in flight (148 x 6 x 64 KiB = 57 MB) fit the L2.
// Only full 64 KiB chunks take this path (compile-time length: mask, ip_limit are constants); the short
// last chunk of a stream runs compress_block_v5 on the walker alone.
#pragma once

namespace snb {
namespace pipe {

#ifndef SNB_PIPE_GROUPS
#define SNB_PIPE_GROUPS 6
#endif
#ifndef SNB_PIPE_RESOLVERS
#define SNB_PIPE_RESOLVERS 2
#endif
#ifndef SNB_PIPE_RING
#define SNB_PIPE_RING 8
#endif
constexpr uint32_t kGroups = SNB_PIPE_GROUPS;        // chunks in flight per CTA, one 32 KiB table each
constexpr uint32_t kResolvers = SNB_PIPE_RESOLVERS;  // resolver warps per chunk (windows interleaved)
constexpr uint32_t kRing = SNB_PIPE_RING;            // record slots per chunk; slot = window % kRing, so a slot has ONE writer
static_assert(kRing % kResolvers == 0, "every ring slot must belong to one resolver");
constexpr uint32_t kWarpsPerGroup = 1 + kResolvers;
constexpr uint32_t kThreads = kGroups * kWarpsPerGroup * 32;
constexpr uint32_t kLen = kMaxBlock;
constexpr uint32_t kIpLimit = kLen - kInputMargin;
constexpr uint32_t kMask = kMaxTableEntries - 1;
constexpr uint32_t kWinMax = (kIpLimit - 33) / 32 + 1;   // windows 0 .. kWinMax-1 have b + 32 < ip_limit (the walker's window path)
constexpr uint32_t kStop = 0xffffu;                      // ctl window field: the job is over
constexpr uint32_t kExit = 0xffffffffu;                  // ctl: no more jobs

struct Group {            // shared memory, one per chunk group
  uint32_t ctl;           // (seq << 16) | next window the walker wants (kStop: job over); kExit: kernel done
  uint32_t ack;           // resolvers that have zeroed their share of the table, summed over jobs
  uint32_t in_lo, in_hi;  // the job's input
  uint32_t pad[12];
};
constexpr size_t kTableBytes = (size_t)kMaxTableEntries * 2;
constexpr size_t kRingBytes = (size_t)kRing * 32 * sizeof(uint2);
constexpr size_t kSmemBytes = kGroups * (kTableBytes + kRingBytes + sizeof(Group)) + sizeof(CrcSmem);
static_assert(kSmemBytes <= 227 * 1024, "pipe kernel: shared memory");

// Flags and records cross warps through shared memory WITHOUT fences: a fence (MEMBAR.*.CTA) also waits for the
// issuing warp's global loads and stores in flight -- the resolver's next-window prefetch, the walker's output bytes --
// i.e. a DRAM round trip per window.  No ordering between different shared-memory words is assumed instead:
//   * ctl is a single word that carries no data dependency (a cursor position);
//   * a record is ONE 8-byte word per lane, written by one STS.64 and read by one LDS.64, and carries its own stamp
//     (window index + job sequence): a lane either sees the whole record of the window it asked for or retries.
// The two job-level hand-overs (input pointer, table zeroed) happen once per 64 KiB chunk and do use fences.
#ifdef SNB_EMULATE
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) { return *(volatile const uint32_t*)p; }
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) { *(volatile uint32_t*)p = v; }
__device__ __forceinline__ uint2 ld_record(const uint2* p) { uint2 r; r.x = ((volatile const uint2*)p)->x; r.y = ((volatile const uint2*)p)->y; return r; }
__device__ __forceinline__ void st_record(uint2* p, uint2 r) { ((volatile uint2*)p)->x = r.x; ((volatile uint2*)p)->y = r.y; }
__device__ __forceinline__ void spin_pause() { emu::yield(); }
__device__ __forceinline__ void spin_sleep(unsigned) { emu::yield(); }
#else
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 ld_record(const uint2* p) {
  uint2 r;
  asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return r;
}
__device__ __forceinline__ void st_record(uint2* p, uint2 r) {
  asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(r.x), "r"(r.y) : "memory");
}
__device__ __forceinline__ void spin_pause() {}
__device__ __forceinline__ void spin_sleep(unsigned ns) { __nanosleep(ns); }
#endif

// Match of the 12 bytes a* at position p against the 12 bytes c* at T: length 0..12 and the first byte of the
// copy element a 4..11-byte match emits (encoder.nim:113-125); the second byte is off & 0xff in both forms.
__device__ __forceinline__ uint32_t match12(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t c0, uint32_t c1, uint32_t c2) {
  const uint32_t x0 = a0 ^ c0, x1 = a1 ^ c1, x2 = a2 ^ c2;
  const uint32_t xs = x0 ? x0 : (x1 ? x1 : x2);
  const uint32_t base = x0 ? 0u : (x1 ? 4u : 8u);
  return xs ? base + first_diff_byte(xs) : 12u;
}
__device__ __forceinline__ uint32_t copy_tag(uint32_t off, uint32_t ml) {
  return off >= 2048 ? (((ml - 1) << 2) | 2u) : (((off >> 8) << 5) | ((ml - 4) << 2) | 1u);
}
// record: x = h | stamp[12..13] << 14 | T << 16;  y = ml | tag << 4 | literal byte << 12 | stamp[0..11] << 20
// stamp = (seq % 7 + 1) << 11 | window: never 0 (a zeroed slot is never valid), differs between consecutive jobs.
__device__ __forceinline__ uint32_t stamp_base(uint32_t seq) { return (seq % 7u + 1u) << 11; }   // once per job
__device__ __forceinline__ uint2 make_record(uint32_t h, uint32_t T, uint32_t ml, uint32_t tag, uint32_t a0, uint32_t stamp) {
  uint2 r;
  r.x = h | ((stamp >> 12) << 14) | (T << 16);
  r.y = ml | ((tag & 0xffu) << 4) | ((a0 & 0xffu) << 12) | (stamp << 20);
  return r;
}
__device__ __forceinline__ bool record_is(uint2 r, uint32_t stamp) {
  return (r.y >> 20) == (stamp & 0xfffu) && ((r.x >> 14) & 3u) == (stamp >> 12);
}

// ------------------------------------------------------------------------------------------ resolver
__device__ __forceinline__ void resolver_main(Group* g, uint16_t* table, uint2* ring, uint32_t r, uint32_t lane) {
  const uint32_t wlane = umin32(lane, 11u);
  uint32_t seq = 0;
  for (;;) {
    seq = (seq + 1) & 0xffffu;
    uint32_t c;
    for (;;) {
      c = ld_flag(&g->ctl);
      if (c == kExit || (c >> 16) == seq) break;
      spin_sleep(200);
    }
    if (c == kExit) return;
    __threadfence_block();   // the job's input pointer was written before ctl
    InView v;
    {
      const uint8_t* in = (const uint8_t*)(((unsigned long long)g->in_hi << 32) | g->in_lo);
      v.mis = (uint32_t)((uintptr_t)in & 3);
      v.w = (const uint32_t*)(in - v.mis);
      v.last = (kLen + v.mis - 1) >> 2;
    }
    {  // zero this warp's ring slots (no record of an earlier job survives) and its share of the table (encoder.nim:237-238)
      uint2 z; z.x = 0; z.y = 0;
      for (uint32_t sl = r; sl < kRing; sl += kResolvers) st_record(&ring[sl * 32 + lane], z);
      uint32_t* t32 = (uint32_t*)table + r * (kMaxTableEntries / 2 / kResolvers);
      for (uint32_t i = lane; i < kMaxTableEntries / 2 / kResolvers; i += 32) t32[i] = 0;
      __syncwarp();
      if (lane == 0) { __threadfence_block(); atomicAdd(&g->ack, 1u); }
    }
    const uint32_t own_q = (lane + v.mis) >> 2, own_s = ((lane + v.mis) & 3) * 8;
    const uint32_t sbase = stamp_base(seq);
    uint32_t win = r;
    uint32_t Wn = 0, Wn_win = 0xffffffffu;
    for (;;) {
      c = ld_flag(&g->ctl);
      if ((c >> 16) != seq) break;
      const uint32_t want = c & 0xffffu;
      if (want == kStop) break;
      if (win < want) win = want + (r + kResolvers - want % kResolvers) % kResolvers;   // the cursor jumped: skip
      if (win >= kWinMax || win >= want + kRing) { spin_sleep(win >= kWinMax ? 1000 : 100); continue; }   // ring full: do not eat the walkers' issue slots
      const uint32_t b = win << 5, p = b + lane;
      const uint32_t W = (Wn_win == win) ? Wn : __ldg(v.w + (b >> 2) + wlane);
      Wn_win = win + kResolvers;
      Wn = __ldg(v.w + umin32((Wn_win << 3) + wlane, v.last));
      const uint32_t w0 = __shfl_sync(kFull, W, (int)own_q), w1 = __shfl_sync(kFull, W, (int)own_q + 1);
      const uint32_t w2 = __shfl_sync(kFull, W, (int)own_q + 2), w3 = __shfl_sync(kFull, W, (int)own_q + 3);
      const uint32_t a0 = __funnelshift_r(w0, w1, own_s), a1 = __funnelshift_r(w1, w2, own_s), a2 = __funnelshift_r(w2, w3, own_s);
      const uint32_t h = hash14(a0, kMask);
      // Not necessarily the entry the walker will see (it may still commit earlier windows, or another resolver may
      // still be zeroing): clamp so that the candidate read stays inside the chunk; the walker checks T.
      const uint32_t T = umin32(*(const volatile uint16_t*)(table + h), p);
      uint32_t c0, c1, c2;
      in_ld12u(v, T, c0, c1, c2);
      const uint32_t ml = match12(a0, a1, a2, c0, c1, c2);
      st_record(&ring[(win % kRing) * 32 + lane], make_record(h, T, ml, copy_tag(p - T, ml), a0, sbase | win));
      win += kResolvers;
    }
  }
}

// ------------------------------------------------------------------------------------------ walker
// compress_block_v5 with the resolve stage replaced by "take the record, check T against the table".
// Everything after that -- walk, masks, stale rule, commit, the serial steps -- is v5's, so the emitted bytes and the
// table evolution are the reference's (encoder.nim:184-383).  len == kLen.
__device__ __forceinline__ uint32_t compress_block_v6(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                      uint16_t* table, uint32_t abort_thr, uint32_t lane, Group* g,
                                                      const uint2* ring, uint32_t seq) {
  InView v;
  v.mis = (uint32_t)((uintptr_t)in & 3);
  v.w = (const uint32_t*)(in - v.mis);
  v.last = (kLen + v.mis - 1) >> 2;
  constexpr uint32_t len = kLen, mask = kMask, ip_limit = kIpLimit;
  const uint32_t seqhi = seq << 16, sbase = stamp_base(seq);
  const uint32_t below = (1u << lane) - 1u;
  const uint32_t notbelow = ~below;
  uint32_t op = 0;
  uint32_t ip = 0;
  uint32_t next_emit = 0;
  uint32_t pp = 1;
  uint32_t k = 0;
  bool chain = false, force_slow = false;
  bool m1_done = false;
  uint32_t want_pub = 0;   // the window ctl currently announces (the job starts with 0)

  // the table is zero once every resolver has acknowledged this job
  while (ld_flag(&g->ack) != seq * kResolvers) spin_pause();
  __threadfence_block();

#define SNB_FINISH(from)                                                          \
  do {                                                                            \
    uint32_t f_ = (from);                                                         \
    if (f_ < len) {                                                               \
      uint32_t n_ = len - f_;                                                     \
      if (op + literal_size(n_) > abort_thr) return kAborted;                     \
      op += emit_literal_v2(out + op, v, f_, n_, lane);                           \
    }                                                                             \
    return op;                                                                    \
  } while (0)

  for (;;) {
    __syncwarp();
    const uint32_t b = (chain ? ip : pp) & ~31u;

    SNB_STAT(8);
    if (force_slow || b + 32 >= ip_limit || (!chain && k >= 32)) {
      force_slow = false;
      m1_done = false;
      if (chain) {
        SNB_STAT(1);
        const uint32_t d1 = in_ld4(v, ip - 1), d = in_ld4(v, ip);
        const uint32_t h1 = hash14(d1, mask), hh = hash14(d, mask);
        const uint32_t tv = table[hh];
        const uint32_t cand = (hh == h1) ? ip - 1 : tv;
        __syncwarp();
        if (lane == 0) { table[h1] = (uint16_t)(ip - 1); table[hh] = (uint16_t)ip; }
        const uint32_t m = common_prefix_v2(v, len, cand, ip, lane);
        if (m < 4) { chain = false; next_emit = ip; pp = ip + 1; k = 0; continue; }
        op += emit_copy_v2(out + op, ip - cand, m, lane);
        ip += m;
        if (op > abort_thr) return kAborted;
        if (ip > ip_limit) SNB_FINISH(ip);
        continue;
      }
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

    // ---- window [b, b+32): tell the resolvers where the cursor is, then take the window's record
    const uint32_t win = b >> 5;
    if (win != want_pub) {
      want_pub = win;
      if (lane == 0) st_flag(&g->ctl, seqhi | win);
    }
    const uint32_t p = b + lane;
    if (chain && ip == b && !m1_done) {
      const uint32_t d1 = in_ld4(v, ip - 1);
      if (lane == 0) table[hash14(d1, mask)] = (uint16_t)(ip - 1);
      __syncwarp();
    }
    m1_done = false;

    SNB_STAT(0);
    uint2 rec;
    {
      const uint2* rp = ring + (win % kRing) * 32 + lane;
      const uint32_t stamp = sbase | win;
      for (;;) {   // every lane needs its own record of THIS window (a warp-wide store may be seen half done)
        rec = ld_record(rp);
        if (__all_sync(kFull, record_is(rec, stamp))) break;
        SNB_STAT(10);
        spin_pause();
      }
    }
    const uint32_t h = rec.x & kMask;
    const uint32_t T = table[h];
    uint32_t ml = rec.y & 15u, tag = (rec.y >> 4) & 0xffu;
    const uint32_t lit = (rec.y >> 12) & 0xffu;
    const uint32_t off = p - T;
    if (T != (rec.x >> 16)) {   // the entry moved after the resolver looked: re-measure against the entry that counts
      SNB_STAT(11);
      uint32_t a0, a1, a2, c0, c1, c2;
      in_ld12u(v, p, a0, a1, a2);
      in_ld12u(v, T, c0, c1, c2);
      ml = match12(a0, a1, a2, c0, c1, c2);
      tag = copy_tag(off, ml);
    }
    const bool c3 = off >= 2048;
    const uint32_t grp = __match_any_sync(kFull, h);
    const uint32_t eg = grp & below;
    const uint32_t HB = __ballot_sync(kFull, ml >= 4);
    const uint32_t C3 = __ballot_sync(kFull, c3);
    const uint32_t qb = HB & notbelow;
    const uint32_t evt = qb ? (uint32_t)__ffs((int)qb) - 1u : 32u;
    const uint32_t eml = __shfl_sync(kFull, ml, (int)(evt & 31u));
    const uint32_t J = evt >= 32 ? 128u : (eml >= 12 ? 64u + evt : evt + eml);
    const uint32_t lkmask = notbelow & ((2u << umin32(evt, 31u)) - 1u);

    uint32_t ins = 0;
    uint32_t c = 0;
    uint32_t noprev = 0;
    if (!chain) {
      const uint32_t l0 = pp - b;
      const uint32_t bend = umin32(32u, l0 + (32u - k));
      const uint32_t range = (0xffffffffu << l0) & (bend >= 32 ? 0xffffffffu : ((1u << bend) - 1u));
      const uint32_t qh = HB & range;
      const uint32_t q0 = qh ? (uint32_t)__ffs((int)qh) - 1u : 32u;
      const uint32_t run = range & (q0 >= 32 ? 0xffffffffu : ((2u << q0) - 1u));
      const uint32_t stale = __ballot_sync(kFull, (eg & run) != 0) & run;
      if (stale) { SNB_STAT(5); force_slow = true; continue; }
      if (q0 >= 32) {
        SNB_STAT(6);
        __syncwarp();
        if (((run >> lane) & 1u) && ((grp & run) >> lane) == 1u) table[h] = (uint16_t)p;
        k += (uint32_t)__popc(run);
        pp = b + bend;
        continue;
      }
      {
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

    bool done = false;
    for (;;) {
      SNB_STAT(3);
      uint32_t A = 0, cc = c;
      do { A |= 1u << cc; cc = __shfl_sync(kFull, J, (int)cc); } while (cc < 32);
      bool onp = (A >> lane) & 1u;
      uint32_t LK = __reduce_or_sync(kFull, onp ? lkmask : 0u);
      uint32_t insb = LK | ((A & ~noprev) >> 1);
      const uint32_t stale = __ballot_sync(kFull, ((LK >> lane) & 1u) && (eg & (ins | insb)) != 0);
      uint32_t slow_at = 32;
      if (stale) {
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
      const uint32_t Ev = LK & HB;
      uint32_t EvT = Ev, tail = 0;
      if (cc >= 64 && cc < 96) EvT &= ~(1u << (cc - 64));
      if (cc == 128) tail = 0xffffffffu << (31u - (uint32_t)__clz((int)A));
      const uint32_t LitB = LK & ~Ev & ~tail;
      const uint32_t LS = A & ~HB & ~tail;
      const uint32_t E3 = EvT & C3;
      {
        const uint32_t offl = (uint32_t)(__popc(LitB & below) + __popc(LS & below) + 2 * __popc(EvT & below) + __popc(E3 & below));
        uint8_t* o = out + op + offl;
        if ((LS >> lane) & 1u) { o[0] = (uint8_t)((evt - lane - 1u) << 2); o += 1; }
        if ((LitB >> lane) & 1u) o[0] = (uint8_t)lit;
        if ((EvT >> lane) & 1u) {
          o[0] = (uint8_t)tag; o[1] = (uint8_t)off;
          if (c3) o[2] = (uint8_t)(off >> 8);
        }
      }
      op += (uint32_t)(__popc(LitB) + __popc(LS) + 2 * __popc(EvT) + __popc(E3));
      if (op > abort_thr) return kAborted;
      if (slow_at < 32) { ip = b + slow_at; force_slow = true; break; }
      if (cc < 64) { ip = b + cc; break; }
      if (cc == 128) {
        const uint32_t cl = 31u - (uint32_t)__clz((int)A);
        chain = false; next_emit = b + cl; pp = b + 32; k = 31 - cl;
        break;
      }
      const uint32_t e = cc - 64;
      SNB_STAT(4);
      const uint32_t cand = __shfl_sync(kFull, T, (int)e);
      const uint32_t m = 12 + common_prefix_v2(v, len, cand + 12, b + e + 12, lane);
      op += emit_copy_v2(out + op, b + e - cand, m, lane);
      ip = b + e + m;
      if (op > abort_thr) return kAborted;
      if (ip > ip_limit) { done = true; break; }
      c = ip - b;
      if (c >= 32) break;
    }
    if (chain && !force_slow && ip > ip_limit) done = true;
    bool pend_insert = false;
    if (chain && !force_slow && !done && (ip & 31u) == 0) {
      if (ip == b + 32) ins |= 1u << 31; else pend_insert = true;
      m1_done = true;
    }
    __syncwarp();
    if (((ins >> lane) & 1u) && ((grp & ins) >> lane) == 1u) table[h] = (uint16_t)p;
    if (done) SNB_FINISH(ip);
    if (pend_insert) {
      __syncwarp();
      const uint32_t d1 = in_ld4(v, ip - 1);
      if (lane == 0) table[hash14(d1, mask)] = (uint16_t)(ip - 1);
    }
  }
#undef SNB_FINISH
}

}  // namespace pipe

// One job of the streaming kernels: encodeFrame (encoder.nim:385-426) or a bare encodeBlock, `enc(in, len, out, thr)`
// being the block encoder.  Same layout decisions as compress_job.
template <class Enc>
__device__ __forceinline__ uint32_t compress_job_with(const CompressArgs& a, uint32_t job, uint32_t lane, uint32_t crc, Enc enc) {
  const uint64_t ioff = (uint64_t)job * kMaxBlock;
  const uint32_t len = (uint32_t)((a.total_len - ioff) < kMaxBlock ? (a.total_len - ioff) : kMaxBlock);
  const uint8_t* in = a.in_base + ioff;
  uint8_t* out = a.out_base + (uint64_t)job * a.out_stride;
  const bool frame = a.mode == kModeFrame;
  const uint32_t hl = a.mode == kModeRawBlock ? 0u : varint32_len(len);
  const uint32_t pre = frame ? 8u + hl : hl;
  const uint32_t thr = frame ? len - len / 8 : kNoAbort;
  uint32_t bl = frame ? kAborted : 0u;
  if (frame ? len >= kMinNonLiteral : len > 0) bl = enc(in, len, out + pre, thr);
  if (!frame) return hl + bl;
  uint32_t fl;
  if (bl != kAborted && bl <= thr) {
    put_varint32(out + 8, len, lane);
    fl = hl + bl + 4;
    if (lane == 0) out[0] = 0x00;
  } else {
    fl = len + 4;
    if (lane == 0) out[0] = 0x01;
    if (!a.defer_stored) warp_copy(out + 8, in, len, lane);
  }
  if (lane == 0) {
    out[1] = (uint8_t)fl; out[2] = (uint8_t)(fl >> 8); out[3] = (uint8_t)(fl >> 16);
    out[4] = (uint8_t)crc; out[5] = (uint8_t)(crc >> 8); out[6] = (uint8_t)(crc >> 16); out[7] = (uint8_t)(crc >> 24);
  }
  return fl + 4;
}

// Streaming compress, v6 geometry: kGroups chunk groups per CTA, each one walker + kResolvers resolver warps.
// Arrival wait, CRC, in-order packing and group flags are compress_stream_kernel's, done by the walker.
__global__ void __launch_bounds__(pipe::kThreads, 1) compress_pipe_kernel(StreamArgs a) {
  SNB_DYN_SMEM(smem);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t grp = warp / pipe::kWarpsPerGroup, role = warp % pipe::kWarpsPerGroup;
  uint16_t* table = (uint16_t*)(smem + (size_t)grp * pipe::kTableBytes);
  uint2* ring = (uint2*)(smem + pipe::kGroups * pipe::kTableBytes + (size_t)grp * pipe::kRingBytes);
  pipe::Group* g = (pipe::Group*)(smem + pipe::kGroups * (pipe::kTableBytes + pipe::kRingBytes)) + grp;
  CrcSmem* crc_t = (CrcSmem*)(smem + pipe::kGroups * (pipe::kTableBytes + pipe::kRingBytes + sizeof(pipe::Group)));
  const bool frame = a.c.mode == kModeFrame;
  if (role == 0 && lane == 0) { g->ctl = 0; g->ack = 0; }   // seq 0 is never used by a job
  if (frame) crc_stage_tables(crc_t, a.crc_tables);   // ends with __syncthreads()
  else __syncthreads();
  if (role != 0) { pipe::resolver_main(g, table, ring, role - 1, lane); return; }

  uint32_t seq = 0;
  for (;;) {
    uint32_t job = 0;
    if (lane == 0) job = atomicAdd(a.c.counter, 1u);
    job = __shfl_sync(kFull, job, 0);
    if (job >= a.c.count) break;
    if (a.arrived) {
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
    const unsigned long long ioff = (unsigned long long)job * kMaxBlock;
    const uint32_t len = (uint32_t)((a.c.total_len - ioff) < kMaxBlock ? (a.c.total_len - ioff) : kMaxBlock);
    const bool piped = len == pipe::kLen;
    if (piped) {   // the resolvers start (zero the table, fill the ring) while this warp does the CRC
      seq = (seq + 1) & 0xffffu;
      if (lane == 0) {
        const unsigned long long ip64 = (unsigned long long)(uintptr_t)(a.c.in_base + ioff);
        g->in_lo = (uint32_t)ip64; g->in_hi = (uint32_t)(ip64 >> 32);
        __threadfence_block();
        pipe::st_flag(&g->ctl, seq << 16);
      }
    }
    uint32_t crc = 0;
    if (frame) crc = crc_mask(crc32c_warp(*crc_t, a.crc_tables, a.c.in_base + ioff, len, lane));
    const uint32_t total = compress_job_with(a.c, job, lane, crc, [&](const uint8_t* in, uint32_t n, uint8_t* out, uint32_t thr) {
      if (piped) return pipe::compress_block_v6(in, out, table, thr, lane, g, ring, seq);
      return compress_block_v5(in, n, out, table, thr, lane);
    });
    __syncwarp();
    if (lane == 0) {
      if (piped) pipe::st_flag(&g->ctl, (seq << 16) | pipe::kStop);
      st_release_gpu(a.c.out_len + job, total);
      __threadfence();
    }
    __syncwarp();
    stream_place(a, lane);
    __syncwarp();
  }
  if (lane == 0) pipe::st_flag(&g->ctl, pipe::kExit);
}

}  // namespace snb
