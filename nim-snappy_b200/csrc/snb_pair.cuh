// snb_pair.cuh -- encodeBlock v7: a PAIR of warps per 64 KiB chunk, leapfrogging over its 32-position windows
// (included by snb_device.cuh).
//
// compress_block_v5 runs the whole window pass -- resolve (own bytes, hash, table entry, candidate bytes, match
// length), walk, emit, commit -- as one dependent chain on one warp: about 2 800 cycles per window unloaded
// (scripts/gpu_cyc.py), so 22 chunks per SM had to be in flight and their tables + histories (213 MB) thrashed the
// 126 MB L2 (33 x the algorithmic DRAM traffic).  v7 keeps v5's pass, and with it the reference's exact sequence of
// lookups, inserts, emitLiteral / emitCopy decisions (encoder.nim:184-383), but takes everything that does not
// depend on the match cursor off the critical path:
//
//   turn      the two warps of a pair take the window passes of a chunk alternately.  The cursor state (ip | pp,
//             next_emit, k, op, flags) is handed over through one 16-byte record in shared memory.
//   pre       while its partner walks window b, a warp resolves the window it will most likely get next: own bytes,
//             hashes, the (not yet final) table entries, the candidate bytes from L2, match lengths.  When its turn
//             comes it re-reads table[h] -- now final, the partner has committed -- and re-measures only the lanes
//             whose entry moved, so the result never depends on timing.
//   commit    v5 found the lanes that share a hash with match.any: 200-390 cycles of latency on sm_100a
//             (scripts/ubench/lat.cu), spent in every pass to catch a conflict that happens in about 4 % of them.  v7
//             walks optimistically (no groups), lets every inserted lane write the table, reads the entries back and,
//             if two inserted lanes did share a hash, restores the entries and redoes the pass on v5's exact path.
//
// A chunk advances a window per hand-over-to-hand-over chain (table re-read, ballots, walk, commit), so a few pairs
// per SM fill it: every hash table lives in shared memory (north_star layout) and the chunk histories in flight
// (148 x 6..7 x 64 KiB) fit the L2.  Only full 64 KiB chunks take this path (compile-time length); the short last
// chunk of a stream runs compress_block_v5 on the pair's first warp.
#pragma once

namespace snb {
namespace pair {

#ifndef SNB_RING
#define SNB_RING 2          // warps per chunk
#endif
#ifndef SNB_PAIR_MAXT
#define SNB_PAIR_MAXT 448
#endif
constexpr uint32_t kRing = SNB_RING;
static_assert(kRing >= 2 && kRing <= 8, "ring size");
constexpr uint32_t kMaxThreads = SNB_PAIR_MAXT;
constexpr uint32_t kMaxPairs = kMaxThreads / (32 * kRing);
constexpr uint32_t kLen = kMaxBlock;
constexpr uint32_t kIpLimit = kLen - kInputMargin;
constexpr uint32_t kMask = kMaxTableEntries - 1;
constexpr uint32_t kNoWindow = 0xffffffffu;
constexpr uint32_t kJobExit = 0xffffffffu;
// hand-over record: x = turn, y = cursor (chain ? ip : pp) | flags, z = next_emit | k << 20, w = op (result once finished)
constexpr uint32_t kFlagChain = 1u << 20, kFlagSlow = 1u << 21, kFlagM1 = 1u << 22, kFlagFin = 1u << 23;

struct Box {            // shared memory, one per warp group
  uint4 rec;            // the cursor for turn x: y = cursor | flags | Fin hops << 24 | nonce << 28, z = next_emit | k << 20
                        // (Fin: z = the turn that finished, w = the result).  One STS.128 / LDS.128: seen whole or not at all
  uint4 com;            // turn x has committed: the table holds every insert up to it; y = nonce of the valid rec(x + 1), z = op after it
  uint32_t job;         // leader -> partners: the job (kJobExit: none left)
  uint32_t job_seq;     // leader -> partners: counts jobs
  uint32_t ack;         // partners -> leader: += 1 per partner whose share of the table is zero
  uint32_t fin;         // partners -> leader: += 1 per partner whose output bytes are fenced
  uint32_t pad[4];
};
static_assert(sizeof(Box) == 64, "Box layout");

// Flags and the record cross warps through shared memory without fences (a MEMBAR.CTA also waits for the warp's
// global loads in flight, i.e. an L2 round trip per window): a warp's shared-memory accesses are performed in
// program order, the record is one 16-byte word, and nothing else is assumed.
#ifdef SNB_EMULATE
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) { return *(volatile const uint32_t*)p; }
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) { *(volatile uint32_t*)p = v; }
__device__ __forceinline__ uint4 ld_rec(const uint4* p) {
  uint4 r; const volatile uint4* q = (const volatile uint4*)p; r.x = q->x; r.y = q->y; r.z = q->z; r.w = q->w; return r;
}
__device__ __forceinline__ void st_rec(uint4* p, uint4 r) { volatile uint4* q = (volatile uint4*)p; q->y = r.y; q->z = r.z; q->w = r.w; q->x = r.x; }
__device__ __forceinline__ void spin_pause() { emu::yield(); }
#else
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_rec(const uint4* p) {
  uint4 r;
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return r;
}
__device__ __forceinline__ void st_rec(uint4* p, uint4 r) {
  asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"((uint32_t)__cvta_generic_to_shared(p)),
               "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}
#ifndef SNB_SPIN_NS
#define SNB_SPIN_NS 0
#endif
__device__ __forceinline__ void spin_pause() { if (SNB_SPIN_NS) __nanosleep(SNB_SPIN_NS); }
#endif
__device__ __forceinline__ uint32_t ld_tab(const uint16_t* p) { return *(const volatile uint16_t*)p; }

// Match of the 12 bytes a* at position p against the 12 bytes c* at T: 0..12.
__device__ __forceinline__ uint32_t match12(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t c0, uint32_t c1, uint32_t c2) {
  const uint32_t x0 = a0 ^ c0, x1 = a1 ^ c1, x2 = a2 ^ c2;
  const uint32_t xs = x0 ? x0 : (x1 ? x1 : x2);
  const uint32_t base = x0 ? 0u : (x1 ? 4u : 8u);
  return xs ? base + first_diff_byte(xs) : 12u;
}


// The group kernel is large and its warps sit at unrelated program counters, so what decides its speed is how much
// code the hot turn touches (ncu: 5.7 of 15 cycles per instruction were instruction-fetch stalls with every helper
// inlined at each of its call sites).  Everything that is not part of the common turn is called, not inlined.
#ifdef SNB_EMULATE
#define SNB_COLD
#else
#define SNB_COLD __noinline__
#endif
__device__ SNB_COLD uint32_t emit_literal_cold(uint8_t* o, const uint32_t* vw, uint32_t vmis, uint32_t vlast, uint32_t s, uint32_t n, uint32_t lane) {
  InView v; v.w = vw; v.mis = vmis; v.last = vlast;
  return emit_literal_v2(o, v, s, n, lane);
}
// literals of up to 32 bytes (one tag byte when n <= 60) inline, the rest through the call
__device__ __forceinline__ uint32_t emit_literal_v7(uint8_t* o, const InView& v, uint32_t s, uint32_t n, uint32_t lane) {
  if (n - 1u < 32u) {
    if (lane == 0) o[0] = (uint8_t)((n - 1) << 2);
    if (lane < n) o[1 + lane] = (uint8_t)in_byte(v, s + lane);
    return n + 1;
  }
  return emit_literal_cold(o, v.w, v.mis, v.last, s, n, lane);
}
__device__ SNB_COLD uint32_t common_prefix_cold(const uint32_t* vw, uint32_t vmis, uint32_t vlast, uint32_t len, uint32_t s1, uint32_t s2, uint32_t lane) {
  InView v; v.w = vw; v.mis = vmis; v.last = vlast;
  return common_prefix_v2(v, len, s1, s2, lane);
}
__device__ SNB_COLD uint32_t emit_copy_cold(uint8_t* o, uint32_t off, uint32_t len, uint32_t lane) { return emit_copy_v2(o, off, len, lane); }
__device__ SNB_COLD uint32_t crc32c_cold(const CrcSmem* t, const CrcTables* g, const uint8_t* p, uint32_t n, uint32_t lane) {
  return crc32c_warp(*t, g, p, n, lane);
}
__device__ SNB_COLD uint32_t compress_block_v5_cold(const uint8_t* in, uint32_t len, uint8_t* out, uint16_t* table, uint32_t abort_thr, uint32_t lane) {
  return compress_block_v5(in, len, out, table, abort_thr, lane);
}

// One full 64 KiB block; all kRing warps of the group call it (role r takes the turns = r mod kRing).  `turn`: in, the
// job's first turn (a multiple of kRing, the same in every warp); out, the turn that finished the block -- the next
// job starts at (turn / kRing + 2) * kRing.  Returns the encoded size or kAborted in every warp.  The table is zero,
// rec(turn) and com(turn - 1) are in the box.
__device__ __forceinline__ uint32_t compress_block_v7(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                      uint16_t* table, uint32_t abort_thr, uint32_t lane,
                                                      uint32_t role, Box* box, uint32_t& turn) {
  InView v;
  v.mis = (uint32_t)((uintptr_t)in & 3);
  v.w = (const uint32_t*)(in - v.mis);
  v.last = (kLen + v.mis - 1) >> 2;
  constexpr uint32_t len = kLen, mask = kMask, ip_limit = kIpLimit;
  const uint32_t below = (1u << lane) - 1u;
  const uint32_t notbelow = ~below;
  const uint32_t own_q = (lane + v.mis) >> 2, own_s = ((lane + v.mis) & 3) * 8;
  const uint32_t wlane = umin32(lane, 11u);
  uint32_t my_turn = turn + role;
  uint32_t b_pred = role * 32;          // the window this warp expects to get at its next turn
  uint32_t Wn = 0, Wn_b = kNoWindow;    // prefetched own words of window Wn_b
  // the resolved window r_b: own 12 bytes, hash, table entry seen, match length
  uint32_t ra0 = 0, ra1 = 0, ra2 = 0, rh = 0, rT = 0, rml = 0, r_b = kNoWindow;
  SNB_CYC_DECL;

  // resolve window bb against the table as it is now
#define SNB_RESOLVE(bb)                                                                                          \
  do {                                                                                                           \
    const uint32_t b_ = (bb);                                                                                    \
    const uint32_t W_ = (Wn_b == b_) ? Wn : __ldg(v.w + (b_ >> 2) + wlane);                                      \
    Wn_b = b_ + 32 * kRing;                                                                                      \
    Wn = __ldg(v.w + umin32((Wn_b >> 2) + wlane, v.last));                                                       \
    const uint32_t w0_ = __shfl_sync(kFull, W_, (int)own_q), w1_ = __shfl_sync(kFull, W_, (int)own_q + 1);       \
    const uint32_t w2_ = __shfl_sync(kFull, W_, (int)own_q + 2), w3_ = __shfl_sync(kFull, W_, (int)own_q + 3);   \
    ra0 = __funnelshift_r(w0_, w1_, own_s); ra1 = __funnelshift_r(w1_, w2_, own_s); ra2 = __funnelshift_r(w2_, w3_, own_s); \
    rh = hash14(ra0, mask);                                                                                      \
    rT = umin32(ld_tab(table + rh), b_ + lane);   /* a half-zeroed / not yet final table: stay inside the chunk */ \
    uint32_t c0_, c1_, c2_;                                                                                      \
    in_ld12u(v, rT, c0_, c1_, c2_);                                                                              \
    rml = match12(ra0, ra1, ra2, c0_, c1_, c2_);                                                                 \
    r_b = b_;                                                                                                    \
  } while (0)

  // a Fin record has arrived: pass it on to the warps that still wait for their turn, and leave
#define SNB_LEAVE_ON_FIN(s_)                                                                        \
  do {                                                                                              \
    const uint32_t hops_ = ((s_).y >> 24) & 15u;                                                    \
    __syncwarp();   /* every lane holds its copy of the record before it is overwritten */         \
    if (hops_ && lane == 0) {                                                                       \
      uint4 f_ = (s_); f_.x = (s_).x + 1; f_.y = ((s_).y & ~(15u << 24)) | ((hops_ - 1) << 24);     \
      st_rec(&box->rec, f_);                                                                        \
    }                                                                                               \
    turn = (s_).z;                                                                                  \
    SNB_CYC_FLUSH();                                                                                \
    return (s_).w;                                                                                  \
  } while (0)
  // the block is over: publish the result (rec, then com for a warp that waits for this turn's commit) and leave
#define SNB_PUBLISH_FIN(res)                                                                        \
  do {                                                                                              \
    const uint32_t res_ = (res);                                                                    \
    __syncwarp();                                                                                   \
    if (lane == 0) {                                                                                \
      uint4 f_; f_.x = my_turn + 1; f_.y = kFlagFin | ((kRing - 2) << 24) | (pub_nonce << 28); f_.z = my_turn; f_.w = res_; \
      st_rec(&box->rec, f_);                                                                        \
      uint4 c_; c_.x = my_turn; c_.y = pub_nonce; c_.z = 0; c_.w = 0;                               \
      st_rec(&box->com, c_);                                                                        \
    }                                                                                               \
    turn = my_turn;                                                                                 \
    SNB_CYC_FLUSH();                                                                                \
    return res_;                                                                                    \
  } while (0)
#define SNB_FINISH(from)                                                          \
  do {                                                                            \
    uint32_t f_ = (from);                                                         \
    if (f_ < len) {                                                               \
      uint32_t n_ = len - f_;                                                     \
      if (op + literal_size(n_) > abort_thr) SNB_PUBLISH_FIN(kAborted);           \
      op += emit_literal_v7(out + op, v, f_, n_, lane);                           \
    }                                                                             \
    SNB_PUBLISH_FIN(op);                                                          \
  } while (0)
#define SNB_UNPACK(s_)                                                                               \
  do {                                                                                               \
    chain = ((s_).y & kFlagChain) != 0; force_slow = ((s_).y & kFlagSlow) != 0; m1_done = ((s_).y & kFlagM1) != 0; \
    ip = chain ? ((s_).y & 0xfffffu) : 0u; pp = chain ? 1u : ((s_).y & 0xfffffu);                    \
    next_emit = (s_).z & 0xfffffu; k = (s_).z >> 20; rnonce = ((s_).y >> 28) & 1u;                   \
  } while (0)

  for (;;) {
    // ---------------------------------------------------------------- pre: resolve the window I expect to get
    SNB_CYC_START(b_pred);
    if (b_pred + 32 < ip_limit) {
      SNB_STAT(12);
      SNB_RESOLVE(b_pred);
    }
    SNB_CYC_MARK(0, rml);   // pre-resolve

    // ---------------------------------------------------------------- wait for the cursor
    uint4 s;
    for (;;) {
      s = ld_rec(&box->rec);
      if (s.x == my_turn) break;
      SNB_STAT(10);
      spin_pause();
    }
    SNB_CYC_MARK(1, s.x);   // waiting for the turn
#ifdef SNB_DEBUG_RING
    if (lane == 0) fprintf(stderr, "role %u turn %u: rec y=%08x z=%08x w=%u r_b=%u\n", role, my_turn, s.y, s.z, s.w, r_b);
#endif
    if (s.y & kFlagFin) SNB_LEAVE_ON_FIN(s);
    bool chain, force_slow, m1_done;
    uint32_t ip;          // chain mode: position of the pending post-copy lookup
    uint32_t pp;          // literal mode: next probe position
    uint32_t next_emit;   // literal mode: start of the pending literal
    uint32_t k;           // literal mode: probes already done in this run
    uint32_t rnonce;      // which publication of this turn's record I hold
    SNB_UNPACK(s);
    uint32_t op = 0;      // known once the previous turn has committed
    bool have_com = false;
    uint32_t pub_nonce = 0;   // 1: my first record for the next turn was wrong, this one replaces it
    bool exact = false;       // the window pass runs v5's exact conflict handling (after an optimistic attempt failed)

    // ---------------------------------------------------------------- fast turn: walk on the pre-resolved lengths, hand the
    // cursor over, and only then emit and commit.  Nothing is written before the previous turn's commit has been seen.
    {
      const uint32_t b = (chain ? ip : pp) & ~31u;
      bool fast = r_b == b && !force_slow && b + 32 < ip_limit && (chain || k < 32) && !(chain && ip == b && !m1_done);
      if (fast) {
        SNB_STAT(0);
        const uint32_t p = b + lane, h = rh;
        uint32_t A = 0, cc = 255, HB = 0, evt = 32, lkmask = 0, ins0 = 0, noprev = 0, run = 0, bend = 32, entry_n = 0;
        bool entry = false, allmiss = false, verified = false;
        for (;;) {
          const uint32_t ml = rml;
          HB = __ballot_sync(kFull, ml >= 4);
          const uint32_t qb = HB & notbelow;
          evt = qb ? (uint32_t)__ffs((int)qb) - 1u : 32u;
          const uint32_t eml = __shfl_sync(kFull, ml, (int)(evt & 31u));
          const uint32_t J = evt >= 32 ? 128u : (eml >= 12 ? 64u + evt : evt + eml);
          lkmask = notbelow & ((2u << umin32(evt, 31u)) - 1u);
          uint32_t c = 0;
          ins0 = 0; noprev = 0; entry = false; allmiss = false;
          if (!chain) {   // the literal run entering the window: stride-1 probes at lanes [l0, bend)
            const uint32_t l0 = pp - b;
            bend = umin32(32u, l0 + (32u - k));
            const uint32_t range = (0xffffffffu << l0) & (bend >= 32 ? 0xffffffffu : ((1u << bend) - 1u));
            const uint32_t qh = HB & range;
            const uint32_t q0 = qh ? (uint32_t)__ffs((int)qh) - 1u : 32u;
            run = range & (q0 >= 32 ? 0xffffffffu : ((2u << q0) - 1u));   // probes performed
            if (q0 >= 32) allmiss = true;
            else { entry = true; entry_n = b + q0 - next_emit; ins0 = run; c = q0; noprev = 1u << q0; }
          } else {
            c = ip - b;
          }
          A = 0; cc = 255;
          if (!allmiss) { cc = c; do { A |= 1u << cc; cc = __shfl_sync(kFull, J, (int)cc); } while (cc < 32); }
          if (verified) break;
          // ---- the previous turn's commit (and with it the output offset); my entries are final from here on
          uint4 cm;
          for (;;) {
            cm = ld_rec(&box->com);
            if (cm.x == my_turn - 1) break;
            SNB_STAT(15);
            spin_pause();
          }
          have_com = true; op = cm.z;
          if (cm.y != rnonce) { fast = false; break; }   // the record I walked on has been replaced
          const uint32_t T = ld_tab(table + h);
          const bool moved = T != rT;
          if (__any_sync(kFull, moved)) {   // an entry moved after I looked: does the walk depend on it?
            SNB_STAT(11);
            bool rel = (run >> lane) & 1u;   // a probe of the entering run
            if (!allmiss) {
              const uint32_t al = A & ((2u << lane) - 1u);                  // arrivals at or below me
              const uint32_t a = al ? 31u - (uint32_t)__clz((int)al) : 0u;
              const uint32_t ea = __shfl_sync(kFull, evt, (int)a);
              rel = rel || (al != 0 && ea >= lane)                          // looked up
                    || (((A >> 1) >> lane) & 1u)                            // inserted as the lane before an arrival
                    || (cc == 32 && lane == 31);                            // ... before the arrival at the next window's first lane
            }
            if (__any_sync(kFull, moved && rel)) {
              if (moved) {
                uint32_t c0, c1, c2;
                in_ld12u(v, T, c0, c1, c2);
                rml = match12(ra0, ra1, ra2, c0, c1, c2);
                rT = T;
              }
              verified = true;
              continue;   // walk again, on final lengths
            }
          }
          break;
        }
        // ---- where the cursor goes; the cases that need more than one batch take the slow turn
        bool n_chain = chain;
        uint32_t n_ip = ip, n_pp = pp, n_k = k, n_ne = next_emit;
        if (fast) {
          if (allmiss) { n_k = k + (uint32_t)__popc(run); n_pp = b + bend; }
          else if (cc >= 64 && cc < 96) fast = false;                 // a match reaches the 12-byte horizon
          else if (cc == 128) { const uint32_t cl = 31u - (uint32_t)__clz((int)A); n_chain = false; n_ne = b + cl; n_pp = b + 32; n_k = 31 - cl; }
          else { n_chain = true; n_ip = b + cc; if (n_ip > ip_limit) fast = false; }   // the block ends here
        }
        if (fast) {
          bool pend_insert = false, n_m1 = false;
          uint32_t ins = allmiss ? run : ins0;
          if (n_chain && (n_ip & 31u) == 0) {
            if (n_ip == b + 32) ins |= 1u << 31; else pend_insert = true;
            n_m1 = true;
          }
          // ---- hand the cursor over
          const uint32_t cur = n_chain ? n_ip : n_pp;
          if (lane == 0) {
            uint4 r;
            r.x = my_turn + 1;
            r.y = cur | (n_chain ? kFlagChain : 0u) | (n_m1 ? kFlagM1 : 0u);
            r.z = n_ne | (n_k << 20);
            r.w = 0;
            st_rec(&box->rec, r);
          }
          SNB_CYC_MARK(2, cur);   // cursor received -> handed over
          // ---- emit and commit, off the critical path
          const uint32_t T = rT, ml = rml, a0 = ra0;
          const uint32_t op0 = op;
          bool fail = false;
          if (entry) {
            if (op + literal_size(entry_n) > abort_thr) fail = true;
            else op += emit_literal_v7(out + op, v, next_emit, entry_n, lane);
          }
          if (!fail && !allmiss) {
            const bool onp = (A >> lane) & 1u;
            const uint32_t LK = __reduce_or_sync(kFull, onp ? lkmask : 0u);
            ins |= LK | ((A & ~noprev) >> 1);
            const uint32_t off = p - T;
            const bool c3 = off >= 2048;
            const uint32_t tb = c3 ? ((((ml - 1) << 2) | 2u) | (off << 8))
                                   : ((((off >> 8) << 5) | ((ml - 4) << 2) | 1u) | ((off & 0xff) << 8));
            const uint32_t C3 = __ballot_sync(kFull, c3);
            const uint32_t Ev = LK & HB;
            uint32_t tail = 0;
            if (cc == 128) tail = 0xffffffffu << (31u - (uint32_t)__clz((int)A));
            const uint32_t LitB = LK & ~Ev & ~tail;
            const uint32_t LS = A & ~HB & ~tail;
            const uint32_t E3 = Ev & C3;
            {
              const uint32_t offl = (uint32_t)(__popc(LitB & below) + __popc(LS & below) + 2 * __popc(Ev & below) + __popc(E3 & below));
              uint8_t* o = out + op + offl;
              if ((LS >> lane) & 1u) { o[0] = (uint8_t)((evt - lane - 1u) << 2); o += 1; }
              if ((LitB >> lane) & 1u) o[0] = (uint8_t)a0;
              if ((Ev >> lane) & 1u) {
                o[0] = (uint8_t)tb; o[1] = (uint8_t)(tb >> 8);
                if (c3) o[2] = (uint8_t)(tb >> 16);
              }
            }
            op += (uint32_t)(__popc(LitB) + __popc(LS) + 2 * __popc(Ev) + __popc(E3));
            if (op > abort_thr) fail = true;
          }
          if (!fail) {   // every inserted lane writes; two of them sharing a hash show up as a lost write
            __syncwarp();
            const bool mine = (ins >> lane) & 1u;
            if (mine) table[h] = (uint16_t)p;
            __syncwarp();
            if (__any_sync(kFull, mine && ld_tab(table + h) != p)) {
              __syncwarp();
              if (mine) table[h] = (uint16_t)T;   // every lane of a group read the same entry
              fail = true;
            }
          }
          if (!fail) {
            if (pend_insert) {
              __syncwarp();
              const uint32_t d1 = in_ld4(v, n_ip - 1);
              if (lane == 0) table[hash14(d1, mask)] = (uint16_t)(n_ip - 1);
            }
            __syncwarp();   // every lane's table writes are issued before the record
            if (lane == 0) { uint4 c_; c_.x = my_turn; c_.y = 0; c_.z = op; c_.w = 0; st_rec(&box->com, c_); }
            SNB_CYC_MARK(3, op);   // emit + commit
            SNB_CYC_COUNT(12);
            my_turn += kRing;
            b_pred = (cur & ~31u) + 32 * (kRing - 1);
            continue;
          }
          // the optimistic turn did not hold: take it again on the exact path and replace the record
          SNB_STAT(13);
          op = op0; exact = true; pub_nonce = 1;
        }
      }
    }

    // ---------------------------------------------------------------- slow turn: serial steps, then one window pass, in place
    if (!have_com) {
      uint4 cm;
      for (;;) {
        cm = ld_rec(&box->com);
        if (cm.x == my_turn - 1) break;
        SNB_STAT(15);
        spin_pause();
      }
      op = cm.z;
      if (cm.y != rnonce) rnonce = 2;   // reload below
    } else if (!exact) {
      const uint4 cm = ld_rec(&box->com);
      if (cm.y != rnonce) rnonce = 2;
    }
    if (rnonce == 2) {   // the record I hold was replaced before its turn committed: the replacement is there now
      s = ld_rec(&box->rec);
      if (s.y & kFlagFin) SNB_LEAVE_ON_FIN(s);
      SNB_UNPACK(s);
    }
    SNB_CYC_COUNT(10);
    for (;;) {
      __syncwarp();
      const uint32_t b = (chain ? ip : pp) & ~31u;
      SNB_STAT(8);
      if (force_slow || b + 32 >= ip_limit || (!chain && k >= 32)) {
        force_slow = false;
        m1_done = false;
        SNB_CYC_COUNT(chain ? 13 : 14);
        if (chain) {
          SNB_STAT(1);
          // ---- one serial chain step straight on the table (encoder.nim:364-380)
          const uint32_t d1 = in_ld4(v, ip - 1), d = in_ld4(v, ip);
          const uint32_t h1 = hash14(d1, mask), hh = hash14(d, mask);
          const uint32_t tv = table[hh];
          const uint32_t cand = (hh == h1) ? ip - 1 : tv;
          __syncwarp();
          if (lane == 0) { table[h1] = (uint16_t)(ip - 1); table[hh] = (uint16_t)ip; }
          const uint32_t m = common_prefix_cold(v.w, v.mis, v.last, len, cand, ip, lane);
          if (m < 4) { chain = false; next_emit = ip; pp = ip + 1; k = 0; continue; }   // :379-380
          op += emit_copy_cold(out + op, ip - cand, m, lane);
          ip += m;
          if (op > abort_thr) SNB_PUBLISH_FIN(kAborted);
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
            __syncwarp();
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
          if (op + literal_size(n) > abort_thr) SNB_PUBLISH_FIN(kAborted);
          op += emit_literal_v7(out + op, v, next_emit, n, lane);
        }
        uint32_t m = 4 + common_prefix_cold(v.w, v.mis, v.last, len, hit_cand + 4, hit_ip + 4, lane);
        op += emit_copy_cold(out + op, hit_ip - hit_cand, m, lane);
        ip = hit_ip + m;
        if (op > abort_thr) SNB_PUBLISH_FIN(kAborted);
        if (ip > ip_limit) SNB_FINISH(ip);
        chain = true;
        continue;
      }

      // ---- window [b, b+32): every lane is a valid lookup position (b + 32 < ip_limit)
      // what an optimistic pass may have to take back
      const uint32_t s_ip = ip, s_pp = pp, s_k = k, s_ne = next_emit, s_op = op;
      const bool s_chain = chain, s_m1 = m1_done;
#define SNB_REDO_EXACT()                                                                                  \
  do { ip = s_ip; pp = s_pp; k = s_k; next_emit = s_ne; op = s_op; chain = s_chain; m1_done = s_m1;      \
       force_slow = false; exact = true; SNB_STAT(13); } while (0)
      const uint32_t p = b + lane;
      if (chain && ip == b && !m1_done) {  // arrival on the window boundary after a serial step: ip-1 is not a lane here
        const uint32_t d1 = in_ld4(v, ip - 1);
        if (lane == 0) table[hash14(d1, mask)] = (uint16_t)(ip - 1);
        __syncwarp();
      }
      m1_done = false;

      // ---- resolve: take the pre-resolved window if it is this one, and check it against the final table
      SNB_STAT(9);
      if (r_b != b) { SNB_STAT(14); SNB_RESOLVE(b); }
      const uint32_t h = rh, a0 = ra0;
      const uint32_t T = ld_tab(table + h);
      if (__any_sync(kFull, T != rT)) {   // the entry moved after I looked: re-measure against the entry that counts
        if (T != rT) {
          uint32_t c0, c1, c2;
          in_ld12u(v, T, c0, c1, c2);
          rml = match12(ra0, ra1, ra2, c0, c1, c2);
          rT = T;
        }
      }
      const uint32_t ml = rml;
      const uint32_t off = p - T;
      const bool c3 = off >= 2048;   // three-byte copy tag (ml <= 11 whenever the tag is used, encoder.nim:113-125)
      const uint32_t tb = c3 ? ((((ml - 1) << 2) | 2u) | (off << 8))
                             : ((((off >> 8) << 5) | ((ml - 4) << 2) | 1u) | ((off & 0xff) << 8));
      const uint32_t grp = exact ? __match_any_sync(kFull, h) : (1u << lane);
      const uint32_t eg = grp & below;                   // earlier lanes of my hash group (exact pass only)
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

      uint32_t ins = 0;      // lanes inserted in this pass
      uint32_t c = 0;        // entry arrival lane
      uint32_t noprev = 0;   // arrival whose predecessor lane is not inserted again (a probe hit)
      bool redo = false;
      if (!chain) {
        // ---- the literal run entering the window: stride-1 probes at lanes [l0, bend)
        const uint32_t l0 = pp - b;
        const uint32_t bend = umin32(32u, l0 + (32u - k));
        const uint32_t range = (0xffffffffu << l0) & (bend >= 32 ? 0xffffffffu : ((1u << bend) - 1u));
        const uint32_t qh = HB & range;
        const uint32_t q0 = qh ? (uint32_t)__ffs((int)qh) - 1u : 32u;
        const uint32_t run = range & (q0 >= 32 ? 0xffffffffu : ((2u << q0) - 1u));   // probes performed
        if (exact) {
          const uint32_t stale = __ballot_sync(kFull, (eg & run) != 0) & run;
          if (stale) { SNB_STAT(5); force_slow = true; continue; }
        }
        if (q0 >= 32) {  // every probe missed: insert them and move on (k == 32 switches to the strided scan)
          SNB_STAT(6);
          __syncwarp();
          const bool mine = (run >> lane) & 1u;
          if (exact) {
            if (mine && ((grp & run) >> lane) == 1u) table[h] = (uint16_t)p;
          } else {
            if (mine) table[h] = (uint16_t)p;
            __syncwarp();
            if (__any_sync(kFull, mine && ld_tab(table + h) != p)) {   // two probes share a hash: take it back
              __syncwarp();
              if (mine) table[h] = (uint16_t)T;
              SNB_REDO_EXACT();
              continue;
            }
          }
          k += (uint32_t)__popc(run);
          pp = b + bend;
          break;   // pass complete: hand over
        }
        {  // :336-338; the copy at q0 is the first event of the walk
          const uint32_t n = b + q0 - next_emit;
          if (op + literal_size(n) > abort_thr) {
            if (!exact) { SNB_REDO_EXACT(); continue; }
            SNB_PUBLISH_FIN(kAborted);
          }
          op += emit_literal_v7(out + op, v, next_emit, n, lane);
        }
        ins = run;
        chain = true;
        c = q0;
        noprev = 1u << q0;
      } else {
        c = ip - b;
      }

      bool done = false;  // block finished inside this pass (ip > ip_limit after a copy)
      for (;;) {          // one batch per turn; a match that reaches the horizon ends a batch
        SNB_STAT(3);
        uint32_t A = 0, cc = c;
        do { A |= 1u << cc; cc = __shfl_sync(kFull, J, (int)cc); } while (cc < 32);
        bool onp = (A >> lane) & 1u;
        uint32_t LK = __reduce_or_sync(kFull, onp ? lkmask : 0u);
        uint32_t insb = LK | ((A & ~noprev) >> 1);
        uint32_t slow_at = 32;
        if (exact) {
          const uint32_t stale = __ballot_sync(kFull, ((LK >> lane) & 1u) && (eg & (ins | insb)) != 0);
          if (stale) {  // keep the arrivals whose lookups all precede the first stale lane
            const uint32_t xs = (uint32_t)__ffs((int)stale) - 1u;
            slow_at = 31u - (uint32_t)__clz((int)(A & ((2u << xs) - 1u)));
            A &= (1u << slow_at) - 1u;
            onp = (A >> lane) & 1u;
            LK = __reduce_or_sync(kFull, onp ? lkmask : 0u);
            insb = LK | ((A & ~noprev) >> 1);
            cc = 32;
          }
        }
        ins |= insb;
        noprev = 0;
        SNB_STATN(7, __popc(A));
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
        if (op > abort_thr) {
          if (!exact) { redo = true; break; }
          SNB_PUBLISH_FIN(kAborted);
        }
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
        const uint32_t m = 12 + common_prefix_cold(v.w, v.mis, v.last, len, cand + 12, b + e + 12, lane);
        op += emit_copy_cold(out + op, b + e - cand, m, lane);
        ip = b + e + m;
        SNB_CYC_COUNT(11);
        if (op > abort_thr) {
          if (!exact) { redo = true; break; }
          SNB_PUBLISH_FIN(kAborted);
        }
        if (ip > ip_limit) { done = true; break; }
        c = ip - b;
        if (c >= 32) break;
      }
      if (redo) { SNB_REDO_EXACT(); continue; }
      if (chain && !force_slow && ip > ip_limit) done = true;   // encoder.nim:362-363
      // the next arrival's ip-1 insert when no window pass will own it (lane c-1 of the arrival rule)
      bool pend_insert = false;
      if (chain && !force_slow && !done && (ip & 31u) == 0) {
        if (ip == b + 32) ins |= 1u << 31; else pend_insert = true;
        m1_done = true;
      }
      // ---- commit
      __syncwarp();   // after every lane's table reads of this pass
      {
        const bool mine = (ins >> lane) & 1u;
        if (exact) {   // the last inserted lane of each hash group writes the table
          if (mine && ((grp & ins) >> lane) == 1u) table[h] = (uint16_t)p;
        } else {       // every inserted lane writes; two of them sharing a hash show up as a lost write
          if (mine) table[h] = (uint16_t)p;
          __syncwarp();
          if (__any_sync(kFull, mine && ld_tab(table + h) != p)) {
            __syncwarp();
            if (mine) table[h] = (uint16_t)T;   // every lane of a group read the same entry
            SNB_REDO_EXACT();
            continue;
          }
        }
      }
      if (done) SNB_FINISH(ip);
      if (pend_insert) {
        __syncwarp();
        const uint32_t d1 = in_ld4(v, ip - 1);
        if (lane == 0) table[hash14(d1, mask)] = (uint16_t)(ip - 1);
      }
      break;   // pass complete: hand over
    }
#undef SNB_REDO_EXACT

    // ---------------------------------------------------------------- hand over: the cursor, and the commit with it
    const uint32_t cur = chain ? ip : pp;
    __syncwarp();   // every lane's table writes are issued before the records
    if (lane == 0) {
      uint4 r;
      r.x = my_turn + 1;
      r.y = cur | (chain ? kFlagChain : 0u) | (force_slow ? kFlagSlow : 0u) | (m1_done ? kFlagM1 : 0u) | (pub_nonce << 28);
      r.z = next_emit | (k << 20);
      r.w = 0;
      st_rec(&box->rec, r);
      uint4 c_; c_.x = my_turn; c_.y = pub_nonce; c_.z = op; c_.w = 0;
      st_rec(&box->com, c_);
    }
    SNB_CYC_MARK(4, cur);   // slow turn
    SNB_CYC_COUNT(12);
    my_turn += kRing;
    b_pred = (cur & ~31u) + 32 * (kRing - 1);
  }
#undef SNB_UNPACK
#undef SNB_FINISH
#undef SNB_PUBLISH_FIN
#undef SNB_LEAVE_ON_FIN
#undef SNB_RESOLVE
}

}  // namespace pair

// One job of the streaming kernels: encodeFrame (encoder.nim:385-426) or a bare encodeBlock, `enc(in, len, out, thr)`
// being the block encoder.  Same layout decisions as compress_job.
struct JobGeom {
  const uint8_t* in;
  uint8_t* out;      // the job's scratch slot
  uint32_t len, hl, pre, thr;
  bool frame;
};
__device__ __forceinline__ JobGeom job_geom(const CompressArgs& a, uint32_t job) {
  JobGeom j;
  const uint64_t ioff = (uint64_t)job * kMaxBlock;
  j.len = (uint32_t)((a.total_len - ioff) < kMaxBlock ? (a.total_len - ioff) : kMaxBlock);
  j.in = a.in_base + ioff;
  j.out = a.out_base + (uint64_t)job * a.out_stride;
  j.frame = a.mode == kModeFrame;
  j.hl = a.mode == kModeRawBlock ? 0u : varint32_len(j.len);
  j.pre = j.frame ? 8u + j.hl : j.hl;
  j.thr = j.frame ? j.len - j.len / 8 : kNoAbort;
  return j;
}
// what follows the block encoder: header / stored fallback (encoder.nim:408-426).  bl: encoded size or kAborted.
__device__ __forceinline__ uint32_t job_finish(const CompressArgs& a, const JobGeom& j, uint32_t bl, uint32_t crc, uint32_t lane) {
  if (!j.frame) return j.hl + bl;
  uint32_t fl;
  if (bl != kAborted && bl <= j.thr) {
    put_varint32(j.out + 8, j.len, lane);
    fl = j.hl + bl + 4;
    if (lane == 0) j.out[0] = 0x00;
  } else {
    fl = j.len + 4;
    if (lane == 0) j.out[0] = 0x01;
    if (!a.defer_stored) warp_copy(j.out + 8, j.in, j.len, lane);
  }
  if (lane == 0) {
    j.out[1] = (uint8_t)fl; j.out[2] = (uint8_t)(fl >> 8); j.out[3] = (uint8_t)(fl >> 16);
    j.out[4] = (uint8_t)crc; j.out[5] = (uint8_t)(crc >> 8); j.out[6] = (uint8_t)(crc >> 16); j.out[7] = (uint8_t)(crc >> 24);
  }
  return fl + 4;
}

namespace pair {
__device__ SNB_COLD void stream_place_cold(const StreamArgs& a, uint32_t lane) { stream_place(a, lane); }
}
// Streaming compress, v7 geometry: blockDim.x / (32 kRing) warp groups per CTA, one chunk per group; the first
// a.c.smem_warps groups keep their table in shared memory, further groups in a.c.gtables.  Arrival wait, CRC,
// in-order packing and group flags are compress_stream_kernel's, done by the group's first warp.
__global__ void __launch_bounds__(pair::kMaxThreads, 1) compress_pair_kernel(StreamArgs a) {
  SNB_DYN_SMEM(smem);
  constexpr uint32_t R = pair::kRing;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t pr = warp / R, role = warp % R, npairs = blockDim.x / (32 * R);
  const uint32_t spairs = a.c.smem_warps;
  const size_t tbytes = (size_t)kMaxTableEntries * 2;
  uint16_t* table = pr < spairs ? (uint16_t*)(smem + pr * tbytes)
                                : a.c.gtables + ((size_t)blockIdx.x * (npairs - spairs) + (pr - spairs)) * kMaxTableEntries;
  pair::Box* box = (pair::Box*)(smem + spairs * tbytes) + pr;
  CrcSmem* crc_t = (CrcSmem*)(smem + spairs * tbytes + npairs * sizeof(pair::Box));
  const bool frame = a.c.mode == kModeFrame;
  if (a.span && threadIdx.x == 0) atomicMin(a.span, snb_now_ns());
  if (role == 0 && lane == 0) {
    uint4 z; z.x = 0xffffffffu; z.y = 0; z.z = 0; z.w = 0;
    box->rec = z; box->com = z; box->job = 0; box->job_seq = 0; box->ack = 0; box->fin = 0;
  }
  if (frame) crc_stage_tables(crc_t, a.crc_tables);   // ends with __syncthreads()
  else __syncthreads();

  constexpr uint32_t kShare = kMaxTableEntries / 2 / R + 1;   // words of the table each warp zeroes
  uint32_t turn = R, jseq = 0;   // the first job's first turn (turn - 1 must not be 0xffffffff: the boxes start out with that)
  for (;;) {
    jseq++;
    if (role == 0) {
      uint32_t job = 0;
      if (lane == 0) job = atomicAdd(a.c.counter, 1u);
      job = __shfl_sync(kFull, job, 0);
      if (job >= a.c.count) {
        if (lane == 0) { box->job = pair::kJobExit; pair::st_flag(&box->job_seq, jseq); }
        if (a.span && lane == 0) atomicMax(a.span + 1, snb_now_ns());
        return;
      }
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
      const JobGeom j = job_geom(a.c, job);
      const bool paired = j.len == pair::kLen;
#ifdef SNB_DEBUG_RING
      if (lane == 0) fprintf(stderr, "group %u leader: job %u len %u paired %d npairs %u blockDim %u\n", pr, job, j.len, (int)paired, npairs, (unsigned)blockDim.x);
#endif
      if (paired) {   // the first records, then the job: the partners zero their share of the table and start resolving
        if (lane == 0) {
          uint4 r; r.x = turn; r.y = 1u; r.z = 0; r.w = 0;   // literal mode, next probe at 1 (encoder.nim:281)
          pair::st_rec(&box->rec, r);
          uint4 c; c.x = turn - 1; c.y = 0; c.z = 0; c.w = 0;   // nothing to wait for, output offset 0
          pair::st_rec(&box->com, c);
          box->job = job;
          pair::st_flag(&box->job_seq, jseq);
        }
        uint32_t* t32 = (uint32_t*)table;
        for (uint32_t i = lane; i < kShare; i += 32) t32[i] = 0;
      } else if (lane == 0) {
        box->job = job;
        pair::st_flag(&box->job_seq, jseq);
      }
      uint32_t crc = 0;
      if (frame) crc = crc_mask(pair::crc32c_cold(crc_t, a.crc_tables, j.in, j.len, lane));
      uint32_t bl = frame ? kAborted : 0u;
      if (frame ? j.len >= kMinNonLiteral : j.len > 0) {
        if (paired) {
          while (pair::ld_flag(&box->ack) != jseq * (R - 1)) pair::spin_pause();   // the partners' shares of the table are zero
          __syncwarp();
          bl = pair::compress_block_v7(j.in, j.out + j.pre, table, j.thr, lane, 0, box, turn);
          turn = (turn / R + 2) * R;
#ifdef SNB_DEBUG_RING
          if (lane == 0) fprintf(stderr, "group %u leader: block done bl=%u turn=%u fin=%u\n", pr, bl, turn, box->fin);
#endif
          while (pair::ld_flag(&box->fin) != jseq * (R - 1)) pair::spin_pause();   // the partners' output bytes are out
          __syncwarp();
#ifdef SNB_DEBUG_RING
          if (lane == 0) fprintf(stderr, "group %u leader: partners done\n", pr);
#endif
        } else {
          bl = pair::compress_block_v5_cold(j.in, j.len, j.out + j.pre, table, j.thr, lane);
          if (lane == 0) { box->ack += R - 1; box->fin += R - 1; }   // the counters keep step with job_seq
        }
      } else if (lane == 0) { box->ack += R - 1; box->fin += R - 1; }
      const uint32_t total = job_finish(a.c, j, bl, crc, lane);
      __syncwarp();
      if (lane == 0) {
        __threadfence();                              // the partners' fenced bytes and mine ...
        st_release_gpu(a.c.out_len + job, total);     // ... before the frame's length
        __threadfence();                              // ... and the length before this warp looks at the frontier
      }
      __syncwarp();
      pair::stream_place_cold(a, lane);
      __syncwarp();
    } else {
      while (pair::ld_flag(&box->job_seq) != jseq) { SNB_STAT(15); pair::spin_pause(); snb_sleep(200); }
      const uint32_t job = *(volatile uint32_t*)&box->job;
      if (job == pair::kJobExit) return;
      const JobGeom j = job_geom(a.c, job);
      if (j.len != pair::kLen || !(frame ? j.len >= kMinNonLiteral : j.len > 0)) continue;   // a short chunk: the first warp encodes it alone
      uint32_t* t32 = (uint32_t*)table;
      for (uint32_t i = role * kShare + lane; i < umin32((role + 1) * kShare, kMaxTableEntries / 2); i += 32) t32[i] = 0;
      __syncwarp();
      if (lane == 0) atomicAdd(&box->ack, 1u);
#ifdef SNB_DEBUG_RING
      if (lane == 0) fprintf(stderr, "group %u role %u: acked job %u (ack %u)\n", pr, role, job, box->ack);
#endif
      pair::compress_block_v7(j.in, j.out + j.pre, table, j.thr, lane, role, box, turn);
      turn = (turn / R + 2) * R;
      __syncwarp();
      if (lane == 0) { __threadfence(); atomicAdd(&box->fin, 1u); }
#ifdef SNB_DEBUG_RING
      if (lane == 0) fprintf(stderr, "group %u role %u: fin %u\n", pr, role, box->fin);
#endif
    }
  }
}

}  // namespace snb
