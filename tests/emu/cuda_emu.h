// tests/emu/cuda_emu.h -- single-threaded SIMT emulator (TEST DOUBLE, never shipped).
//
// Lets g++ compile the product's device code (nim-snappy_b200/csrc/*.cuh) and run each CUDA
// thread as a ucontext fiber, so the kernels' warp-level logic (ballot / shfl / match_any /
// redux / syncwarp / syncthreads) can be checked against the oracle on a machine without a
// GPU.  Every warp collective is a rendezvous of all 32 lanes of the warp; kernels must
// call collectives convergently with the full mask (which is how the product code is written).
//
// This is test infrastructure in the same sense as oracle/: only tests/ builds and loads it.
// The product library has no CPU path.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct emu_dim3 { unsigned x, y, z; };

namespace emu {

constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 128 * 1024;

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
};

struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = 0;       // running thread id
  int nthreads = 0;
  // per-warp rendezvous state
  int warp_arrived[kMaxThreads / 32];
  int warp_gen[kMaxThreads / 32];
  uint64_t slots[2][kMaxThreads];
  int warp_coll[kMaxThreads / 32];  // collective counter (for slot double-buffering)
  int blk_arrived = 0, blk_gen = 0;
  uint8_t* dyn_smem = nullptr;
  emu_dim3 block_idx{0, 0, 0}, block_dim{1, 1, 1}, grid_dim{1, 1, 1};
  std::function<void()> body;
};

inline Block*& B() { static Block* b = nullptr; return b; }

inline void yield() { Block* b = B(); swapcontext(&b->fibers[b->cur].ctx, &b->sched); }

inline void warp_rendezvous() {
  Block* b = B();
  int w = b->cur >> 5;
  int gen = b->warp_gen[w];
  if (++b->warp_arrived[w] == 32) { b->warp_arrived[w] = 0; b->warp_gen[w]++; return; }
  while (b->warp_gen[w] == gen) yield();
}

inline void block_rendezvous() {
  Block* b = B();
  int gen = b->blk_gen;
  if (++b->blk_arrived == b->nthreads) { b->blk_arrived = 0; b->blk_gen++; return; }
  while (b->blk_gen == gen) yield();
}

// exchange: every lane deposits v; after the rendezvous each lane may read any lane's value.
// The slot buffer alternates per collective so a fast lane cannot clobber unread values.
inline uint64_t* exchange(uint64_t v) {
  Block* b = B();
  int w = b->cur >> 5, lane = b->cur & 31;
  int buf = b->warp_coll[w] & 1;  // all lanes observe the same counter value before rendezvous
  uint64_t* s = &b->slots[buf][w * 32];
  s[lane] = v;
  int gen = b->warp_gen[w];
  if (++b->warp_arrived[w] == 32) { b->warp_arrived[w] = 0; b->warp_coll[w]++; b->warp_gen[w]++; }
  else while (b->warp_gen[w] == gen) yield();
  return s;
}

inline void fiber_entry() {
  Block* b = B();
  b->body();
  b->fibers[b->cur].done = true;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

template <class F>
void launch(unsigned grid, unsigned block, size_t smem_bytes, F&& kernel_call) {
  if (block % 32 != 0 || block > (unsigned)kMaxThreads) abort();
  Block blk;
  blk.nthreads = block;
  blk.fibers.resize(block);
  for (auto& f : blk.fibers) f.stack = (char*)malloc(kStack);
  std::vector<uint8_t> smem(smem_bytes + 64);
  blk.dyn_smem = (uint8_t*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
  blk.block_dim = {block, 1, 1};
  blk.grid_dim = {grid, 1, 1};
  blk.body = kernel_call;
  Block* saved = B();
  B() = &blk;
  for (unsigned bi = 0; bi < grid; bi++) {
    blk.block_idx = {bi, 0, 0};
    memset(blk.warp_arrived, 0, sizeof blk.warp_arrived);
    memset(blk.warp_gen, 0, sizeof blk.warp_gen);
    memset(blk.warp_coll, 0, sizeof blk.warp_coll);
    blk.blk_arrived = 0;
    for (unsigned t = 0; t < block; t++) {
      Fiber& f = blk.fibers[t];
      f.done = false;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    unsigned remaining = block;
    while (remaining) {
      unsigned progressed = 0;
      for (unsigned t = 0; t < block; t++) {
        Fiber& f = blk.fibers[t];
        if (f.done) continue;
        blk.cur = (int)t;
        swapcontext(&blk.sched, &f.ctx);
        if (f.done) { remaining--; }
        progressed++;
      }
      if (!progressed) abort();
    }
  }
  B() = saved;
  for (auto& f : blk.fibers) free(f.stack);
}

inline emu_dim3 tid() { Block* b = B(); return {(unsigned)b->cur, 0, 0}; }

}  // namespace emu

#define threadIdx (emu::tid())
#define blockIdx (emu::B()->block_idx)
#define blockDim (emu::B()->block_dim)
#define gridDim (emu::B()->grid_dim)
#define SNB_DYN_SMEM(name) uint8_t* name = emu::B()->dyn_smem

// ------------------------------------------------------------------ warp collectives
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_rendezvous(); }
static inline void __syncthreads() { emu::block_rendezvous(); }

template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  uint64_t* s = emu::exchange(raw);
  T out;
  memcpy(&out, &s[src & 31], sizeof(T));
  return out;
}
template <class T>
static inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
  int lane = emu::B()->cur & 31;
  return __shfl_sync(m, v, lane >= (int)d ? lane - (int)d : lane);
}
template <class T>
static inline T __shfl_down_sync(unsigned m, T v, unsigned d) {
  int lane = emu::B()->cur & 31;
  return __shfl_sync(m, v, lane + (int)d < 32 ? lane + (int)d : lane);
}
template <class T>
static inline T __shfl_xor_sync(unsigned m, T v, int x) {
  int lane = emu::B()->cur & 31;
  return __shfl_sync(m, v, lane ^ x);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  uint64_t* s = emu::exchange(pred ? 1 : 0);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(s[i] & 1) << i;
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline unsigned __match_any_sync(unsigned, unsigned v) {
  int lane = emu::B()->cur & 31;
  uint64_t* s = emu::exchange(v);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(s[i] == s[lane]) << i;
  return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  uint64_t* s = emu::exchange(v);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)s[i];
  return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  uint64_t* s = emu::exchange(v);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r += (unsigned)s[i];
  return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  uint64_t* s = emu::exchange(v);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r = (unsigned)s[i] > r ? (unsigned)s[i] : r;
  return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  uint64_t* s = emu::exchange(v);
  unsigned r = 0xffffffffu;
  for (int i = 0; i < 32; i++) r = (unsigned)s[i] < r ? (unsigned)s[i] : r;
  return r;
}

// ------------------------------------------------------------------ scalar intrinsics
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  return (unsigned)((((uint64_t)hi << 32) | lo) >> (sh & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
  return (unsigned)(((((uint64_t)hi << 32) | lo) << (sh & 31)) >> 32);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  uint64_t v = ((uint64_t)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
template <class T>
static inline T __ldg(const T* p) {
  if (((uintptr_t)p) % alignof(T) != 0) abort();   // the GPU faults on a misaligned load: so does the emulator
  return *p;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p; *p = o + v; return o;
}
static inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __threadfence_block() {}
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
  unsigned long long o = *p; if (o == cmp) *p = v; return o;
}
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
