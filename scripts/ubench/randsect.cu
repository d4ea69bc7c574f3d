// What the HBM sustains for the encoder's access pattern: random 32-byte sectors out of a working set larger than the
// L2, read (and partly written) by 148 x 22 warps.  Independent loads (the memory system's limit) and dependent
// chains of one load per lane (what a latency-bound warp sees).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
template <int MLP, bool WRITE>
__global__ void __launch_bounds__(704, 1) k(const uint32_t* buf, uint32_t* wbuf, uint64_t words, uint32_t iters, uint32_t* sink) {
  uint32_t x = mix(blockIdx.x * 1024 + threadIdx.x + 1), acc = 0;
  const uint32_t smask = (uint32_t)(words / 8) - 1;   // power of two
  for (uint32_t i = 0; i < iters; i++) {
    uint32_t v[MLP];
#pragma unroll
    for (int j = 0; j < MLP; j++) { x = mix(x + j + 1); v[j] = __ldg(buf + (uint64_t)(x & smask) * 8 + (threadIdx.x & 7)); }
#pragma unroll
    for (int j = 0; j < MLP; j++) acc += v[j];
    if (WRITE) { x = mix(x); wbuf[(uint64_t)(x & smask) * 8] = (uint16_t)acc; }
    if (MLP == 1) x ^= acc & 1;   // dependent chain
  }
  if (acc == 0x12345678u) *sink = acc;
}
int main() {
  uint32_t *buf, *wbuf, *sink; cudaMalloc(&buf, 1ull << 30); cudaMalloc(&wbuf, 1ull << 30); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, 1ull << 30); cudaMemset(wbuf, 0, 1ull << 30);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(MLP, WR, name) { const uint32_t iters = 2000 / MLP; k<MLP, WR><<<148, 704>>>(buf, wbuf, words, 20, sink); cudaEventRecord(e0); \
    k<MLP, WR><<<148, 704>>>(buf, wbuf, words, iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); \
    const double sect = 148.0 * 704 * iters * (MLP + (WR ? 1 : 0)); \
    printf("%5llu MiB  %-46s %8.3f ms  %7.1f G sectors/s  %6.2f TB/s of 32-byte sectors\n", (unsigned long long)(words >> 18), name, ms, sect / ms / 1e6, sect * 32 / ms / 1e9); }
  for (uint64_t mib : {32ull, 64ull, 128ull, 256ull, 512ull, 1024ull}) {
    const uint64_t words = (mib << 20) / 4;
    RUN(1, false, "dependent, 1 load in flight per lane")
    RUN(8, false, "8 independent loads per lane")
    RUN(8, true, "8 loads + one 16-bit store per lane")
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
