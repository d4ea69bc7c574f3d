// Dependent-chain latency of the warp collectives the encoder's walk is made of (one warp, sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 512
template <int OP> __global__ void k(uint32_t* out, uint32_t seed, const uint32_t* gmem) {
  __shared__ uint32_t sm[1024];
  uint32_t lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i * 7 + 3) & 1023;
  __syncwarp();
  uint32_t x = seed + lane * 2654435761u;
  uint32_t t0 = clock();
#pragma unroll 1
  for (int i = 0; i < N; i++) {
    if (OP == 0) x = __shfl_sync(0xffffffffu, x, (x >> 3) & 31) + 1;
    if (OP == 1) x = __ballot_sync(0xffffffffu, (x >> lane) & 1) + lane;
    if (OP == 2) x = __match_any_sync(0xffffffffu, x & 0x3fff) + i;
    if (OP == 3) x = __reduce_or_sync(0xffffffffu, x) + lane;
    if (OP == 4) x = __popc(x) + seed;
    if (OP == 5) x = __ffs(x) + seed;
    if (OP == 6) x = sm[x & 1023];
    if (OP == 7) x = x * 0x1e35a7bdu + 1;
    if (OP == 8) x = (x ^ (x >> 3)) + 1;   // 2 dependent ALU ops
    if (OP == 9) x = __match_any_sync(0xffffffffu, (x + lane) | 0x10000) + i;   // all distinct
    if (OP == 10) x = __ldg(gmem + (x & 1023));
    if (OP == 11) { x = __shfl_sync(0xffffffffu, x, (x >> 3) & 31); if (x & 0x80000000u) x += 3; else x += 1; }
    if (OP == 12) x = __funnelshift_r(x, seed, x & 31) + 1;
    if (OP == 13) { __syncwarp(); x = x + 1; }
  }
  uint32_t t1 = clock();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) out[32] = t1 - t0;
}
int main() {
  uint32_t *d, *g; cudaMalloc(&d, 256); cudaMalloc(&g, 4096);
  uint32_t hg[1024]; for (int i = 0; i < 1024; i++) hg[i] = (i * 13 + 5) & 1023; cudaMemcpy(g, hg, 4096, cudaMemcpyHostToDevice);
  const char* names[] = {"shfl.idx+add", "ballot+add", "match.any(14 bit)+add", "redux.or+add", "popc+add", "ffs+add", "lds (dependent)", "imad", "2 alu", "match.any distinct+add", "ldg L1 hit", "shfl+branchy add", "shf+add", "syncwarp+add"};
#define RUN(OP) { k<OP><<<1, 32>>>(d, 12345, g); k<OP><<<1, 32>>>(d, 12345, g); cudaDeviceSynchronize(); uint32_t h[33]; cudaMemcpy(h, d, 132, cudaMemcpyDeviceToHost); printf("%-28s %6.1f cycles per iteration\n", names[OP], h[32] / (double)N); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
