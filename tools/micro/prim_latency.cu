// Dev micro-benchmark: dependent-chain latency (cycles) of the warp primitives the CC kernels are built from.
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void k(int* out, long long* cyc, int seed) {
  __shared__ int sm[1024];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7 + seed) & 1023;
  __syncthreads();
  int x = seed + lane;
  long long t0, t1;
  // shfl chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_up_sync(0xffffffffu, x, 1) + 1;
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  // ballot chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __ballot_sync(0xffffffffu, (x >> lane) & 1) + lane;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  // redux min chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __reduce_min_sync(0xffffffffu, x + lane) + 1;
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  // match_any chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __match_any_sync(0xffffffffu, x & 3) + lane;
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  // LDS dependent chain
  int p = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) p = ((volatile int*)sm)[p];
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  // ATOMS with result, spread addresses, dependent
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) p = atomicMin(&sm[(p + lane) & 1023], p) & 1023;
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  // ATOMS no result, same address (all lanes)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) atomicAdd(&sm[5], x + i);
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  // ATOMS no result, same address, one lane per warp
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) if (lane == 0) atomicAdd(&sm[6], x + i);
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0);
  // reduce_add with match groups (non-uniform masks)
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) { unsigned m = __match_any_sync(0xffffffffu, (x + i) & 7); x += __reduce_add_sync(m, lane); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0);
  // integer division by a runtime value
  int d = (seed & 127) + 3;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x / d + 100000 + i;
  t1 = clock64(); if (threadIdx.x == 0) cyc[9] = (t1 - t0);
  out[threadIdx.x] = x + p;
}
int main() {
  int* out; long long* cyc;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 16 * 8);
  const char* names[] = {"shfl_up", "ballot", "redux.min", "match.any", "LDS chain", "ATOMS.min w/ result (spread)", "ATOMS.add same addr 32 lanes (no result)", "ATOMS.add same addr 1 lane", "match.any + reduce_add(groups of 8 keys)", "int div runtime"};
  for (int threads : {32, 512}) {
    k<<<1, threads>>>(out, cyc, 3); cudaDeviceSynchronize();
    k<<<1, threads>>>(out, cyc, 5); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads/CTA = %d\n", threads);
    for (int i = 0; i < 10; ++i) printf("  %-45s %7.1f cycles/op\n", names[i], (double)h[i] / N);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
