// Dev micro-benchmark (B200): latencies the mask decoder's cluster kernel (dec_tok.cu) is built from:
// mma.sync m16n8k16 bf16 dependent chain / 4 independent chains, movmatrix, an L2-hit LDG.128, st.shared::cluster +
// barrier.cluster round trip for cluster sizes 8, cluster barrier alone, __syncthreads with 512 threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 64
__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__global__ void __cluster_dims__(8, 1, 1) k(float* out, long long* cyc, const uint4* gmem, int seed) {
  __shared__ float land[8][512];
  const int tid = threadIdx.x, lane = tid & 31;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool rec = tid == 0 && rank == 0;
  long long t0, t1;
  float d[4][4] = {};
  uint32_t a = 0x3f803f80u + seed, b = 0x3f803f80u;
  // 1. dependent mma chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) mma(d[0], a, a, a, a, b, b);
  t1 = clock64(); if (rec) cyc[0] = (t1 - t0);
  // 2. four independent chains
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) { mma(d[0], a, a, a, a, b, b); mma(d[1], a, a, a, a, b, b); mma(d[2], a, a, a, a, b, b); mma(d[3], a, a, a, a, b, b); }
  t1 = clock64(); if (rec) cyc[1] = (t1 - t0);
  // 3. movmatrix chain
  uint32_t m = a;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(m) : "r"(m));
  t1 = clock64(); if (rec) cyc[2] = (t1 - t0);
  // 4. dependent L2-hit LDG.128 chain (pointer chasing through the x component; buffer >> L1)
  uint32_t idx = lane + seed;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { uint4 v = __ldcg(gmem + (idx & 0xffff)); idx = v.x + lane; }
  t1 = clock64(); if (rec) cyc[3] = (t1 - t0);
  // 5. cluster barrier alone
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  t1 = clock64(); if (rec) cyc[4] = (t1 - t0);
  // 6. every thread stores one float to all 8 CTAs + cluster barrier
  uint32_t la = (uint32_t)__cvta_generic_to_shared(&land[rank][tid]);
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      uint32_t ra; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(r));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(d[0][0] + i) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  t1 = clock64(); if (rec) cyc[5] = (t1 - t0);
  // 7. __syncthreads
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) __syncthreads();
  t1 = clock64(); if (rec) cyc[6] = (t1 - t0);
  // 8. only 64 threads store 8 floats each to all CTAs + barrier
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    if (tid < 64) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        uint32_t ra; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(r));
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra + j * 256), "f"(d[0][0] + i) : "memory");
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  t1 = clock64(); if (rec) cyc[7] = (t1 - t0);
  // 9. warp_sum (5 shuffles) chain
  float x = d[0][0];
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) { for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o); }
  t1 = clock64(); if (rec) cyc[8] = (t1 - t0);
  out[blockIdx.x * blockDim.x + tid] = d[0][0] + d[1][1] + d[2][2] + d[3][3] + m + idx + land[tid & 7][tid] + x;
}
int main() {
  float* out; long long* cyc; uint4* g;
  cudaMalloc(&out, 8 * 512 * 4); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&g, 65536 * 16); cudaMemset(g, 0, 65536 * 16);
  const char* names[] = {"mma.sync m16n8k16 bf16, dependent", "mma.sync x4 independent chains (per 4)", "movmatrix dependent", "LDG.128 L2-hit dependent",
                         "barrier.cluster (8 CTAs x 512 thr)", "8 remote stores/thread (512 thr) + barrier.cluster", "__syncthreads (512 thr)",
                         "64 remote stores/thread (64 thr) + barrier.cluster", "warp_sum (5 shfl) dependent"};
  for (int rep = 0; rep < 2; ++rep) { k<<<8, 512>>>(out, cyc, g, rep); cudaDeviceSynchronize(); }
  long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  for (int i = 0; i < 9; ++i) printf("  %-55s %8.1f cycles/op\n", names[i], (double)h[i] / N);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
