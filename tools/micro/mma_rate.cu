// Micro-benchmark: issue rate / throughput of tcgen05.mma (kind::f16, bf16 -> f32, cta_group::1, M = 128) by shape:
// N in {64, 128, 256}, A from shared memory (SS) or tensor memory (TS), B K-major or MN-major.  One CTA per SM runs
// `reps` back-to-back MMAs from ONE thread on garbage operands and reports clock64 cycles per MMA (issue loop only, and
// until the commit lands).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I video-llava-seg_b200/csrc
//   tools/micro/mma_rate.cu -o tools/micro/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace vls;

template <int N, bool TS, bool BMN>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N) | (BMN ? (1u << 16) : 0u);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 65536);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int kk = r & 3;
      const uint64_t bd = BMN ? make_desc_sw128(b_addr + (r & 7) * 2048) : make_desc_sw128(b_addr + ((r >> 2) & 3) * (N * 128)) + 2 * kk;
      if (TS) umma_ts(tmem, tmem + 256 + (r & 15) * 8, bd, idesc, 1u);
      else umma_ss(tmem, make_desc_sw128(a_addr + ((r >> 2) & 3) * (128 * 128)) + 2 * kk, bd, idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, bool TS, bool BMN>
void run(const char* name, int ctas, long long* d) {
  const int reps = 512, smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<N, TS, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int w = 0; w < 2; ++w) rate_kernel<N, TS, BMN><<<ctas, 128, smem>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s ctas=%3d  issue %6.1f cyc/MMA   complete %6.1f cyc/MMA   (ideal %d)  %s\n", name, ctas, double(h[0]) / reps,
         double(h[1]) / reps, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int ctas : {1, 148}) {
    run<64, false, false>("SS N=64  B K-major", ctas, d);
    run<128, false, false>("SS N=128 B K-major", ctas, d);
    run<256, false, false>("SS N=256 B K-major", ctas, d);
    run<64, true, false>("TS N=64  B K-major", ctas, d);
    run<64, true, true>("TS N=64  B MN-major", ctas, d);
    run<128, true, false>("TS N=128 B K-major", ctas, d);
    run<256, true, false>("TS N=256 B K-major", ctas, d);
  }
  return 0;
}
