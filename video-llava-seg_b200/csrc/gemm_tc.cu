// tcgen05 / TMEM / TMA GEMM with fused epilogues for every dense contraction on the path
// (q/k/v/out projections with axial RoPE, FFN, 1x1 convs, ConvTranspose-as-GEMM, pwconvs).
//
//   C[b][m][n] = epi( sum_k A[b][m][k] * W[n][k] )      A, W: bf16, K-major;  accumulate f32 in TMEM
//
// One CTA computes a 128 x BN tile.  Warp 0 = TMA producer (4-stage mbarrier ring, 128B-swizzled
// [128 x 64] / [BN x 64] boxes), warp 1 = single-thread tcgen05.mma issuer + TMEM owner,
// warps 2-5 = epilogue (each warp drains its 32-lane TMEM quarter with tcgen05.ld 32x32b.x32 and
// applies bias / ReLU / GELU / RoPE / residual before the store).
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int THREADS = 192;
constexpr int A_BYTES = BM * BK * 2;

struct Epi {
  int M, N, K;
  const float* bias;
  int bias_mode, act;
  const float* rope_cos;
  const float* rope_sin;
  int rope_period, rope_rows;
  const float* residual;
  long long ld_res, res_bstride;
  void* C;
  int c_bf16;
  long long ldc, c_bstride;
  int w_batched, a_batched;
};

template <int BN>
constexpr int smem_bytes() {
  return STAGES * (A_BYTES + BN * BK * 2) + 128 + 1024;
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Epi e) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int bz = blockIdx.z;
  const int kblocks = (e.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int wz = e.w_batched ? bz : 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        uint8_t* sa = smem + s * STAGE_BYTES;
        tma_load_3d(sa, &tmA, &full[s], kb * BK, m0, e.a_batched ? bz : 0);
        tma_load_3d(sa + A_BYTES, &tmB, &full[s], kb * BK, n0, wz);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_desc_sw128(sa);
        const uint64_t bdesc = make_desc_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < e.M;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const float bias_row = (e.bias_mode == 2 && row_ok) ? e.bias[row] : 0.0f;
    const bool do_rope = e.rope_cos != nullptr && row < e.rope_rows;
    const int pos = do_rope ? (row % e.rope_period) : 0;
    const float* res_row = e.residual ? e.residual + (long long)bz * e.res_bstride + (long long)row * e.ld_res : nullptr;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem + (uint32_t(q * 32) << 16) + c * 32, r);
      tc_wait_ld();
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= e.N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (e.bias_mode == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < e.N) v[j] += __ldg(e.bias + col0 + j);
      } else if (e.bias_mode == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += bias_row;
      }
      if (e.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
      } else if (e.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      }
      if (do_rope) {
        const int pair0 = (col0 & 255) >> 1;
        const float* cs = e.rope_cos + (long long)pos * 128 + pair0;
        const float* sn = e.rope_sin + (long long)pos * 128 + pair0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float co = __ldg(cs + j), si = __ldg(sn + j);
          const float a = v[2 * j], b = v[2 * j + 1];
          v[2 * j] = a * co - b * si;
          v[2 * j + 1] = a * si + b * co;
        }
      }
      if (res_row) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < e.N) v[j] += res_row[col0 + j];
      }
      const long long off = (long long)bz * e.c_bstride + (long long)row * e.ldc + col0;
      const bool full_chunk = col0 + 32 <= e.N;
      if (e.c_bf16) {
        bf16* out = reinterpret_cast<bf16*>(e.C) + off;
        if (full_chunk && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
          uint4* o4 = reinterpret_cast<uint4*>(out);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                               pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < e.N) out[j] = __float2bfloat16_rn(v[j]);
        }
      } else {
        float* out = reinterpret_cast<float*>(e.C) + off;
        if (full_chunk && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
          float4* o4 = reinterpret_cast<float4*>(out);
#pragma unroll
          for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < e.N) out[j] = v[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, BN);
}

template <int BN>
int launch_bn(const GemmArgs& a, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  const bool a_batched = a.a_bstride != 0 && a.batch > 1;
  VLS_TRY(make_tmap_bf16(&tmA, a.A, a.K, a.M, a_batched ? a.batch : 1, a.lda, a.a_bstride, BM));
  const bool w_batched = a.w_bstride != 0 && a.batch > 1;
  VLS_TRY(make_tmap_bf16(&tmB, a.W, a.K, a.N, w_batched ? a.batch : 1, a.ldw, a.w_bstride, BN));
  Epi e;
  e.M = a.M; e.N = a.N; e.K = a.K;
  e.bias = a.bias; e.bias_mode = a.bias ? a.bias_mode : 0; e.act = a.act;
  e.rope_cos = a.rope_cos; e.rope_sin = a.rope_sin; e.rope_period = a.rope_period > 0 ? a.rope_period : 1;
  e.rope_rows = a.rope_rows;
  e.residual = a.residual; e.ld_res = a.ld_res; e.res_bstride = a.res_bstride;
  e.C = a.C; e.c_bf16 = a.c_bf16; e.ldc = a.ldc; e.c_bstride = a.c_bstride;
  e.w_batched = w_batched ? 1 : 0;
  e.a_batched = a_batched ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    VLS_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN>()));
    attr_set = true;
  }
  dim3 grid((a.M + BM - 1) / BM, (a.N + BN - 1) / BN, a.batch);
  gemm_tn_kernel<BN><<<grid, THREADS, smem_bytes<BN>(), stream>>>(tmA, tmB, e);
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  VLS_REQUIRE(a.A && a.W && a.C, "gemm: null operand");
  VLS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && a.batch > 0, "gemm: bad shape M=%d N=%d K=%d batch=%d", a.M, a.N,
              a.K, a.batch);
  VLS_REQUIRE(a.lda % 8 == 0 && a.ldw % 8 == 0, "gemm: lda/ldw must be multiples of 8 elements");
  VLS_REQUIRE(!a.rope_cos || (a.rope_sin && a.rope_period > 0), "gemm: incomplete RoPE arguments");
  const long long tiles128 = (long long)((a.M + 127) / 128) * ((a.N + 127) / 128) * a.batch;
  if (tiles128 < 120 || a.N <= 64) return launch_bn<64>(a, stream);
  return launch_bn<128>(a, stream);
}

}  // namespace vls
