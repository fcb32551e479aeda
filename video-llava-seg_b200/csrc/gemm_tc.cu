// tcgen05 / TMEM / TMA GEMM with fused epilogues for every dense contraction on the path
// (q/k/v/out projections with axial RoPE, FFN, 1x1 convs, ConvTranspose-as-GEMM, pwconvs).
//
//   C[b][m][n] = epi( sum_k A[b][m][k] * W[n][k] )      A, W: bf16, K-major;  accumulate f32 in TMEM
//
// One CTA computes a 128 x BN tile.  Warp 0 = TMA producer (6/8-stage mbarrier ring = 192 KB of
// 128B-swizzled [128 x 64] / [BN x 64] boxes in flight), warp 1 = single-thread tcgen05.mma issuer +
// TMEM owner, warps 2-5 = epilogue: each warp drains its 32-lane TMEM quarter (tcgen05.ld 32x32b.x32)
// into an f32 staging tile that reuses the idle operand ring, then the 128 epilogue threads sweep the
// tile row-wise so that bias / ReLU / GELU / RoPE-table / residual reads and the stores are all
// coalesced 128-bit accesses.  blockIdx.z indexes (object, layer) batches with independent
// div/mod maps per operand, so e.g. the memory K/V projections of all 4 layers are one launch.
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace vls {

extern int g_gemm_ring2_above;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int THREADS = 192;
constexpr int A_BYTES = BM * BK * 2;

template <int BN, int STAGES_>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = STAGES_;                        // = K blocks for short K, 8/6 for long K
  static constexpr int PITCH = BN + 4;                          // f32 staging pitch (16 B aligned, conflict-free)
  static constexpr int RING = STAGES * STAGE_BYTES;
  static constexpr int STAGING = BM * PITCH * 4;                // epilogue tile, aliases the idle operand ring
  static constexpr int SMEM = (RING > STAGING ? RING : STAGING) + 256 + 1024;
  // resident CTAs per SM (smem-limited, capped at 3): one CTA's epilogue overlaps the others' loads/MMAs
  static constexpr int MINB = (232448 / SMEM) >= 3 ? 3 : ((232448 / SMEM) >= 2 ? 2 : 1);
  static constexpr int G = MINB >= 3 ? 4 : 8;                   // rows whose global loads are batched per thread
};

struct Epi {
  int M, N, K;
  const float* bias;
  int bias_mode, act;
  long long bias_bstride;
  int bias_div, bias_mod;
  const float* rope_cos;
  const float* rope_sin;
  int rope_period, rope_rows;
  const float* residual;
  long long ld_res, res_bstride;
  void* C;
  int c_bf16;
  long long ldc, c_bstride;
  int c_colblock;                   // > 0: column n lives at (n / colblock) * colblock_stride + n % colblock ("planes")
  long long c_colblock_stride;
  int a_div, a_mod, w_div, w_mod;   // operand batch index = (blockIdx.z / div) % mod
};

template <int BN, int STAGES_>
__global__ void __launch_bounds__(THREADS, Cfg<BN, STAGES_>::MINB)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Epi e) {
  using C_ = Cfg<BN, STAGES_>;
  constexpr int STAGES = C_::STAGES;
  constexpr int STAGE_BYTES = C_::STAGE_BYTES;
  constexpr int PITCH = C_::PITCH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (C_::RING > C_::STAGING ? C_::RING : C_::STAGING));
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
  pdl_enter();   // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail; global memory from here on

  if (warp == 0) {
    if (elect_one()) {   // elect.sync: a lane test makes the compiler wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop
      const int az = (bz / e.a_div) % e.a_mod;
      const int wz = (bz / e.w_div) % e.w_mod;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        uint8_t* sa = smem + s * STAGE_BYTES;
        tma_load_3d(sa, &tmA, &full[s], kb * BK, m0, az);
        tma_load_3d(sa + A_BYTES, &tmB, &full[s], kb * BK, n0, wz);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // elect.sync: a lane test makes the compiler wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop
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
    // ---- epilogue: TMEM -> registers -> f32 staging tile in the (now idle) operand ring -> coalesced pass
    const int q = warp & 3;          // TMEM lane quarter of this warp
    const int et = threadIdx.x - 64; // 0..127
    float* stage = reinterpret_cast<float*>(smem);
    mbar_wait(acc_full, 0);          // all MMAs done => every TMA load has landed and been consumed
    tc_fence_after();
    {
      float* srow = stage + (q * 32 + lane) * PITCH;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + (uint32_t(q * 32) << 16) + c * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(srow + c * 32 + 4 * j) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
    constexpr int CG = BN / 4;            // float4 column groups per row
    constexpr int RPP = 128 / CG;         // rows handled per pass by the 128 epilogue threads
    const int cg = et % CG, rsub = et / CG;
    const int col = n0 + cg * 4;
    const bool col_ok = col < e.N;        // N % 4 == 0 is required by the launcher
    const float* bias = e.bias ? e.bias + (long long)((bz / e.bias_div) % e.bias_mod) * e.bias_bstride : nullptr;
    float4 bcol = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias_mode == 1 && col_ok) bcol = *reinterpret_cast<const float4*>(bias + col);
    const bool rope = e.rope_cos != nullptr;
    const int act = e.act, bias_mode = e.bias_mode, c_bf16 = e.c_bf16;
    // Everything that depends on the row is a RUNNING value advanced by RPP rows per step: no 64-bit multiplies, no
    // integer modulo (row % rope_period) and no mode decoding inside the sweep.  The first ncu capture of the memory K
    // projection (M=28736, N=256, K=64) showed 15.4 k warp instructions per 128x128 tile, i.e. an instruction-bound
    // epilogue at 0.8 TB/s of output.
    const int row_first = m0 + rsub;
    const long long c_elem = c_bf16 ? 2 : 4;
    const long long col_off = e.c_colblock > 0 ? (long long)(col / e.c_colblock) * e.c_colblock_stride + col % e.c_colblock : col;
    char* cptr = reinterpret_cast<char*>(e.C) + ((long long)bz * e.c_bstride + (long long)row_first * e.ldc + col_off) * c_elem;
    const long long c_step = (long long)RPP * e.ldc * c_elem;
    const float* rptr = e.residual ? e.residual + (long long)bz * e.res_bstride + (long long)row_first * e.ld_res + col : nullptr;
    const long long r_step = (long long)RPP * e.ld_res;
    const float* brow_ptr = bias_mode == 2 ? bias + row_first : nullptr;
    int rmod = rope ? row_first % e.rope_period : 0;            // one modulo per thread, then incremental
    const int rope_inc = rope ? RPP % e.rope_period : 0;
    const int pair0 = (col & 255) >> 1;
    const float* srd = stage + rsub * PITCH + cg * 4;
    int row = row_first;
    constexpr int G = C_::G;  // rows per thread whose global loads are issued together (memory-level parallelism)
    // The sweep is instantiated per (GELU, RoPE): as run-time conditions the compiler if-converted them, and every float4
    // paid for ~100 predicated-off erf / rotation instructions (12 k warp instructions per 128x128 tile in ncu).
    auto sweep = [&](auto gelu_tag, auto rope_tag) {
    constexpr bool GELU = decltype(gelu_tag)::value;
    constexpr bool ROPE = decltype(rope_tag)::value;
#pragma unroll 1
    for (int r0 = 0; r0 < BM; r0 += RPP * G) {
      float4 v[G], rr[G];
      float2 co[G], si[G];
      float brow[G];
      bool ok[G], rot[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int rw = row + g * RPP;
        ok[g] = col_ok && rw < e.M;
        rot[g] = ROPE && ok[g] && rw < e.rope_rows;
        v[g] = *reinterpret_cast<const float4*>(srd + (r0 + g * RPP) * PITCH);
        if constexpr (ROPE) {
          if (rot[g]) {
            const int t = (rmod << 7) + pair0;
            co[g] = *reinterpret_cast<const float2*>(e.rope_cos + t);
            si[g] = *reinterpret_cast<const float2*>(e.rope_sin + t);
          }
          rmod += rope_inc;
          if (rmod >= e.rope_period) rmod -= e.rope_period;
        }
        if (rptr && ok[g]) rr[g] = *reinterpret_cast<const float4*>(rptr + g * r_step);
        brow[g] = (brow_ptr && ok[g]) ? __ldg(brow_ptr + g * RPP) : 0.f;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (!ok[g]) continue;
        float4 x = v[g];
        if (bias_mode == 1) {
          x.x += bcol.x; x.y += bcol.y; x.z += bcol.z; x.w += bcol.w;
        } else if (bias_mode == 2) {
          x.x += brow[g]; x.y += brow[g]; x.z += brow[g]; x.w += brow[g];
        }
        if constexpr (GELU) {
          x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w);
        } else if (act == 1) {
          x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
        }
        if constexpr (ROPE) {
          if (rot[g]) {
            const float a0 = x.x, b0 = x.y, a1 = x.z, b1 = x.w;
            x.x = a0 * co[g].x - b0 * si[g].x; x.y = a0 * si[g].x + b0 * co[g].x;
            x.z = a1 * co[g].y - b1 * si[g].y; x.w = a1 * si[g].y + b1 * co[g].y;
          }
        }
        if (rptr) {
          x.x += rr[g].x; x.y += rr[g].y; x.z += rr[g].z; x.w += rr[g].w;
        }
        char* dst = cptr + g * c_step;
        if (c_bf16)
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
        else
          *reinterpret_cast<float4*>(dst) = x;
      }
      row += RPP * G;
      cptr += G * c_step;
      if (rptr) rptr += G * r_step;
      if (brow_ptr) brow_ptr += G * RPP;
    }
    };
    if (act == 2) sweep(std::true_type{}, std::false_type{});       // GELU outputs are never rotated
    else if (rope) sweep(std::false_type{}, std::true_type{});
    else sweep(std::false_type{}, std::false_type{});
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, BN);
}

template <int BN, int STAGES_>
int launch_cfg(const GemmArgs& a, cudaStream_t stream) {
  using C_ = Cfg<BN, STAGES_>;
  CUtensorMap tmA, tmB;
  VLS_TRY(make_tmap_bf16(&tmA, a.A, a.K, a.M, a.a_batches, a.lda, a.a_bstride, BM));
  VLS_TRY(make_tmap_bf16(&tmB, a.W, a.K, a.N, a.w_batches, a.ldw, a.w_bstride, BN));
  Epi e;
  e.M = a.M; e.N = a.N; e.K = a.K;
  e.bias = a.bias; e.bias_mode = a.bias ? a.bias_mode : 0; e.act = a.act;
  e.bias_bstride = a.bias_bstride; e.bias_div = a.bias_div > 0 ? a.bias_div : 1; e.bias_mod = a.bias_batches > 0 ? a.bias_batches : 1;
  e.rope_cos = a.rope_cos; e.rope_sin = a.rope_sin; e.rope_period = a.rope_period > 0 ? a.rope_period : 1;
  e.rope_rows = a.rope_rows;
  e.residual = a.residual; e.ld_res = a.ld_res; e.res_bstride = a.res_bstride;
  e.C = a.C; e.c_bf16 = a.c_bf16; e.ldc = a.ldc; e.c_bstride = a.c_bstride;
  e.c_colblock = a.c_colblock; e.c_colblock_stride = a.c_colblock_stride;
  e.a_div = a.a_div; e.a_mod = a.a_batches; e.w_div = a.w_div; e.w_mod = a.w_batches;
  static unsigned long long attr_set = 0;
  if (first_use_on_device(&attr_set)) {
    VLS_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, STAGES_>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
  }
  dim3 grid((a.M + BM - 1) / BM, (a.N + BN - 1) / BN, a.batch);
  VLS_CUDA(launch_k(gemm_tn_kernel<BN, STAGES_>, dim3(grid), dim3(THREADS), C_::SMEM, stream, tmA, tmB, e));
  VLS_POST_LAUNCH(1);
  return 0;
}

template <int BN>
int launch_bn(const GemmArgs& a, cudaStream_t stream) {
  // short K: ring = exactly the K blocks (1/2/4 stages) so 2-3 CTAs stay resident per SM and one CTA's epilogue
  // overlaps the others' main loops; long K: 192 KB of loads in flight to cover the L2->smem latency
  const int kblocks = (a.K + BK - 1) / BK;
  const long long ctas = (long long)((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN) * a.batch;
  if (kblocks <= 1) return launch_cfg<BN, 1>(a, stream);
  // short K: a 2-stage ring keeps 3-4 CTAs resident per SM and lets a CTA start in the shared memory one retiring CTA of a
  // co-running launch frees (g_gemm_ring2_above)
  if (kblocks <= 2 || (kblocks <= 4 && ctas > g_gemm_ring2_above)) return launch_cfg<BN, 2>(a, stream);
  if (kblocks <= 4) return launch_cfg<BN, 4>(a, stream);
  return launch_cfg<BN, BN == 64 ? 8 : 6>(a, stream);
}

}  // namespace

int g_gemm_bn64_below = 296;   // vls_set_tuning("gemm_bn64_below")
// vls_set_tuning("gemm_ring2_above"): K <= 256 GEMMs with more CTAs than this use a 2-stage operand ring.  Was 2 x 148 (only
// multi-wave grids); 0 since the pipelined frame: the M = 4096 projections of the memory attention run next to the background key
// projection, whose three resident 69 KB CTAs per SM leave no room for a 97 KB four-stage CTA until two of them retire -- a 50 KB
// two-stage CTA starts when one does.  Frame 0.8865 -> 0.8752 ms (three runs each on one box; 100: 0.8763, 130: 0.8915).
int g_gemm_ring2_above = 0;

int launch_gemm(const GemmArgs& a_in, cudaStream_t stream) {
  GemmArgs a = a_in;
  VLS_REQUIRE(a.A && a.W && a.C, "gemm: null operand");
  VLS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && a.batch > 0, "gemm: bad shape M=%d N=%d K=%d batch=%d", a.M, a.N,
              a.K, a.batch);
  VLS_REQUIRE(a.lda % 8 == 0 && a.ldw % 8 == 0, "gemm: lda/ldw must be multiples of 8 elements");
  VLS_REQUIRE(a.N % 4 == 0 && a.ldc % 4 == 0, "gemm: N and ldc must be multiples of 4");
  VLS_REQUIRE(!a.residual || a.ld_res % 4 == 0, "gemm: ld_res must be a multiple of 4");
  VLS_REQUIRE(a.c_colblock == 0 || (a.c_colblock % 4 == 0 && a.c_colblock_stride % 4 == 0), "gemm: column blocks must be multiples of 4");
  VLS_REQUIRE(!a.rope_cos || (a.rope_sin && a.rope_period > 0), "gemm: incomplete RoPE arguments");
  VLS_REQUIRE(!(a.rope_cos && a.act == 2), "gemm: GELU + RoPE epilogue is not instantiated");
  // default batch indexing: operand batch = blockIdx.z when it has a batch stride, else shared
  if (a.a_batches <= 0) { a.a_batches = (a.a_bstride != 0 && a.batch > 1) ? a.batch : 1; a.a_div = 1; }
  if (a.w_batches <= 0) { a.w_batches = (a.w_bstride != 0 && a.batch > 1) ? a.batch : 1; a.w_div = 1; }
  if (a.a_div <= 0) a.a_div = 1;
  if (a.w_div <= 0) a.w_div = 1;
  const long long tiles128 = (long long)((a.M + 127) / 128) * ((a.N + 127) / 128) * a.batch;
  // up to two waves of 128-wide tiles (M = 4096 projections): 64-wide tiles put 2-3 CTAs on every SM, so one CTA's
  // epilogue runs under the others' loads and MMAs (the q/k projection + RoPE: 14.9 -> see DESIGN)
  if (tiles128 < g_gemm_bn64_below || a.N <= 64) return launch_bn<64>(a, stream);
  // (The memory K projection -- M = 28736, N = 256, K = 64 -- is 450 tiles on 148 SMs x 3 resident CTAs = 1.01 waves
  // (ncu: launch__waves_per_multiprocessor), i.e. the time of two.  900 tiles of 128 x 64 at 4 CTAs per SM = 1.52 half-size
  // waves were measured as well: the frame got 7 us SLOWER (0.9205 vs 0.913 ms, A/B on one box), so the rule stays.)
  return launch_bn<128>(a, stream);
}

}  // namespace vls
