// The middle of a memory-attention layer in ONE cluster kernel (memory_attention.py:58-72):
//
//     x      = x + ao Wo^T + bo                 self-attention output projection + residual      (K = 256)
//     t      = LayerNorm2(x)
//     q      = RoPE(t Wq^T + bq)                cross-attention query projection, axial RoPE     (K = 256)
//
// r2 ran this as GEMM (7.8 us) -> LayerNorm (3.1) -> GEMM + RoPE epilogue (7.3) with the gaps between: three latency chains
// for 1.1 GFLOP.  Here a CLUSTER of 4 CTAs owns a 128-row tile and CTA r owns output columns [64 r, 64 r + 64) of BOTH
// projections, so a CTA streams only a quarter of each weight matrix (2 x 32 KB) next to the ao tile (64 KB) and its
// quarter of the f32 residual tile (32 KB, staged by TMA):
//   * Y0[128][64] = ao_tile Wo_r^T: 16 SS MMAs (M128 N64 K16) into tensor memory
//   * epilogue 1 (thread = row): + bias + residual -> x (written back, this CTA's 64 columns) kept in registers; the per-row
//     (sum, sum of squares) of the 64 columns are pushed to all four CTAs over distributed shared memory
//   * after a cluster barrier every CTA has the four partial statistics of its rows: t = LN2(x) for its columns is written as
//     bf16 into panel r of the 128B-swizzled A operand of ALL four CTAs (the 64 columns of a CTA are exactly one 64-wide
//     K panel), i.e. the all-gather of t is 8 remote 16-byte stores per row and destination
//   * D1[128][64] = t_tile Wq_r^T: 16 MMAs; epilogue 2 adds the bias, rotates the column pairs (position_encoding.py:168-222)
//     and stores bf16 q rows.
// TMEM: Y0 64 columns | D1 64.  SMEM: 64 + 32 + 32 + 32 KB operands / staging + 4 KB statistics.
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int C = 256;
constexpr int BM = 128;
constexpr int CL = 4;
constexpr int NS = 64;                       // output columns per CTA
constexpr int A_BYTES = BM * C * 2;          // 4 panels [128 rows x 64 cols]
constexpr int W_BYTES = NS * C * 2;          // 4 panels [64 rows x 64 cols]
constexpr int X_BYTES = BM * NS * 4;         // 2 boxes [128 rows x 32 floats]
constexpr int ST_BYTES = CL * BM * 8;        // [source rank][row] (sum, sum of squares)
constexpr int PRM_BYTES = 4 * NS * 4;        // this CTA's slices of bo, LN weight, LN bias, bq
constexpr int SMEM_BYTES = A_BYTES + 2 * W_BYTES + X_BYTES + ST_BYTES + PRM_BYTES + 256 + 1024;
constexpr int THREADS = 192;                 // warp 0: TMA, warp 1: MMA + TMEM, warps 2-5: epilogue (thread = row)
constexpr uint32_t TM_Y0 = 0, TM_D1 = 64;

struct MidParams {
  int M;
  const float* b0;           // [256] out-proj bias
  const float* ln_w; const float* ln_b; float ln_eps;
  const float* b1;           // [256] q-proj bias
  float* x;                  // f32 [B][M][256] residual stream, updated in place
  long long x_bstride;
  const float* rope_cos; const float* rope_sin; int rope_period;   // [period][128]
  bf16* q; long long ldq, q_bstride;
  long long* trace;          // optional dev trace (vls_ffn_trace buffer): clock64 stamps of the first epilogue thread of CTA (0,0,0)
};

#define MID_TRACE(slot)                                                                              \
  do {                                                                                               \
    if (p.trace && threadIdx.x == 64 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)       \
      p.trace[slot] = clock64();                                                                     \
  } while (0)

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
mid_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW0,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmQ, const MidParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                          // ao tile, then the t tile (all-gathered)
  uint8_t* sW0 = sA + A_BYTES;
  uint8_t* sW1 = sW0 + W_BYTES;
  uint8_t* sX = sW1 + W_BYTES;
  float2* stats = reinterpret_cast<float2*>(sX + X_BYTES);
  float* prm = reinterpret_cast<float*>(sX + X_BYTES + ST_BYTES);   // [bo | ln_w | ln_b | bq] x 64
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + X_BYTES + ST_BYTES + PRM_BYTES);
  uint64_t* g0_full = bars;        // ao tile + Wo slice landed
  uint64_t* x_full = bars + 1;     // residual slice landed
  uint64_t* w1_full = bars + 2;    // Wq slice landed
  uint64_t* g0_done = bars + 3;    // out-proj MMAs complete
  uint64_t* d1_full = bars + 4;    // q-proj MMAs complete
  uint64_t* t_full = bars + 5;     // the three remote panels of the t tile have landed (bulk DSMEM copies, complete_tx)
  uint64_t* a_ready = bars + 6;    // this CTA's own panel of t has been written (128 epilogue threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = blockIdx.y * BM, bz = blockIdx.z;
  const int n0 = (int)rank * NS;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_init(a_ready, 128);
    mbar_expect_tx(t_full, (CL - 1) * BM * 128);   // armed before any peer can send (they send after the first cluster barrier)
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW0);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmQ);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_enter();
  MID_TRACE(0);
  if (warp >= 2) {   // parameter slices -> shared memory (a dependent global load costs ~330 cycles each time it is needed)
    const int i = threadIdx.x - 64;   // 0..127: two floats each of the 4 x 64 block
    const int which = i >> 5, c2 = (i & 31) * 2;
    const float* src = which == 0 ? p.b0 : which == 1 ? p.ln_w : which == 2 ? p.ln_b : p.b1;
    *reinterpret_cast<float2*>(&prm[which * NS + c2]) = __ldg(reinterpret_cast<const float2*>(src + n0 + c2));
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(g0_full, A_BYTES + W_BYTES);
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        tma_load_3d(sA + kp * (BM * 128), &tmA, g0_full, kp * 64, m0, bz);
        tma_load_3d(sW0 + kp * (NS * 128), &tmW0, g0_full, kp * 64, n0, 0);
      }
      mbar_expect_tx(x_full, X_BYTES);
#pragma unroll
      for (int k = 0; k < 2; ++k) tma_load_3d(sX + k * (BM * 128), &tmX, x_full, n0 + 32 * k, m0, bz);
      mbar_expect_tx(w1_full, W_BYTES);
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) tma_load_3d(sW1 + kp * (NS * 128), &tmW1, w1_full, kp * 64, n0, 0);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, NS);
      mbar_wait(g0_full, 0);
      tc_fence_after();
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        const uint64_t ad = make_desc_sw128(smem_u32(sA + kp * (BM * 128)));
        const uint64_t bd = make_desc_sw128(smem_u32(sW0 + kp * (NS * 128)));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + TM_Y0, ad + 2 * kk, bd + 2 * kk, idesc, (kp | kk) != 0 ? 1u : 0u);
      }
      umma_commit(g0_done);
    }
  }

  // ---- epilogue 1: x = residual + Y0 + bias (this CTA's 64 columns of the rows), per-row partial statistics
  const int q4 = warp & 3, rl = q4 * 32 + lane;
  const uint32_t lane_off = uint32_t(q4 * 32) << 16;
  const int row = m0 + rl;
  const bool row_ok = row < p.M;
  float xm[NS];
  if (warp >= 2) {
    mbar_wait(g0_done, 0);
    MID_TRACE(1);
    mbar_wait(x_full, 0);
    tc_fence_after();
    MID_TRACE(2);
    float sum = 0.f, ss = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_off + TM_Y0 + k * 32, r);
      uint8_t* xs = sX + k * (BM * 128) + rl * 128;
      float4 xv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + ((i ^ (rl & 7)) << 4));
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bb = *reinterpret_cast<const float4*>(&prm[k * 32 + 4 * i]);
        float4 m;
        m.x = __uint_as_float(r[4 * i]) + bb.x + xv[i].x; m.y = __uint_as_float(r[4 * i + 1]) + bb.y + xv[i].y;
        m.z = __uint_as_float(r[4 * i + 2]) + bb.z + xv[i].z; m.w = __uint_as_float(r[4 * i + 3]) + bb.w + xv[i].w;
        sum += (m.x + m.y) + (m.z + m.w);
        ss += (m.x * m.x + m.y * m.y) + (m.z * m.z + m.w * m.w);
        xm[k * 32 + 4 * i] = m.x; xm[k * 32 + 4 * i + 1] = m.y; xm[k * 32 + 4 * i + 2] = m.z; xm[k * 32 + 4 * i + 3] = m.w;
        // the new residual row goes back into the staging tile (same swizzled slot) and out through a TMA store: thread = row
        // global stores touch 32 cache lines per warp instruction (2.8 k cycles for this epilogue)
        *reinterpret_cast<float4*>(xs + ((i ^ (rl & 7)) << 4)) = m;
      }
    }
    fence_proxy_async();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 64) {
#pragma unroll
      for (int k = 0; k < 2; ++k) tma_store_3d(sX + k * (BM * 128), &tmX, n0 + 32 * k, m0, bz);
      tma_store_commit();
    }
    const uint32_t dst = smem_u32(&stats[rank * BM + rl]);
#pragma unroll
    for (int r = 0; r < CL; ++r)
      asm volatile("st.shared::cluster.v2.f32 [%0], {%1,%2};" ::"r"(mapa_u32(dst, (uint32_t)r)), "f"(sum), "f"(ss) : "memory");
  }
  MID_TRACE(3);
  tc_fence_before();
  cluster_sync_all();   // statistics of all four column quarters are here; every CTA's out-proj MMAs have read its ao tile

  MID_TRACE(4);
  // ---- LayerNorm2 of this CTA's columns -> bf16 panel `rank` of the t operand in ALL four CTAs
  if (warp >= 2) {
    float sum = 0.f, ss = 0.f;
#pragma unroll
    for (int r = 0; r < CL; ++r) { const float2 s2 = stats[r * BM + rl]; sum += s2.x; ss += s2.y; }   // same order everywhere
    const float mean = sum * (1.0f / C);
    const float rstd = rsqrtf(fmaxf(ss * (1.0f / C) - mean * mean, 0.f) + p.ln_eps);
    uint8_t* prow = sA + rank * (BM * 128) + rl * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // 16-byte chunk j of the 128-byte panel row = columns 8 j .. 8 j + 7 of this CTA's slice
      const float4 wa = *reinterpret_cast<const float4*>(&prm[NS + 8 * j]), wb = *reinterpret_cast<const float4*>(&prm[NS + 8 * j + 4]);
      const float4 ba = *reinterpret_cast<const float4*>(&prm[2 * NS + 8 * j]), bb = *reinterpret_cast<const float4*>(&prm[2 * NS + 8 * j + 4]);
      const uint32_t p0 = pack_bf16x2((xm[8 * j] - mean) * rstd * wa.x + ba.x, (xm[8 * j + 1] - mean) * rstd * wa.y + ba.y);
      const uint32_t p1 = pack_bf16x2((xm[8 * j + 2] - mean) * rstd * wa.z + ba.z, (xm[8 * j + 3] - mean) * rstd * wa.w + ba.w);
      const uint32_t p2 = pack_bf16x2((xm[8 * j + 4] - mean) * rstd * wb.x + bb.x, (xm[8 * j + 5] - mean) * rstd * wb.y + bb.y);
      const uint32_t p3 = pack_bf16x2((xm[8 * j + 6] - mean) * rstd * wb.z + bb.z, (xm[8 * j + 7] - mean) * rstd * wb.w + bb.w);
      *reinterpret_cast<uint4*>(prow + ((j ^ (rl & 7)) << 4)) = make_uint4(p0, p1, p2, p3);
    }
    fence_proxy_async();      // generic-proxy writes -> visible to the async proxy (tensor core reads, bulk copies)
    tc_fence_before();
    mbar_arrive(a_ready);
    asm volatile("bar.sync 1, 128;" ::: "memory");   // the whole local panel is written
    if (threadIdx.x == 64) {
      // all-gather of t: this CTA's 16 KB panel -> the same panel slot of the three peers, as bulk shared->shared::cluster
      // copies that complete on the DESTINATION's t_full barrier (128 scattered remote 16-byte stores per warp took 7 k cycles)
      const uint32_t src = smem_u32(sA + rank * (BM * 128)), bar = smem_u32(t_full);
#pragma unroll
      for (int d = 1; d < CL; ++d) {
        const uint32_t peer = (rank + d) % CL;
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(mapa_u32(src, peer)), "r"(src), "r"(BM * 128), "r"(mapa_u32(bar, peer)) : "memory");
      }
    }
  }
  MID_TRACE(5);

  if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, NS);
      mbar_wait(w1_full, 0);
      mbar_wait(a_ready, 0);
      mbar_wait(t_full, 0);
      tc_fence_after();
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        const uint64_t ad = make_desc_sw128(smem_u32(sA + kp * (BM * 128)));
        const uint64_t bd = make_desc_sw128(smem_u32(sW1 + kp * (NS * 128)));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + TM_D1, ad + 2 * kk, bd + 2 * kk, idesc, (kp | kk) != 0 ? 1u : 0u);
      }
      umma_commit(d1_full);
    }
  } else if (warp >= 2) {
    // ---- epilogue 2: q = RoPE(D1 + bias): pairs (2i, 2i+1) rotate by the angle of (row % period, pair)
    const int rmod = row % p.rope_period;
    const float* cs = p.rope_cos + (long long)rmod * 128 + (n0 >> 1);
    const float* sn = p.rope_sin + (long long)rmod * 128 + (n0 >> 1);
    float4 co[8], si[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {   // 32 pairs: requested before the accumulator is ready
      co[i] = __ldg(reinterpret_cast<const float4*>(cs + 4 * i));
      si[i] = __ldg(reinterpret_cast<const float4*>(sn + 4 * i));
    }
    mbar_wait(d1_full, 0);
    tc_fence_after();
    MID_TRACE(7);
    // q rows are staged in a dead panel of the t tile (a peer's panel: consumed by this CTA's MMAs, never read remotely) in
    // the 128B-swizzled layout of a TMA box and stored by the TMA engine
    uint8_t* qrow = sA + ((rank + 1) & 3) * (BM * 128) + rl * 128;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_off + TM_D1 + k * 32, r);
      tc_wait_ld();
      uint32_t o[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {   // 4 columns = 2 pairs per step
        const float4 bb = *reinterpret_cast<const float4*>(&prm[3 * NS + k * 32 + 4 * i]);
        const float a0 = __uint_as_float(r[4 * i]) + bb.x, b0 = __uint_as_float(r[4 * i + 1]) + bb.y;
        const float a1 = __uint_as_float(r[4 * i + 2]) + bb.z, b1 = __uint_as_float(r[4 * i + 3]) + bb.w;
        const int pi = k * 16 + 2 * i;                       // pair index inside this CTA's 32 pairs
        const float c0 = reinterpret_cast<const float*>(co)[pi], s0 = reinterpret_cast<const float*>(si)[pi];
        const float c1 = reinterpret_cast<const float*>(co)[pi + 1], s1 = reinterpret_cast<const float*>(si)[pi + 1];
        o[2 * i] = pack_bf16x2(a0 * c0 - b0 * s0, a0 * s0 + b0 * c0);
        o[2 * i + 1] = pack_bf16x2(a1 * c1 - b1 * s1, a1 * s1 + b1 * c1);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(qrow + (((4 * k + i) ^ (rl & 7)) << 4)) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
    fence_proxy_async();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 64) {
      tma_store_3d(sA + ((rank + 1) & 3) * (BM * 128), &tmQ, n0, m0, bz);
      tma_store_commit();
      tma_store_wait_read();   // both store groups have read their shared-memory sources
    }
  }
  MID_TRACE(8);
  tc_fence_before();
  cluster_sync_all();   // peers may still be storing into this CTA's shared memory / reading its statistics
  MID_TRACE(9);
  if (warp == 1) tmem_dealloc(tmem, 128);
}

}  // namespace

int g_mid_fused = 1;   // memory attention: 1 = self-attention out-proj + LN2 + cross-attention q-proj (+RoPE) in one launch

// x += ao Wo^T + bo;  q = RoPE(LN(x) Wq^T + bq).  ao bf16 [B][M][256]; wo, wq bf16 [256][256]; x f32 [B][M][256] in place;
// q bf16 rows at b*q_bstride + m*ldq; RoPE tables f32 [period][128].
int launch_mid_fused(const MidArgs& a, cudaStream_t stream) {
  VLS_REQUIRE(a.ao && a.wo && a.bo && a.ln_w && a.ln_b && a.wq && a.bq && a.x && a.q && a.rope_cos && a.rope_sin && a.B > 0 && a.M > 0 &&
              a.rope_period > 0, "mid_fused: bad arguments");
  VLS_REQUIRE(a.ldq % 8 == 0 && a.q_bstride % 8 == 0, "mid_fused: q strides must be multiples of 8");
  CUtensorMap tmA, tmW0, tmW1, tmX, tmQ;
  VLS_TRY(make_tmap_bf16(&tmA, a.ao, C, a.M, a.B, C, (long long)a.M * C, BM));
  VLS_TRY(make_tmap_bf16(&tmW0, a.wo, C, C, 1, C, (long long)C * C, NS));
  VLS_TRY(make_tmap_bf16(&tmW1, a.wq, C, C, 1, C, (long long)C * C, NS));
  VLS_TRY(make_tmap_f32(&tmX, a.x, C, a.M, a.B, C, (long long)a.M * C, BM));
  VLS_TRY(make_tmap_bf16(&tmQ, a.q, C, a.M, a.B, a.ldq, a.q_bstride, BM));
  MidParams p = {};
  p.M = a.M; p.b0 = a.bo; p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.ln_eps = a.ln_eps; p.b1 = a.bq;
  p.x = a.x; p.x_bstride = (long long)a.M * C;
  p.rope_cos = a.rope_cos; p.rope_sin = a.rope_sin; p.rope_period = a.rope_period;
  p.q = reinterpret_cast<bf16*>(a.q); p.ldq = a.ldq; p.q_bstride = a.q_bstride;
  p.trace = g_ffn_trace;
  static unsigned long long attr_set = 0;
  if (first_use_on_device(&attr_set)) VLS_CUDA(cudaFuncSetAttribute(mid_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  VLS_CUDA(launch_k(mid_fused_kernel, dim3(CL, (a.M + BM - 1) / BM, a.B), dim3(THREADS), SMEM_BYTES, stream, tmA, tmW0, tmW1, tmX, tmQ, p));
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
