// Image side of a two-way-transformer layer of the mask decoder in ONE cluster kernel (sam/transformer.py:205-214 of
// layer l, and the image-side projections of what follows: :193-198 / :205 of layer l + 1, or :127-132):
//
//     a      = softmax(q_i2t k_tok^T / 4) v_tok            image -> token attention, 8 heads x 16, <= 16 tokens per row
//     keys   = LayerNorm4(keys + a Wo^T + bo)              (f32 residual stream in place + bf16 copy for the GEMMs)
//     planes = keys_bf16 Wn^T + bn + pe_add                next [K_t2i | V_t2i | Q_i2t] (or final [K | V]) as head planes
//
// r2 ran this as i2t_attn (3.9 us) -> GEMM (6.8) -> LayerNorm (3.5) -> GEMM (8.7): four launches of per-row work.  A cluster
// of 4 CTAs owns a 128-row tile (the pattern of mid_fused.cu): every CTA computes the whole attention tile itself (thread =
// (row, 4 heads): 1 k FMAs, written as the 128B-swizzled A operand: 4 heads x 16 channels are exactly one panel row), CTA r
// owns columns [64 r, 64 r + 64) of the output projection and LayerNorm (statistics over DSMEM, t all-gathered with bulk
// shared -> shared::cluster copies) and a quarter of the next projection's columns (96 or 64: whole 16-column planes).
// TMEM: Y0 64 columns | D1 96.
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int C = 256;
constexpr int BM = 128;
constexpr int CL = 4;
constexpr int NS = 64;                       // output-projection / LayerNorm columns per CTA
constexpr int NQ_MAX = 96;                   // next-projection columns per CTA (384 / 4; 64 for the final projection)
constexpr int TOK_MAX = 16;
constexpr int A_BYTES = BM * 128 * 2;        // attention tile: 2 panels [128 rows x 64 cols] bf16
constexpr int T_BYTES = BM * C * 2;          // t tile: 4 panels
constexpr int W0_BYTES = NS * 128 * 2;       // Wo slice [64 rows x 128]: 2 panels
constexpr int W1_BYTES = NQ_MAX * C * 2;     // Wn slice [96 rows x 256]: 4 panels of [96 x 64]
constexpr int X_BYTES = BM * NS * 4;         // residual slice f32: 2 boxes [128 rows x 32 floats]
constexpr int KV_BYTES = 2 * TOK_MAX * 128 * 4;
constexpr int ST_BYTES = 2 * CL * BM * 8;     // [source rank][column half][row] (sum, sum of squares)
constexpr int PRM_FLOATS = 3 * NS + NQ_MAX;  // bo | ln_w | ln_b | bn slices
constexpr int SMEM_BYTES = A_BYTES + T_BYTES + W0_BYTES + W1_BYTES + X_BYTES + KV_BYTES + ST_BYTES + PRM_FLOATS * 4 + 256 + 1024;
static_assert(SMEM_BYTES <= 232448, "dec_img: shared memory");
constexpr int THREADS = 320;                 // warp 0: TMA, warp 1: MMA + TMEM, warps 2-9: epilogue (two warps per TMEM lane quarter)
constexpr uint32_t TM_Y0 = 0, TM_D1 = 64;

struct DecImgParams {
  int T, Nt, nq;                   // rows per batch element, token rows, next-projection columns per CTA (96 or 64; 0 = none)
  const bf16* planes_in;           // bf16 [B][planes][T][16]: q of head h = plane qplane + h
  long long planes_in_bstride; int qplane;
  const float* kt; const float* vt;   // f32 [B][Nt][128]
  const float* b0;                 // [256]
  const float* ln_w; const float* ln_b; float ln_eps;
  float* keys; long long keys_bstride;            // f32 [B][T][256] in place
  bf16* keys_h;                                   // bf16 [B][T][256]
  const float* b1;                 // [4 nq]
  const float* pe_add;             // f32 [T][4 nq]
  bf16* planes_out; long long planes_out_bstride; // bf16 [B][4 nq / 16][T][16]
  long long* trace;                // optional dev trace (vls_ffn_trace buffer): clock64 stamps of thread 64 of CTA (0,0,0)
};

#define DI_TRACE(slot)                                                                               \
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
dec_img_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH, const DecImgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sT = sA + A_BYTES;
  uint8_t* sW0 = sT + T_BYTES;
  uint8_t* sW1 = sW0 + W0_BYTES;
  uint8_t* sX = sW1 + W1_BYTES;
  float* skv = reinterpret_cast<float*>(sX + X_BYTES);           // k_tok [Nt][128], v_tok [Nt][128]
  float2* stats = reinterpret_cast<float2*>(sX + X_BYTES + KV_BYTES);
  float* prm = reinterpret_cast<float*>(sX + X_BYTES + KV_BYTES + ST_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(prm + PRM_FLOATS);
  uint64_t* w0_full = bars;        // Wo slice landed
  uint64_t* x_full = bars + 1;     // residual slice landed
  uint64_t* w1_full = bars + 2;    // Wn slice landed
  uint64_t* g0_done = bars + 3;    // out-proj MMAs complete
  uint64_t* d1_full = bars + 4;    // next-projection MMAs complete
  uint64_t* t_full = bars + 5;     // three remote panels of t landed
  uint64_t* a_ready = bars + 6;    // attention tile written (256 threads)
  uint64_t* t_ready = bars + 7;    // own panel of t written (256 threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = blockIdx.y * BM, bz = blockIdx.z;
  const int n0 = (int)rank * NS;
  const int nq = p.nq, c0q = (int)rank * nq;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_init(a_ready, 256);
    mbar_init(t_ready, 256);
    mbar_expect_tx(t_full, (CL - 1) * BM * 128);
    fence_barrier_init();
    tma_prefetch_desc(&tmW0);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmH);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_enter();
  DI_TRACE(0);

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w0_full, W0_BYTES);
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) tma_load_3d(sW0 + kp * (NS * 128), &tmW0, w0_full, kp * 64, n0, 0);
      mbar_expect_tx(x_full, X_BYTES);
#pragma unroll
      for (int k = 0; k < 2; ++k) tma_load_3d(sX + k * (BM * 128), &tmX, x_full, n0 + 32 * k, m0, bz);
      if (nq > 0) {
        mbar_expect_tx(w1_full, nq * C * 2);
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) tma_load_3d(sW1 + kp * (nq * 128), &tmW1, w1_full, kp * 64, c0q, 0);
      }
    }
  }

  // epilogue thread = (row, column half): warps 2-5 take half 0, warps 6-9 half 1 of the same TMEM lane quarter
  const int et = threadIdx.x - 64;
  const int q4 = warp & 3, hh = warp >= 6 ? 1 : 0, rl = q4 * 32 + lane;
  const uint32_t lane_off = uint32_t(q4 * 32) << 16;
  const int row = m0 + rl;
  const bool row_ok = row < p.T;
  float xm[32];   // this thread's 32 columns of the new keys row: n0 + 32 hh + [0, 32)
  if (warp >= 2) {
    // ---- parameters and token k / v -> shared memory
    for (int i = et; i < PRM_FLOATS; i += 256) {
      float v;
      if (i < NS) v = __ldg(p.b0 + n0 + i);
      else if (i < 2 * NS) v = __ldg(p.ln_w + n0 + i - NS);
      else if (i < 3 * NS) v = __ldg(p.ln_b + n0 + i - 2 * NS);
      else v = (nq > 0 && i - 3 * NS < nq) ? __ldg(p.b1 + c0q + i - 3 * NS) : 0.f;
      prm[i] = v;
    }
    for (int i = et; i < p.Nt * 128; i += 256) {
      skv[i] = __ldg(p.kt + (long long)bz * p.Nt * 128 + i);
      skv[TOK_MAX * 128 + i] = __ldg(p.vt + (long long)bz * p.Nt * 128 + i);
    }
    // ---- image -> token attention: CTA r computes heads 2 r, 2 r + 1 (thread = (row, one head)) and stores its 32-byte piece
    //      of the attention row into the operand tile of ALL four CTAs (computing all 8 heads in every CTA was 9.6 k cycles:
    //      2.2 k instructions per thread on two warps per scheduler)
    const int head = 2 * (int)rank + hh;
    const uint4* qp = reinterpret_cast<const uint4*>(p.planes_in + (long long)bz * p.planes_in_bstride +
                                                     ((long long)(p.qplane + head) * p.T + (row_ok ? row : 0)) * 16);
    const uint4 q0 = __ldg(qp), q1 = __ldg(qp + 1);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    DI_TRACE(1);
    {
      const uint32_t qu[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      float qv[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(&qu[c]);
        qv[2 * c] = __low2float(q2) * 0.25f;        // 1 / sqrt(16)
        qv[2 * c + 1] = __high2float(q2) * 0.25f;
      }
      float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = 0.f;
      uint32_t ka = smem_u32(skv) + (head * 16) * 4;
#pragma unroll 1
      for (int j = 0; j < p.Nt; ++j, ka += 128 * 4) {
        float4 k4[4], v4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(k4[c].x), "=f"(k4[c].y), "=f"(k4[c].z), "=f"(k4[c].w) : "r"(ka + 16 * c));
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v4[c].x), "=f"(v4[c].y), "=f"(v4[c].z), "=f"(v4[c].w) : "r"(ka + TOK_MAX * 128 * 4 + 16 * c));
        }
        const float s0 = qv[0] * k4[0].x + qv[1] * k4[0].y + qv[2] * k4[0].z + qv[3] * k4[0].w;
        const float s1 = qv[4] * k4[1].x + qv[5] * k4[1].y + qv[6] * k4[1].z + qv[7] * k4[1].w;
        const float s2 = qv[8] * k4[2].x + qv[9] * k4[2].y + qv[10] * k4[2].z + qv[11] * k4[2].w;
        const float s3 = qv[12] * k4[3].x + qv[13] * k4[3].y + qv[14] * k4[3].z + qv[15] * k4[3].w;
        const float sj = (s0 + s1) + (s2 + s3);
        const float mn = fmaxf(m, sj);
        const float corr = __expf(m - mn), pr = __expf(sj - mn);
        l = l * corr + pr;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[4 * c] = acc[4 * c] * corr + pr * v4[c].x; acc[4 * c + 1] = acc[4 * c + 1] * corr + pr * v4[c].y;
          acc[4 * c + 2] = acc[4 * c + 2] * corr + pr * v4[c].z; acc[4 * c + 3] = acc[4 * c + 3] * corr + pr * v4[c].w;
        }
        m = mn;
      }
      const float inv = 1.0f / l;
      // panel head / 4, 16-byte chunks 2 (head % 4) + {0, 1} of this row
      const uint32_t arow = smem_u32(sA + (head >> 2) * (BM * 128) + rl * 128);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const uint32_t v0 = pack_bf16x2(acc[8 * c2] * inv, acc[8 * c2 + 1] * inv), v1 = pack_bf16x2(acc[8 * c2 + 2] * inv, acc[8 * c2 + 3] * inv);
        const uint32_t v2 = pack_bf16x2(acc[8 * c2 + 4] * inv, acc[8 * c2 + 5] * inv), v3 = pack_bf16x2(acc[8 * c2 + 6] * inv, acc[8 * c2 + 7] * inv);
        const uint32_t a = arow + (((2 * (head & 3) + c2) ^ (rl & 7)) << 4);
#pragma unroll
        for (int r = 0; r < CL; ++r)
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(mapa_u32(a, (uint32_t)r)), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy (remote) writes -> the tensor cores' async-proxy reads
  }
  DI_TRACE(2);
  tc_fence_before();
  cluster_sync_all();   // the attention tile is complete in every CTA
  if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, NS);
      mbar_wait(w0_full, 0);
      tc_fence_after();
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        const uint64_t ad = make_desc_sw128(smem_u32(sA + kp * (BM * 128)));
        const uint64_t bd = make_desc_sw128(smem_u32(sW0 + kp * (NS * 128)));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + TM_Y0, ad + 2 * kk, bd + 2 * kk, idesc, (kp | kk) != 0 ? 1u : 0u);
      }
      umma_commit(g0_done);
    }
  }
  if (warp >= 2) {
    // ---- epilogue 1: keys = residual + Y0 + bias: columns n0 + 32 hh + [0, 32) of this row
    mbar_wait(g0_done, 0);
    mbar_wait(x_full, 0);
    tc_fence_after();
    DI_TRACE(3);
    float sum = 0.f, ss = 0.f;
    {
      uint32_t r[32];
      tmem_ld32(tmem + lane_off + TM_Y0 + hh * 32, r);
      const uint8_t* xs = sX + hh * (BM * 128) + rl * 128;
      float4 xv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + ((i ^ (rl & 7)) << 4));
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bb = *reinterpret_cast<const float4*>(&prm[hh * 32 + 4 * i]);
        float4 mm;
        mm.x = __uint_as_float(r[4 * i]) + bb.x + xv[i].x; mm.y = __uint_as_float(r[4 * i + 1]) + bb.y + xv[i].y;
        mm.z = __uint_as_float(r[4 * i + 2]) + bb.z + xv[i].z; mm.w = __uint_as_float(r[4 * i + 3]) + bb.w + xv[i].w;
        sum += (mm.x + mm.y) + (mm.z + mm.w);
        ss += (mm.x * mm.x + mm.y * mm.y) + (mm.z * mm.z + mm.w * mm.w);
        xm[4 * i] = mm.x; xm[4 * i + 1] = mm.y; xm[4 * i + 2] = mm.z; xm[4 * i + 3] = mm.w;
      }
    }
    // per-row partial statistics of this CTA's 64 columns = the two halves: slot [rank][half][row] in every CTA
    const uint32_t dst = smem_u32(&stats[(rank * 2 + hh) * BM + rl]);
#pragma unroll
    for (int r = 0; r < CL; ++r)
      asm volatile("st.shared::cluster.v2.f32 [%0], {%1,%2};" ::"r"(mapa_u32(dst, (uint32_t)r)), "f"(sum), "f"(ss) : "memory");
  }
  DI_TRACE(4);
  tc_fence_before();
  cluster_sync_all();
  DI_TRACE(5);

  if (warp >= 2) {
    float sum = 0.f, ss = 0.f;
#pragma unroll
    for (int r = 0; r < 2 * CL; ++r) { const float2 s2 = stats[r * BM + rl]; sum += s2.x; ss += s2.y; }   // same order everywhere
    const float mean = sum * (1.0f / C);
    const float rstd = rsqrtf(fmaxf(ss * (1.0f / C) - mean * mean, 0.f) + p.ln_eps);
    // keys = LayerNorm4(...): f32 in place (residual of the next layer), bf16 copy, and panel `rank` of the t operand
    uint8_t* prow = sT + rank * (BM * 128) + rl * 128;
    uint8_t* xs = sX + hh * (BM * 128) + rl * 128;   // the f32 rows go back into the residual staging tile (TMA box layout)
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // 16-byte chunk 4 hh + j of the panel row
      const float4 wa = *reinterpret_cast<const float4*>(&prm[NS + 32 * hh + 8 * j]), wb = *reinterpret_cast<const float4*>(&prm[NS + 32 * hh + 8 * j + 4]);
      const float4 ba = *reinterpret_cast<const float4*>(&prm[2 * NS + 32 * hh + 8 * j]), bb = *reinterpret_cast<const float4*>(&prm[2 * NS + 32 * hh + 8 * j + 4]);
      float y[8];
      y[0] = (xm[8 * j] - mean) * rstd * wa.x + ba.x; y[1] = (xm[8 * j + 1] - mean) * rstd * wa.y + ba.y;
      y[2] = (xm[8 * j + 2] - mean) * rstd * wa.z + ba.z; y[3] = (xm[8 * j + 3] - mean) * rstd * wa.w + ba.w;
      y[4] = (xm[8 * j + 4] - mean) * rstd * wb.x + bb.x; y[5] = (xm[8 * j + 5] - mean) * rstd * wb.y + bb.y;
      y[6] = (xm[8 * j + 6] - mean) * rstd * wb.z + bb.z; y[7] = (xm[8 * j + 7] - mean) * rstd * wb.w + bb.w;
      *reinterpret_cast<uint4*>(prow + (((4 * hh + j) ^ (rl & 7)) << 4)) =
          make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
      *reinterpret_cast<float4*>(xs + (((2 * j) ^ (rl & 7)) << 4)) = make_float4(y[0], y[1], y[2], y[3]);
      *reinterpret_cast<float4*>(xs + (((2 * j + 1) ^ (rl & 7)) << 4)) = make_float4(y[4], y[5], y[6], y[7]);
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(t_ready);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 64) {
      const uint32_t src = smem_u32(sT + rank * (BM * 128)), bar = smem_u32(t_full);
      if (nq > 0) {   // all-gather of t: own 16 KB panel -> the same slot of the three peers
#pragma unroll
        for (int d = 1; d < CL; ++d) {
          const uint32_t peer = (rank + d) % CL;
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(mapa_u32(src, peer)), "r"(src), "r"(BM * 128), "r"(mapa_u32(bar, peer)) : "memory");
        }
      }
      // keys: bf16 copy straight from the operand panel, f32 rows from the staging tile -- TMA stores (thread = row global
      // stores touch 32 cache lines per warp instruction: 12 of them per thread were most of this phase's 6-8 k cycles)
      tma_store_3d(sT + rank * (BM * 128), &tmH, n0, m0, bz);
#pragma unroll
      for (int k = 0; k < 2; ++k) tma_store_3d(sX + k * (BM * 128), &tmX, n0 + 32 * k, m0, bz);
      tma_store_commit();
    }
  }

  DI_TRACE(6);
  if (nq > 0) {
    if (warp == 1) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_bf16(BM, nq);
        mbar_wait(w1_full, 0);
        mbar_wait(t_ready, 0);
        mbar_wait(t_full, 0);
        tc_fence_after();
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) {
          const uint64_t ad = make_desc_sw128(smem_u32(sT + kp * (BM * 128)));
          const uint64_t bd = make_desc_sw128(smem_u32(sW1 + kp * (nq * 128)));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + TM_D1, ad + 2 * kk, bd + 2 * kk, idesc, (kp | kk) != 0 ? 1u : 0u);
        }
        umma_commit(d1_full);
      }
    } else if (warp >= 2) {
      // ---- epilogue 2: planes = D1 + bias + pe_add.  16-column planes of this CTA: nq / 16, split between the two halves
      const int npl = nq >> 4, pl0 = hh * (npl >> 1), pl1 = pl0 + (npl >> 1);
      const float* pe = p.pe_add + (long long)(row_ok ? row : 0) * (4 * nq) + c0q;
      float4 pv[3][4];   // up to 3 planes per half: requested before the accumulator is ready
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (pl0 + k < pl1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) pv[k][i] = __ldg(reinterpret_cast<const float4*>(pe + (pl0 + k) * 16 + 4 * i));
        }
      mbar_wait(d1_full, 0);
      tc_fence_after();
      DI_TRACE(7);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (pl0 + k >= pl1) break;
        const int pl = pl0 + k;
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + lane_off + TM_D1 + pl * 16) : "memory");
        tc_wait_ld();
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 bb = *reinterpret_cast<const float4*>(&prm[3 * NS + pl * 16 + 4 * i]);
          o[2 * i] = pack_bf16x2(__uint_as_float(r[4 * i]) + bb.x + pv[k][i].x, __uint_as_float(r[4 * i + 1]) + bb.y + pv[k][i].y);
          o[2 * i + 1] = pack_bf16x2(__uint_as_float(r[4 * i + 2]) + bb.z + pv[k][i].z, __uint_as_float(r[4 * i + 3]) + bb.w + pv[k][i].w);
        }
        // plane pl of this CTA: [128 rows][16] bf16 = 4 KB contiguous in global memory: staged in the dead attention tile
        uint4* stp = reinterpret_cast<uint4*>(sA + pl * (BM * 32) + rl * 32);
        stp[0] = make_uint4(o[0], o[1], o[2], o[3]);
        stp[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) {
        const int rows = min(BM, p.T - m0);
        for (int pl = 0; pl < npl; ++pl) {
          bf16* g = p.planes_out + (long long)bz * p.planes_out_bstride + ((long long)((c0q >> 4) + pl) * p.T + m0) * 16;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       ::"l"(g), "r"(smem_u32(sA + pl * (BM * 32))), "r"(rows * 32) : "memory");
        }
        tma_store_commit();
      }
    }
  }
  if (threadIdx.x == 64) tma_store_wait_read();   // every bulk store has read its shared-memory source
  DI_TRACE(8);
  tc_fence_before();
  cluster_sync_all();   // peers' bulk copies out of / into this CTA's shared memory have completed before anyone exits
  DI_TRACE(9);
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace

int g_dec_img_fused = 1;   // mask decoder: 1 = image side of a layer (i2t attention, out-proj, LN4, next projections) as one launch

bool dec_img_supported(int Nt, int T) { return g_dec_img_fused && Nt >= 1 && Nt <= TOK_MAX && T >= 1; }

int launch_dec_img(const DecImgArgs& a, cudaStream_t stream) {
  VLS_REQUIRE(a.planes_in && a.kt && a.vt && a.wo && a.bo && a.ln_w && a.ln_b && a.keys && a.keys_h && a.B > 0 && a.T > 0,
              "dec_img: bad arguments");
  VLS_REQUIRE(a.Nt >= 1 && a.Nt <= TOK_MAX, "dec_img: at most %d token rows", TOK_MAX);
  VLS_REQUIRE(a.n_next == 0 || a.n_next == 384 || a.n_next == 256, "dec_img: the next projection has 384 or 256 columns");
  VLS_REQUIRE(a.n_next == 0 || (a.wn && a.bn && a.pe_add && a.planes_out), "dec_img: next-projection arguments missing");
  const int nq = a.n_next / 4;
  CUtensorMap tmW0, tmW1, tmX, tmH;
  VLS_TRY(make_tmap_bf16(&tmW0, a.wo, 128, C, 1, 128, (long long)C * 128, NS));
  if (nq > 0) VLS_TRY(make_tmap_bf16(&tmW1, a.wn, C, a.n_next, 1, C, (long long)a.n_next * C, nq));
  else tmW1 = tmW0;
  VLS_TRY(make_tmap_f32(&tmX, a.keys, C, a.T, a.B, C, (long long)a.T * C, BM));
  VLS_TRY(make_tmap_bf16(&tmH, a.keys_h, C, a.T, a.B, C, (long long)a.T * C, BM));
  DecImgParams p = {};
  p.T = a.T; p.Nt = a.Nt; p.nq = nq;
  p.planes_in = reinterpret_cast<const bf16*>(a.planes_in); p.planes_in_bstride = a.planes_in_bstride; p.qplane = a.qplane;
  p.kt = a.kt; p.vt = a.vt; p.b0 = a.bo; p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.ln_eps = a.ln_eps;
  p.keys = a.keys; p.keys_bstride = (long long)a.T * C; p.keys_h = reinterpret_cast<bf16*>(a.keys_h);
  p.b1 = a.bn; p.pe_add = a.pe_add;
  p.planes_out = reinterpret_cast<bf16*>(a.planes_out); p.planes_out_bstride = a.planes_out_bstride;
  p.trace = g_ffn_trace;
  static unsigned long long attr_set = 0;
  if (first_use_on_device(&attr_set)) VLS_CUDA(cudaFuncSetAttribute(dec_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  VLS_CUDA(launch_k(dec_img_kernel, dim3(CL, (a.T + BM - 1) / BM, a.B), dim3(THREADS), SMEM_BYTES, stream, tmW0, tmW1, tmX, tmH, p));
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
