// tcgen05 / TMEM flash attention for the memory-attention stack: one head of dim 256,
// Nq = 4096 queries, Nk up to 7*4096 + 64 keys (sam/transformer.py:311-360 after projection;
// RoPE is already applied to Q/K by the projection GEMM epilogue).
//
// CTA = one 128-query tile x one KV split.  Warp 0: TMA producer (Q once; K / V^T tiles of 64 keys,
// 2 stages each, 128B swizzle).  Warp 1: single-thread tcgen05.mma issuer:
//      S_j = Q K_j^T   (SS, M128 N64 K16 x16)  -> TMEM S[j&1]
//      O  += P_j V_j   (TS, A = bf16 P_j in TMEM aliasing S[j&1], B = V^T tile, M128 N256 K16 x4)
// S_{j+1} is issued before P_j is waited on, so the tensor pipe overlaps the softmax of tile j.
// Warps 2-5: softmax, one thread per query row (no shuffles), running max kept in the log2 domain
// and only refreshed when it grows by more than 8 (lazy rescale of O in TMEM), exp2 with the
// 1/sqrt(d)*log2(e) scale folded in, P written back to TMEM as packed bf16.
// With KV splits, partial (O, m, l) go to a workspace and attn_combine_kernel merges them.
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int D = 256;
constexpr int KV_STAGES = 2;
constexpr int Q_BYTES = BM * D * 2;
constexpr int K_BYTES = BN * D * 2;
constexpr int V_BYTES = D * BN * 2;
constexpr int THREADS = 192;
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + 256 + 1024;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_O = 0;
constexpr uint32_t TM_S = 256;
constexpr float RESCALE_THRESHOLD = 8.0f;

struct AttnParams {
  int Nq, Nk, splits;
  float scale_log2;
  bf16* O;
  long long ldo, o_bstride;
  float* part_o;
  float* part_ml;
};

__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + Q_BYTES;
  uint8_t* sV = sK + KV_STAGES * K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KV_STAGES * V_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 9;
  uint64_t* p_ready = bars + 11;
  uint64_t* pv_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM;
  const int split = blockIdx.y;
  const int bz = blockIdx.z;
  const int nt_all = (p.Nk + BN - 1) / BN;
  const int t0 = (int)((long long)nt_all * split / p.splits);
  const int t1 = (int)((long long)nt_all * (split + 1) / p.splits);
  const int n = t1 - t0;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 128);
    }
    mbar_init(pv_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) tma_load_3d(sQ + kp * (BM * 128), &tmQ, q_full, kp * 64, q0, bz);
      for (int j = 0; j < n; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int kv0 = (t0 + j) * BN;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], K_BYTES);
#pragma unroll
        for (int kp = 0; kp < 4; ++kp)
          tma_load_3d(sK + st * K_BYTES + kp * (BN * 128), &tmK, &k_full[st], kp * 64, kv0, bz);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], V_BYTES);
        tma_load_3d(sV + st * V_BYTES, &tmV, &v_full[st], kv0, 0, bz);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BM, D);
      const uint32_t q_addr = smem_u32(sQ);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * K_BYTES);
        const uint32_t d_s = tmem + TM_S + uint32_t(j & 1) * BN;
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) {
          const uint64_t qd = make_desc_sw128(q_addr + kp * (BM * 128));
          const uint64_t kd = make_desc_sw128(k_addr + kp * (BN * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(d_s, qd + 2 * kk, kd + 2 * kk, idesc_s, (kp | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[j & 1]);
      };
      issue_s(0);
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) issue_s(j + 1);
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&p_ready[st], ph);
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        const uint64_t vd = make_desc_sw128(smem_u32(sV + st * V_BYTES));
        const uint32_t a_p = tmem + TM_S + uint32_t(j & 1) * BN;
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks)
          umma_ts(tmem + TM_O, a_p + ks * 8, vd + 2 * ks, idesc_pv, (j | ks) != 0 ? 1u : 0u);
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q0 + q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    float m_used = -INFINITY;
    float l = 0.0f;
    for (int j = 0; j < n; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t r[64];
      tmem_ld32(tmem + lane_off + TM_S + b * BN, r);
      tmem_ld32(tmem + lane_off + TM_S + b * BN + 32, r + 32);
      tc_wait_ld();
      const int kv0 = (t0 + j) * BN;
      const int valid = p.Nk - kv0;  // >= 1
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float s = __uint_as_float(r[i]);
        if (i >= valid) s = -INFINITY;
        r[i] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
      const float m_new = fmaxf(m_used, mx * p.scale_log2);
      const bool need = m_new > m_used + RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        float alpha = 1.0f;
        if (need) {
          alpha = exp2f(m_used - m_new);
          m_used = m_new;
        }
        l *= alpha;
        if (j > 0) {
          mbar_wait(pv_done, (j - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem + lane_off + TM_O + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem + lane_off + TM_O + c * 32, o);
          }
          tc_wait_st();
        }
      }
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p0 = exp2f(__uint_as_float(r[2 * i]) * p.scale_log2 - m_used);
        const float p1 = exp2f(__uint_as_float(r[2 * i + 1]) * p.scale_log2 - m_used);
        l += p0 + p1;
        pk[i] = pack_bf16x2(p0, p1);
      }
      tmem_st32(tmem + lane_off + TM_S + b * BN, pk);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_ready[b]);
    }
    if (n > 0) {
      mbar_wait(pv_done, (n - 1) & 1);
      tc_fence_after();
    }
    const bool row_ok = row < p.Nq;
    if (p.splits == 1) {
      const float inv = l > 0.0f ? 1.0f / l : 0.0f;
      bf16* out = p.O + (long long)bz * p.o_bstride + (long long)row * p.ldo;
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(tmem + lane_off + TM_O + c * 32, o);
        tc_wait_ld();
        if (row_ok) {
          uint4* o4 = reinterpret_cast<uint4*>(out + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o4[i] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                               pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                               pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                               pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
        }
      }
    } else {
      const long long prow = ((long long)bz * p.splits + split) * p.Nq + row;
      float* po = p.part_o + prow * D;
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(tmem + lane_off + TM_O + c * 32, o);
        tc_wait_ld();
        if (row_ok) {
          float4* o4 = reinterpret_cast<float4*>(po + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o4[i] = make_float4(__uint_as_float(o[4 * i]), __uint_as_float(o[4 * i + 1]), __uint_as_float(o[4 * i + 2]),
                                __uint_as_float(o[4 * i + 3]));
        }
      }
      if (row_ok) {
        p.part_ml[prow * 2] = m_used;
        p.part_ml[prow * 2 + 1] = l;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// Merge KV-split partials: one warp per query row, lane owns 8 channels.
__global__ void attn_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, int B, int Nq,
                                    int splits, bf16* __restrict__ O, long long ldo, long long o_bstride) {
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (long long)B * Nq) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(gw / Nq);
  const int row = (int)(gw % Nq);
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s) m = fmaxf(m, part_ml[(((long long)b * splits + s) * Nq + row) * 2]);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float l = 0.0f;
  for (int s = 0; s < splits; ++s) {
    const long long prow = ((long long)b * splits + s) * Nq + row;
    const float w = exp2f(part_ml[prow * 2] - m);
    l += w * part_ml[prow * 2 + 1];
    const float4* src = reinterpret_cast<const float4*>(part_o + prow * D + lane * 8);
    const float4 a = src[0], c = src[1];
    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
    acc[4] += w * c.x; acc[5] += w * c.y; acc[6] += w * c.z; acc[7] += w * c.w;
  }
  const float inv = l > 0.0f ? 1.0f / l : 0.0f;
  uint4 o = make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                       pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
  *reinterpret_cast<uint4*>(O + (long long)b * o_bstride + (long long)row * ldo + lane * 8) = o;
}

}  // namespace

size_t attn_workspace_bytes(int B, int Nq, int splits) {
  if (splits <= 1) return 0;
  return align256((size_t)B * splits * Nq * D * 4) + align256((size_t)B * splits * Nq * 2 * 4);
}

int attn_pick_splits(int B, int Nq, int Nk) {
  const int qtiles = (Nq + BM - 1) / BM;
  const int ntiles = (Nk + BN - 1) / BN;
  int s = 148 / (qtiles * B > 0 ? qtiles * B : 1);
  if (s < 1) s = 1;
  if (s > 8) s = 8;
  while (s > 1 && ntiles / s < 4) --s;  // keep at least 4 KV tiles per split
  return s;
}

int launch_attention(const AttnArgs& a, cudaStream_t stream) {
  VLS_REQUIRE(a.Q && a.K && a.Vt && a.O, "attention: null operand");
  VLS_REQUIRE(a.Nq > 0 && a.Nk > 0 && a.B > 0 && a.splits >= 1, "attention: bad shape");
  VLS_REQUIRE(a.ldo % 8 == 0, "attention: ldo must be a multiple of 8");
  const int nt = (a.Nk + BN - 1) / BN;
  VLS_REQUIRE(a.splits <= nt, "attention: more KV splits (%d) than KV tiles (%d)", a.splits, nt);
  VLS_REQUIRE(a.splits == 1 || (a.part_o && a.part_ml), "attention: split workspace missing");
  CUtensorMap tmQ, tmK, tmV;
  VLS_TRY(make_tmap_bf16(&tmQ, a.Q, D, a.Nq, a.B, a.ldq, a.q_bstride, BM));
  VLS_TRY(make_tmap_bf16(&tmK, a.K, D, a.Nk, a.B, a.ldk, a.k_bstride, BN));
  VLS_TRY(make_tmap_bf16(&tmV, a.Vt, a.Nk, D, a.B, a.ldvt, a.vt_bstride, D));
  AttnParams p;
  p.Nq = a.Nq; p.Nk = a.Nk; p.splits = a.splits;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.O = reinterpret_cast<bf16*>(a.O); p.ldo = a.ldo; p.o_bstride = a.o_bstride;
  p.part_o = a.part_o; p.part_ml = a.part_ml;
  static bool attr_set = false;
  if (!attr_set) {
    VLS_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((a.Nq + BM - 1) / BM, a.splits, a.B);
  const int slot = a.Nk > a.Nq ? PROF_ATTN_CROSS : PROF_ATTN_SELF;
  prof_begin(slot, stream);
  attn_fwd_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  prof_end(slot, stream);
  VLS_POST_LAUNCH(1);
  if (a.splits > 1) {
    const long long rows = (long long)a.B * a.Nq;
    const int wpb = 8;
    attn_combine_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
        a.part_o, a.part_ml, a.B, a.Nq, a.splits, reinterpret_cast<bf16*>(a.O), a.ldo, a.o_bstride);
    VLS_POST_LAUNCH(1);
  }
  return 0;
}

}  // namespace vls
