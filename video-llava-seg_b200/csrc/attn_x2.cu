// Memory cross-attention, second generation (r2): TWO 128-query tiles per CTA against every 128-key tile.
// softmax(Q K^T / 16) mem with q/k dim 256 and the raw 64-d memory rows as values (sam/transformer.py:311-360,
// memory_attention.py:66-81; the value / output projections are folded into the layer tail).
//
// Why: the one-query-tile kernel (attn_tc.cu) streams 80 KB of K + V from L2 per (128 x 128) unit -- 588 MB per
// launch at B=1, 6.6 TB/s, at the limit of what the L2 slices deliver to 148 SMs (K-tile latency 3900 cycles with two
// tiles in flight in the clock64 trace) -- and its single softmax group serialises with the MMA issue: the tensor
// pipe was busy 1690 of 3600 cycles per unit (ncu: 44 %).  Here a CTA owns 256 queries, so every K / V byte is used
// twice (40 KB per unit), and the two query tiles A and B run as a ping-pong: while the softmax warps of A work on
// S_A(j), the tensor pipe computes S_B(j), then PV_A(j) and S_A(j+1) while the softmax warps of B work, and so on --
// the pipe never waits for a softmax in steady state (softmax of one tile ~1500 cycles < PV + S of the other ~1690).
//
//   warp 0      : TMA producer of Q_A, Q_B (once) and of the K tiles, as 16 KB channel panels [128 keys x 64 ch] through
//                 a 5-slot ring (a panel of tile j+1 is requested as soon as S_B(j) has consumed the slot)
//   warp 1      : single-thread tcgen05.mma issuer.  Per key tile, in this order (the pipe executes in order):
//                   PV_A(j) , S_A(j+1) , PV_B(j) , S_B(j+1)
//                 S_g = Q_g K^T: SS, M128 N128 K16 x 16 -> TMEM S_g (128 f32 columns); P_g (bf16) overwrites the first 64
//                 columns of S_g, which is safe because PV_g(j) is issued before S_g(j+1);
//                 O_g += P_g V: TS, A = P_g in TMEM, B = the V tile as the bank stores it ([key][64 ch] rows, MN-major),
//                 M128 N64 K16 x 8
//   warp 2      : TMA producer of the V tiles (one 16 KB stage)
//   warps 4-7   : softmax of query tile A, thread = query row (no cross-thread exchange): 128 scores from TMEM, running
//                 max in the log2 domain with lazy rescale of O (threshold 8), ex2.approx, P back to TMEM as bf16
//   warps 8-11  : the same for query tile B
// s_full[g] is committed after S_g(j), i.e. after PV_g(j-1) has completed as well, so the softmax threads may touch
// O_g (lazy rescale) without a further barrier.
// KV splits: grid = (query pairs, splits, B); partial (O, m, l) go to the workspace and attn_x2_combine_kernel merges
// them (deterministic: fixed split order).
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int BM = 128;                 // rows of one query tile; a CTA owns two
constexpr int BN = 128;                 // keys per tile
constexpr int D = 256;
constexpr int DV = 64;
constexpr int Q_BYTES = BM * D * 2;     // 64 KB per query tile
constexpr int PANEL_BYTES = BN * 128;   // 16 KB: 128 keys x 64 channels
constexpr int KSLOTS = 5;
constexpr int V_BYTES = BN * DV * 2;    // 16 KB
constexpr int SMEM = 2 * Q_BYTES + KSLOTS * PANEL_BYTES + V_BYTES + 256 + 1024;
static_assert(SMEM <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
constexpr int THREADS = 384;            // warpgroup 0: producers + MMA, warpgroups 1 / 2: softmax of tile A / B
constexpr uint32_t TMEM_COLS = 512;     // 384 used
constexpr uint32_t TM_O = 0;            // O_A: 0..63, O_B: 64..127
constexpr uint32_t TM_S = 128;          // S_A: 128..255, S_B: 256..383
constexpr float RESCALE_THRESHOLD = 8.0f;
constexpr int MAX_SPLITS = 16;

struct X2Params {
  int Nq, Nk, splits, ntiles;
  float scale_log2;
  bf16* O;
  long long ldo, o_bstride;
  float* part_o;     // [B][splits][Nq][64]
  float* part_ml;    // [B][splits][Nq][2]
  long long* trace;  // optional: [role 0..3][tile][8] clock64 stamps of CTA (0,0,0)
};

#define X2_TRACE(role, tile, slot)                                                                          \
  do {                                                                                                      \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 48)                    \
      p.trace[((role) * 48 + (tile)) * 8 + (slot)] = clock64();                                             \
  } while (0)

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) { tmem_ld32(taddr, r); }

__global__ void __launch_bounds__(THREADS, 1)
attn_x2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const X2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // [tile g][4 panels][128 rows][128 B]
  uint8_t* sK = smem + 2 * Q_BYTES;                 // [slot][128 keys][128 B]
  uint8_t* sV = sK + KSLOTS * PANEL_BYTES;          // [128 keys][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + V_BYTES);
  uint64_t* q_full = bars;                // [2]
  uint64_t* k_full = bars + 2;            // [KSLOTS]
  uint64_t* k_empty = bars + 2 + KSLOTS;  // [KSLOTS]
  uint64_t* v_full = bars + 2 + 2 * KSLOTS;
  uint64_t* v_empty = v_full + 1;
  uint64_t* s_full = v_empty + 1;         // [2]
  uint64_t* p_ready = s_full + 2;         // [2]
  uint64_t* o_done = p_ready + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BM);
  const int split = blockIdx.y;
  const int bz = blockIdx.z;
  const int t0 = (int)((long long)p.ntiles * split / p.splits);
  const int n = (int)((long long)p.ntiles * (split + 1) / p.splits) - t0;

  if (threadIdx.x == 0) {
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    for (int s = 0; s < KSLOTS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_ready[g], BM);
    }
    mbar_init(o_done, 1);
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
  pdl_enter();

  if (warp == 0) {
    if (lane == 0 && n > 0) {
      mbar_expect_tx(&q_full[0], Q_BYTES);
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) tma_load_3d(sQ + kp * PANEL_BYTES, &tmQ, &q_full[0], kp * 64, q0, bz);
      int i = 0;   // panels issued so far
      for (int j = 0; j < n; ++j) {
        const int kv0 = (t0 + j) * BN;
#pragma unroll 1
        for (int kp = 0; kp < 4; ++kp, ++i) {
          const int slot = i % KSLOTS;
          mbar_wait(&k_empty[slot], ((i / KSLOTS) & 1) ^ 1);
          if (kp == 0) X2_TRACE(0, j, 0);
          mbar_expect_tx(&k_full[slot], PANEL_BYTES);
          tma_load_3d(sK + slot * PANEL_BYTES, &tmK, &k_full[slot], kp * 64, kv0, bz);
        }
        if (j == 0) {   // the second query tile is needed only after S_A(0)
          mbar_expect_tx(&q_full[1], Q_BYTES);
#pragma unroll
          for (int kp = 0; kp < 4; ++kp) tma_load_3d(sQ + Q_BYTES + kp * PANEL_BYTES, &tmQ, &q_full[1], kp * 64, q0 + BM, bz);
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      for (int j = 0; j < n; ++j) {
        mbar_wait(v_empty, (j & 1) ^ 1);
        X2_TRACE(0, j, 1);
        mbar_expect_tx(v_full, V_BYTES);
        tma_load_3d(sV, &tmV, v_full, 0, (t0 + j) * BN, bz);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BM, DV) | (1u << 16);   // bit 16: B is MN-major
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      auto issue_s = [&](int g, int j) {
        const uint32_t d_s = tmem + TM_S + uint32_t(g) * BN;
#pragma unroll 1
        for (int kp = 0; kp < 4; ++kp) {
          const int i = 4 * j + kp;
          const int slot = i % KSLOTS;
          if (g == 0) {
            mbar_wait(&k_full[slot], (i / KSLOTS) & 1);
            tc_fence_after();
          }
          const uint64_t qd = make_desc_sw128(q_addr + g * Q_BYTES + kp * PANEL_BYTES);
          const uint64_t kd = make_desc_sw128(k_addr + slot * PANEL_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(d_s, qd + 2 * kk, kd + 2 * kk, idesc_s, (kp | kk) != 0 ? 1u : 0u);
          if (g == 1) umma_commit(&k_empty[slot]);   // both query tiles have consumed the panel
        }
        umma_commit(&s_full[g]);
      };
      mbar_wait(&q_full[0], 0);
      X2_TRACE(1, 0, 0);
      issue_s(0, 0);
      mbar_wait(&q_full[1], 0);
      issue_s(1, 0);
      X2_TRACE(1, 0, 1);
      for (int j = 0; j < n; ++j) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_ready[g], j & 1);
          X2_TRACE(1, j, 2 + 3 * g);
          if (g == 0) mbar_wait(v_full, j & 1);
          tc_fence_after();
          const uint32_t a_p = tmem + TM_S + uint32_t(g) * BN;
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks)   // 16 keys = two 8-row groups of 1024 B; a key's 64 channels are one swizzle atom
            umma_ts(tmem + TM_O + uint32_t(g) * DV, a_p + ks * 8, make_desc_sw128(v_addr + ks * 2048), idesc_pv,
                    (j | ks) != 0 ? 1u : 0u);
          if (g == 1) umma_commit(v_empty);
          X2_TRACE(1, j, 3 + 3 * g);
          if (j + 1 < n) issue_s(g, j + 1);
          X2_TRACE(1, j, 4 + 3 * g);
        }
      }
      umma_commit(o_done);
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2;     // query tile of this warpgroup
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int rl = q * 32 + lane;
    const int row = q0 + g * BM + rl;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tS = tmem + lane_off + TM_S + uint32_t(g) * BN;
    const uint32_t tO = tmem + lane_off + TM_O + uint32_t(g) * DV;
    const bool tr = p.trace && threadIdx.x == 128 + g * 128;
    float m_used = -INFINITY;
    float l = 0.0f;
    for (int j = 0; j < n; ++j) {
      mbar_wait(&s_full[g], j & 1);
      if (tr) X2_TRACE(2 + g, j, 0);
      tc_fence_after();
      uint32_t r[128];
      tmem_ld32(tS, r);
      tmem_ld32(tS + 32, r + 32);
      tmem_ld32(tS + 64, r + 64);
      tmem_ld32(tS + 96, r + 96);
      tc_wait_ld();
      const int valid = p.Nk - (t0 + j) * BN;
      if (valid < BN) {   // only the last key tile of the sequence is ragged
#pragma unroll
        for (int i = 0; i < BN; ++i)
          if (i >= valid) r[i] = 0xff800000u;   // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < BN; i += 8) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(r[i + 4]), __uint_as_float(r[i + 5])));
        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(r[i + 6]), __uint_as_float(r[i + 7])));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if (tr) X2_TRACE(2 + g, j, 1);
      const float m_new = fmaxf(m_used, mx * p.scale_log2);
      const bool need = m_new > m_used + RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        float alpha = 1.0f;
        if (need) {
          alpha = ex2_approx(m_used - m_new);
          m_used = m_new;
        }
        l *= alpha;
        if (j > 0) {   // s_full(j) implies PV(j-1) has completed: O may be rescaled in place
#pragma unroll 1
          for (int c = 0; c < DV / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO + c * 32, o);
          }
        }
      }
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = __uint_as_float(r[c * 32 + 2 * i]) * p.scale_log2 - m_used;
          const float x1 = __uint_as_float(r[c * 32 + 2 * i + 1]) * p.scale_log2 - m_used;
          const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          l0 += p0;
          l1 += p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st16(tS + c * 16, pk);
      }
      l += l0 + l1;
      if (tr) X2_TRACE(2 + g, j, 2);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_ready[g]);
      if (tr) X2_TRACE(2 + g, j, 3);
    }
    if (n > 0) {
      mbar_wait(o_done, 0);
      tc_fence_after();
      uint32_t o[64];
      tmem_ld32(tO, o);
      tmem_ld32(tO + 32, o + 32);
      tc_wait_ld();
      if (row < p.Nq) {
        if (p.splits == 1) {
          const float inv = l > 0.0f ? 1.0f / l : 0.0f;
          uint4* o4 = reinterpret_cast<uint4*>(p.O + (long long)bz * p.o_bstride + (long long)row * p.ldo);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o4[i] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                               pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                               pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                               pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
        } else {
          const long long prow = ((long long)bz * p.splits + split) * p.Nq + row;
          float4* o4 = reinterpret_cast<float4*>(p.part_o + prow * DV);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            o4[i] = make_float4(__uint_as_float(o[4 * i]), __uint_as_float(o[4 * i + 1]), __uint_as_float(o[4 * i + 2]),
                                __uint_as_float(o[4 * i + 3]));
          *reinterpret_cast<float2*>(p.part_ml + prow * 2) = make_float2(m_used, l);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// Merge the KV-split partials: one warp per query row, a lane owns 2 of the 64 channels.  All (m, l) pairs and all
// partial vectors of the row are loaded before any arithmetic (two L2 round trips per row, not two per split).
__global__ void attn_x2_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, int B, int Nq,
                                       int splits, bf16* __restrict__ O, long long ldo, long long o_bstride) {
  pdl_enter();
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (long long)B * Nq) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(gw / Nq);
  const int row = (int)(gw % Nq);
  float2 ml[MAX_SPLITS], pv[MAX_SPLITS];
#pragma unroll
  for (int k = 0; k < MAX_SPLITS; ++k) {
    const long long prow = ((long long)b * splits + (k < splits ? k : 0)) * Nq + row;
    ml[k] = k < splits ? *reinterpret_cast<const float2*>(part_ml + prow * 2) : make_float2(-INFINITY, 0.f);
    pv[k] = k < splits ? *reinterpret_cast<const float2*>(part_o + prow * DV + lane * 2) : make_float2(0.f, 0.f);
  }
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < MAX_SPLITS; ++k) m = fmaxf(m, ml[k].x);
  float l = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int k = 0; k < MAX_SPLITS; ++k) {
    if (k < splits) {
      const float w = exp2f(ml[k].x - m);
      l += w * ml[k].y;
      a0 += w * pv[k].x;
      a1 += w * pv[k].y;
    }
  }
  const float inv = l > 0.0f ? 1.0f / l : 0.0f;
  *reinterpret_cast<uint32_t*>(O + (long long)b * o_bstride + (long long)row * ldo + lane * 2) = pack_bf16x2(a0 * inv, a1 * inv);
}

}  // namespace

int g_attn_x2 = 1;   // memory cross-attention (dv = 64, V as bank rows): 1 = two query tiles per CTA (this file), 0 = attn_tc.cu

// KV splits of the two-query-tile kernel: minimise waves x (key tiles per CTA + fixed per-CTA cost in tile times);
// a split is only taken when it is worth > 5 % (partials + a combine launch come with it)
int attn_x2_pick_splits(int B, int Nq, int Nk) {
  const long long pairs = (long long)B * ((Nq + 2 * BM - 1) / (2 * BM));
  const int nt = (Nk + BN - 1) / BN;
  const int overhead = 3;
  long long best_cost = 0;
  int best = 1;
  for (int s = 1; s <= MAX_SPLITS; ++s) {
    if (s > 1 && nt / s < 4) break;
    const long long waves = (pairs * s + 147) / 148;
    const long long cost = waves * ((nt + s - 1) / s + overhead);
    if (s == 1 || cost * 100 < best_cost * 95) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

int launch_attention_x2(const AttnArgs& a, cudaStream_t stream) {
  VLS_REQUIRE(a.dv == DV && a.v_rows, "attention x2: needs value dim 64 given as rows");
  VLS_REQUIRE(a.splits >= 1 && a.splits <= MAX_SPLITS, "attention x2: 1..%d KV splits (got %d)", MAX_SPLITS, a.splits);
  const int nt = (a.Nk + BN - 1) / BN;
  VLS_REQUIRE(a.splits <= nt, "attention: more KV splits (%d) than KV tiles (%d)", a.splits, nt);
  VLS_REQUIRE(a.splits == 1 || (a.part_o && a.part_ml), "attention: split workspace missing");
  VLS_REQUIRE(a.ldo % 8 == 0, "attention: ldo must be a multiple of 8");
  CUtensorMap tmQ, tmK, tmV;
  VLS_TRY(make_tmap_bf16(&tmQ, a.Q, D, a.Nq, a.B, a.ldq, a.q_bstride, BM));
  VLS_TRY(make_tmap_bf16(&tmK, a.K, D, a.Nk, a.B, a.ldk, a.k_bstride, BN));
  VLS_TRY(make_tmap_bf16(&tmV, a.Vt, DV, a.Nk, a.B, a.ldvt, a.vt_bstride, BN));
  static unsigned long long attr_set = 0;
  if (first_use_on_device(&attr_set))
    VLS_CUDA(cudaFuncSetAttribute(attn_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  X2Params p;
  p.Nq = a.Nq; p.Nk = a.Nk; p.splits = a.splits; p.ntiles = nt;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.O = reinterpret_cast<bf16*>(a.O); p.ldo = a.ldo; p.o_bstride = a.o_bstride;
  p.part_o = a.part_o; p.part_ml = a.part_ml;
  p.trace = g_attn_trace;
  const int pairs = (a.Nq + 2 * BM - 1) / (2 * BM);
  VLS_CUDA(launch_k(attn_x2_kernel, dim3(pairs, a.splits, a.B), dim3(THREADS), SMEM, stream, tmQ, tmK, tmV, p));
  VLS_POST_LAUNCH(1);
  if (a.splits > 1) {
    const long long rows = (long long)a.B * a.Nq;
    VLS_REQUIRE(rows < (1ll << 31), "attention: too many query rows");
    const int wpb = 8;
    VLS_CUDA(launch_k(attn_x2_combine_kernel, dim3((unsigned)((rows + wpb - 1) / wpb)), dim3(wpb * 32), 0, stream, a.part_o,
                      a.part_ml, a.B, a.Nq, a.splits, reinterpret_cast<bf16*>(a.O), a.ldo, a.o_bstride));
    VLS_POST_LAUNCH(1);
  }
  return 0;
}

}  // namespace vls
