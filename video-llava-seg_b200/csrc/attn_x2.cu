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
// Shared memory: with both query tiles in shared memory (128 KB) only 5 K panels of 16 KB fitted, a panel of tile j+1
// could be requested only when S_B(j) had consumed its slot, and the first version ran at 5200 cycles per key tile
// because S_A(j+1) waited for its K panels (TMA latency under load ~1750 cycles for 16 KB; in-kernel clock64 trace).
// So Q_A lives in TENSOR MEMORY (128 columns of packed bf16, written once by the softmax threads of tile A) and
// S_A = Q_A K^T is a TS MMA; shared memory then holds Q_B (64 KB), TWO whole K tiles (128 KB) and two V tiles (32 KB):
// tile j+1 is in flight during the whole of tile j.
//
//   warp 0      : TMA producer of Q_B (once) and of the K tiles (64 KB = 4 channel panels [128 keys x 64 ch]; 2 stages)
//   warp 1      : single-thread tcgen05.mma issuer.  Per key tile, in this order (the pipe executes in order):
//                   PV_A(j) , S_A(j+1) , PV_B(j) , S_B(j+1)
//                 S_A = Q_A K^T: TS (A = Q_A in TMEM), S_B = Q_B K^T: SS; M128 N128 K16 x 16 -> TMEM S_g (128 f32
//                 columns); P_g (bf16) overwrites the first 64 columns of S_g, which is safe because PV_g(j) is issued
//                 before S_g(j+1);
//                 O_g += P_g V: TS, A = P_g in TMEM, B = the V tile as the bank stores it ([key][64 ch] rows, MN-major),
//                 M128 N64 K16 x 8
//   warp 2      : TMA producer of the V tiles (two 16 KB stages)
//   warps 4-11  : softmax, 256 threads = 2 per query row (64 scores each; the row max is exchanged through shared memory),
//                 ALL of them on tile A, then on tile B, alternating: running max in the log2 domain with lazy rescale of
//                 O (threshold 8), half of the exponentials on MUFU.EX2 and half as a degree-3 polynomial in packed
//                 FFMA2 arithmetic (exp2_poly2), P back to TMEM as bf16.  (First version: one warpgroup per query tile,
//                 thread = row: 1500-1600 cycles per softmax whatever the MUFU / polynomial mix -- a single warp per
//                 scheduler is latency bound -- and the chain softmax_g(j) -> PV_g(j) -> S_g(j+1) -> softmax_g(j+1) set
//                 the period: 4300 cycles per key tile against 3400 of tensor work.)
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
constexpr int K_BYTES = 4 * PANEL_BYTES; // 64 KB per key tile
constexpr int KST = 2;                  // K stages (whole tiles)
constexpr int VST = 2;                  // V stages
constexpr int V_BYTES = BN * DV * 2;    // 16 KB
constexpr int XCHG_BYTES = 2 * 2 * BM * 4;   // row-max exchange [tile][half][row] f32 (reused for the row sums at the end)
constexpr int TILE_BYTES = Q_BYTES + KST * K_BYTES + VST * V_BYTES;   // 224 KB of 1024-byte aligned operand tiles
// the exchange area and the barriers sit in FRONT of the aligned tiles, inside the alignment slack
constexpr int SMEM = 232448;
static_assert(XCHG_BYTES + 256 + 768 + TILE_BYTES <= SMEM, "shared memory layout does not fit");
constexpr int THREADS = 384;            // warpgroup 0: producers + MMA, warpgroups 1 / 2: softmax of tile A / B
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_O = 0;            // O_A: 0..63, O_B: 64..127
constexpr uint32_t TM_S = 128;          // S_A: 128..255, S_B: 256..383
constexpr uint32_t TM_QA = 384;         // Q_A: 384..511 (column c of lane r = channels 2c, 2c+1 of query row r, bf16)
constexpr float RESCALE_THRESHOLD = 8.0f;
constexpr int MAX_SPLITS = 16;

struct X2Params {
  int Nq, Nk, splits, ntiles;
  float scale_log2;
  const bf16* Q;     // query rows (tile A goes to tensor memory through registers)
  long long ldq, q_bstride;
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

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {   // two FMAs per issue slot (FFMA2)
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
// 2^x for a pair on the FMA pipe instead of the SFU (the softmax is MUFU-bound: 16 384 exponentials per 128 x 128 tile at
// 16 per clock per SM = 1024 cycles, as long as the tile's QK^T MMAs).  x = n + f with n = round(x) through the 1.5 * 2^23
// trick, 2^f on [-0.5, 0.5] by a degree-3 minimax polynomial (relative error 7.5e-5, P is rounded to bf16 = 3.9e-3
// afterwards), 2^n added into the exponent field.  x is clamped at -126 (result 1e-38 instead of 0: harmless).
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 t = fadd2(x, make_float2(12582912.0f, 12582912.0f));
  const float2 nf = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = ffma2(nf, make_float2(-1.0f, -1.0f), x);
  float2 pl = ffma2(make_float2(0.055170830339193344f, 0.055170830339193344f), f, make_float2(0.24260906875133514f, 0.24260906875133514f));
  pl = ffma2(pl, f, make_float2(0.693260908126831f, 0.693260908126831f));
  pl = ffma2(pl, f, make_float2(0.9999281764030457f, 0.9999281764030457f));
  return make_float2(__int_as_float(__float_as_int(pl.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(pl.y) + (__float_as_int(t.y) << 23)));
}

// POLY: of every 4 pairs of scores, POLY go through exp2_poly2 and 4 - POLY through MUFU.EX2
template <int POLY>
__global__ void __launch_bounds__(THREADS, 1)
attn_x2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP, const X2Params p) {
  extern __shared__ uint8_t smem_raw[];
  float* xchg = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + XCHG_BYTES);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + XCHG_BYTES + 256 + 1023) & ~uintptr_t(1023));
  if (smem + TILE_BYTES > smem_raw + SMEM) __trap();   // dynamic shared memory base less aligned than assumed
  uint8_t* sQ = smem;                               // Q_B: [4 panels][128 rows][128 B]
  uint8_t* sK = smem + Q_BYTES;                     // [stage][4 panels][128 keys][128 B]
  uint8_t* sV = sK + KST * K_BYTES;                 // [stage][128 keys][128 B]
  uint64_t* q_full = bars;                // [0]: Q_A is in TMEM (128 arrivals), [1]: Q_B has landed in shared memory
  uint64_t* k_full = bars + 2;            // [KST]
  uint64_t* k_empty = k_full + KST;       // [KST]
  uint64_t* v_full = k_empty + KST;       // [VST]
  uint64_t* v_empty = v_full + VST;       // [VST]
  uint64_t* s_full = v_empty + VST;       // [2]
  uint64_t* p_ready = s_full + 2;         // [2]
  uint64_t* o_done = p_ready + 2;
  uint64_t* spy = o_done + 1;             // [2] trace only: PV_g(j) has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(spy + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BM);
  const int split = blockIdx.y;
  const int bz = blockIdx.z;
  const int t0 = (int)((long long)p.ntiles * split / p.splits);
  const int n = (int)((long long)p.ntiles * (split + 1) / p.splits) - t0;

  if (threadIdx.x == 0) {
    mbar_init(&q_full[0], 2 * BM);
    mbar_init(&q_full[1], 1);
    for (int s = 0; s < KST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < VST; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_ready[g], 2 * BM);
    }
    mbar_init(o_done, 1);
    mbar_init(&spy[0], 1);
    mbar_init(&spy[1], 1);
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
  if (p.trace && threadIdx.x == 0 && (blockIdx.x | blockIdx.z) == 0) {   // SM clock against wall time, per KV split
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.trace[(3 * 48 + 32 + blockIdx.y) * 8 + 4] = clock64();
    p.trace[(3 * 48 + 32 + blockIdx.y) * 8 + 5] = (long long)ns;
  }

  if (warp == 0) {
    if (n > 0 && elect_one()) {
      for (int j = 0; j < n; ++j) {
        const int st = j % KST;
        mbar_wait(&k_empty[st], ((j / KST) & 1) ^ 1);
        X2_TRACE(0, j, 0);
        mbar_expect_tx(&k_full[st], K_BYTES);
#pragma unroll
        for (int kp = 0; kp < 4; ++kp)
          tma_load_3d(sK + st * K_BYTES + kp * PANEL_BYTES, &tmK, &k_full[st], kp * 64, (t0 + j) * BN, bz);
        if (j == 0) {   // the second query tile is needed only after S_A(0)
          mbar_expect_tx(&q_full[1], Q_BYTES);
#pragma unroll
          for (int kp = 0; kp < 4; ++kp) tma_load_3d(sQ + kp * PANEL_BYTES, &tmQ, &q_full[1], kp * 64, q0 + BM, bz);
        }
      }
    }
  } else if (warp == 2) {
    if (elect_one()) {
      for (int j = 0; j < n; ++j) {
        const int st = j % VST;
        mbar_wait(&v_empty[st], ((j / VST) & 1) ^ 1);
        X2_TRACE(0, j, 1);
        mbar_expect_tx(&v_full[st], V_BYTES);
        tma_load_3d(sV + st * V_BYTES, &tmV, &v_full[st], 0, (t0 + j) * BN, bz);
      }
    }
  } else if (warp == 1) {
    // elect.sync, not `lane == 0`: behind a lane test the compiler treats the tcgen05 instructions (uniform datapath) as
    // divergent code and wraps EVERY MMA in an ELECT / BRA.U.ANY loop with three R2UR moves -- 80-100 cycles of issue per
    // MMA against 64-71 of execution, i.e. the issuing thread, not the tensor pipe, set the pace (SASS + clock64 trace)
    if (n > 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BM, DV) | (1u << 16);   // bit 16: B is MN-major
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      auto issue_s = [&](int g, int j) {
        const uint32_t d_s = tmem + TM_S + uint32_t(g) * BN;
        const int st = j % KST;
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) {
          const uint64_t kd = make_desc_sw128(k_addr + st * K_BYTES + kp * PANEL_BYTES);
          if (g == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ts(d_s, tmem + TM_QA + uint32_t(kp * 4 + kk) * 8, kd + 2 * kk, idesc_s, (kp | kk) != 0 ? 1u : 0u);
          } else {
            const uint64_t qd = make_desc_sw128(q_addr + kp * PANEL_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_ss(d_s, qd + 2 * kk, kd + 2 * kk, idesc_s, (kp | kk) != 0 ? 1u : 0u);
          }
        }
        if (g == 1) umma_commit(&k_empty[st]);   // both query tiles have consumed the K tile
        umma_commit(&s_full[g]);
      };
      mbar_wait(&k_full[0], 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      X2_TRACE(1, 0, 0);
      issue_s(0, 0);
      mbar_wait(&q_full[1], 0);
      issue_s(1, 0);
      X2_TRACE(1, 0, 1);
      for (int j = 0; j < n; ++j) {
        const int vs = j % VST;
        // operands that were requested a whole tile ago first: these waits return at once and must not sit between a
        // softmax's arrival and the MMAs that depend on it
        mbar_wait(&v_full[vs], (j / VST) & 1);
        if (j + 1 < n) mbar_wait(&k_full[(j + 1) % KST], ((j + 1) / KST) & 1);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_ready[g], j & 1);
          X2_TRACE(1, j, 2 + 3 * g);
          tc_fence_after();
          const uint32_t a_p = tmem + TM_S + uint32_t(g) * BN;
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks)   // 16 keys = two 8-row groups of 1024 B; a key's 64 channels are one swizzle atom
            umma_ts(tmem + TM_O + uint32_t(g) * DV, a_p + ks * 8, make_desc_sw128(v_addr + vs * V_BYTES + ks * 2048), idesc_pv,
                    (j | ks) != 0 ? 1u : 0u);
          if (g == 1) umma_commit(&v_empty[vs]);
          if (p.trace) umma_commit(&spy[g]);
          X2_TRACE(1, j, 3 + 3 * g);
          if (j + 1 < n) issue_s(g, j + 1);
          X2_TRACE(1, j, 4 + 3 * g);
        }
      }
      umma_commit(o_done);
    }
  } else if (warp == 3) {
    if (lane == 0 && p.trace) {          // trace only: completion times of the PV MMAs
      for (int j = 0; j < n; ++j) {
        mbar_wait(&spy[0], j & 1);
        X2_TRACE(0, j, 2);
        mbar_wait(&spy[1], j & 1);
        X2_TRACE(0, j, 3);
      }
    }
  } else if (warp >= 4) {
    // 256 softmax threads = 2 per query row (columns 0-63 / 64-127 of the S tile, channels 0-31 / 32-63 of O); all of
    // them work on tile A, then on tile B, then on A of the next key tile ...: the SFU is the floor of a softmax (16 384
    // exponentials at 16 per clock = 1024 cycles per tile) whichever way the threads are split, and with two warps per
    // scheduler on the same tile one warp's FMA-pipe polynomial exponentials overlap the other's MUFU ones.
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // column half of the row owned by this thread
    const int rl = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const bool tr = p.trace && threadIdx.x == 128;
    if (n > 0) {
      // Q_A: half a query row (256 B) per thread straight from global memory into tensor memory, once per CTA
      const int row = q0 + rl;
      const uint4* qrow = reinterpret_cast<const uint4*>(p.Q + (long long)bz * p.q_bstride + (long long)row * p.ldq) + half * 16;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t qv[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v = row < p.Nq ? __ldg(qrow + c * 8 + i) : make_uint4(0u, 0u, 0u, 0u);
          qv[4 * i] = v.x; qv[4 * i + 1] = v.y; qv[4 * i + 2] = v.z; qv[4 * i + 3] = v.w;
        }
        tmem_st32(tmem + lane_off + TM_QA + half * 64 + c * 32, qv);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&q_full[0]);
    }
    float m_used[2] = {-INFINITY, -INFINITY};
    float l[2] = {0.0f, 0.0f};
    for (int j = 0; j < n; ++j) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const uint32_t tS = tmem + lane_off + TM_S + uint32_t(g) * BN;
        const uint32_t tO = tmem + lane_off + TM_O + uint32_t(g) * DV + half * 32;
        mbar_wait(&s_full[g], j & 1);
        if (tr) X2_TRACE(2 + g, j, 0);
        tc_fence_after();
        uint32_t r[64];
        tmem_ld32(tS + half * 64, r);
        tmem_ld32(tS + half * 64 + 32, r + 32);
        tc_wait_ld();
        const int valid = p.Nk - (t0 + j) * BN - half * 64;   // may be <= 0 for the upper half of the last tile
        if (valid < 64) {   // only the last key tile of the sequence is ragged
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= valid) r[i] = 0xff800000u;   // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(r[i + 4]), __uint_as_float(r[i + 5])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(r[i + 6]), __uint_as_float(r[i + 7])));
        }
        float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        // both threads of a row need the max over all 128 columns: exchange through shared memory ([tile g][half][row];
        // the buffer of tile g is rewritten one whole softmax of the other tile later, after several barriers)
        float* xb = xchg + g * (2 * BM);
        xb[half * BM + rl] = mx;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mx = fmaxf(mx, xb[(half ^ 1) * BM + rl]);
        if (tr) X2_TRACE(2 + g, j, 1);
        const float m_new = fmaxf(m_used[g], mx * p.scale_log2);
        const bool need = m_new > m_used[g] + RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          float alpha = 1.0f;
          if (need) {
            alpha = ex2_approx(m_used[g] - m_new);
            m_used[g] = m_new;
          }
          l[g] *= alpha;
          if (j > 0) {   // s_full(j) implies PV(j-1) has completed: O may be rescaled in place
            uint32_t o[32];
            tmem_ld32(tO, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO, o);
          }
        }
        float2 lacc = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_used[g], -m_used[g]);
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float2 x = ffma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, nm2);
          const float2 e = (i & 3) < POLY ? exp2_poly2(x) : make_float2(ex2_approx(x.x), ex2_approx(x.y));
          lacc = fadd2(lacc, e);
          pk[i] = pack_bf16x2(e.x, e.y);
        }
        tmem_st32(tS + half * 32, pk);
        l[g] += lacc.x + lacc.y;
        if (tr) X2_TRACE(2 + g, j, 2);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&p_ready[g]);
        if (tr) X2_TRACE(2 + g, j, 3);
      }
    }
    if (n > 0) {
      mbar_wait(o_done, 0);
      tc_fence_after();
      // total row sums = sum of the two halves (same m_used on both)
      // [tile g][half][row], aliasing the max exchange: o_done implies that every thread has arrived on p_ready for the last
      // tile, i.e. has read its partner's last row max long ago
      float* lx = xchg;
      lx[(0 * 2 + half) * BM + rl] = l[0];
      lx[(1 * 2 + half) * BM + rl] = l[1];
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float lt = l[g] + lx[(g * 2 + (half ^ 1)) * BM + rl];
        const int row = q0 + g * BM + rl;
        uint32_t o[32];
        tmem_ld32(tmem + lane_off + TM_O + uint32_t(g) * DV + half * 32, o);
        tc_wait_ld();
        if (row < p.Nq) {
          if (p.splits == 1) {
            const float inv = lt > 0.0f ? 1.0f / lt : 0.0f;
            uint4* o4 = reinterpret_cast<uint4*>(p.O + (long long)bz * p.o_bstride + (long long)row * p.ldo + half * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o4[i] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                                 pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                                 pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                                 pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
          } else {
            const long long prow = ((long long)bz * p.splits + split) * p.Nq + row;
            if (half == 0) *reinterpret_cast<float2*>(p.part_ml + prow * 2) = make_float2(m_used[g], lt);
          }
        }
        if (p.splits > 1) {
          // the f32 partial tile leaves through the TMA engine: box (tile g, half) = [128 rows x 32 floats], 128B-swizzled, in
          // the dead Q_B / K buffers (thread = row stores to global memory touch 32 cache lines per warp instruction)
          uint8_t* st = smem + (g * 2 + half) * (BM * 128) + rl * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(st + ((i ^ (rl & 7)) << 4)) =
                make_float4(__uint_as_float(o[4 * i]), __uint_as_float(o[4 * i + 1]), __uint_as_float(o[4 * i + 2]), __uint_as_float(o[4 * i + 3]));
        }
      }
      if (p.splits > 1) {
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 128) {   // first softmax thread
#pragma unroll 1
          for (int bx = 0; bx < 4; ++bx)
            tma_store_3d(smem + bx * (BM * 128), &tmP, (bx & 1) * 32, q0 + (bx >> 1) * BM, bz * p.splits + split);
          tma_store_commit();
          tma_store_wait_read();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.trace && threadIdx.x == 0 && (blockIdx.x | blockIdx.z) == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.trace[(3 * 48 + 32 + blockIdx.y) * 8 + 6] = clock64();
    p.trace[(3 * 48 + 32 + blockIdx.y) * 8 + 7] = (long long)ns;
  }
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

int g_attn_x2_poly = 2;   // pairs of every 4 whose exponentials run on the FMA pipe (0..3; vls_set_tuning "attn_x2_poly")
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
  VLS_REQUIRE(a.ldo % 8 == 0 && a.ldq % 8 == 0 && a.q_bstride % 8 == 0 && (reinterpret_cast<uintptr_t>(a.Q) & 15) == 0,
              "attention: Q rows and O rows must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV;
  VLS_TRY(make_tmap_bf16(&tmQ, a.Q, D, a.Nq, a.B, a.ldq, a.q_bstride, BM));
  VLS_TRY(make_tmap_bf16(&tmK, a.K, D, a.Nk, a.B, a.ldk, a.k_bstride, BN));
  VLS_TRY(make_tmap_bf16(&tmV, a.Vt, DV, a.Nk, a.B, a.ldvt, a.vt_bstride, BN));
  CUtensorMap tmP = tmQ;   // split partials f32 [B * splits][Nq][64], stored by TMA (rows beyond Nq are clipped)
  if (a.splits > 1) VLS_TRY(make_tmap_f32(&tmP, a.part_o, DV, a.Nq, (uint64_t)a.B * a.splits, DV, (long long)a.Nq * DV, BM));
  static unsigned long long attr_set = 0;
  if (first_use_on_device(&attr_set)) {
    VLS_CUDA(cudaFuncSetAttribute(attn_x2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    VLS_CUDA(cudaFuncSetAttribute(attn_x2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    VLS_CUDA(cudaFuncSetAttribute(attn_x2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    VLS_CUDA(cudaFuncSetAttribute(attn_x2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  X2Params p;
  p.Nq = a.Nq; p.Nk = a.Nk; p.splits = a.splits; p.ntiles = nt;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.Q = reinterpret_cast<const bf16*>(a.Q); p.ldq = a.ldq; p.q_bstride = a.q_bstride;
  p.O = reinterpret_cast<bf16*>(a.O); p.ldo = a.ldo; p.o_bstride = a.o_bstride;
  p.part_o = a.part_o; p.part_ml = a.part_ml;
  p.trace = g_attn_trace;
  const int pairs = (a.Nq + 2 * BM - 1) / (2 * BM);
  const dim3 grid(pairs, a.splits, a.B);
  switch (g_attn_x2_poly) {
    case 0: VLS_CUDA(launch_k(attn_x2_kernel<0>, grid, dim3(THREADS), SMEM, stream, tmQ, tmK, tmV, tmP, p)); break;
    case 1: VLS_CUDA(launch_k(attn_x2_kernel<1>, grid, dim3(THREADS), SMEM, stream, tmQ, tmK, tmV, tmP, p)); break;
    case 3: VLS_CUDA(launch_k(attn_x2_kernel<3>, grid, dim3(THREADS), SMEM, stream, tmQ, tmK, tmV, tmP, p)); break;
    default: VLS_CUDA(launch_k(attn_x2_kernel<2>, grid, dim3(THREADS), SMEM, stream, tmQ, tmK, tmV, tmP, p)); break;
  }
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
