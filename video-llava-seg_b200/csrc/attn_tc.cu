// tcgen05 / TMEM flash attention for the memory-attention stack: one head of dim 256,
// Nq = 4096 queries, Nk up to 7*4096 + 64 keys (sam/transformer.py:311-360 after projection;
// RoPE is already applied to Q/K by the projection GEMM epilogue).
//
// CTA = one 128-query tile x one KV split; CTAs run as CLUSTERS OF 2 (adjacent query tiles, same KV range):
// each CTA TMA-loads half of every K / V^T tile and MULTICASTS it to both, halving the L2->SM operand traffic
// (the r1 ncu capture showed 941 MB / launch = 6.4 TB/s of L2 reads with private loads).
//   warp 0   : TMA producer (Q once; K / V^T tiles of 128 keys, 128B swizzle, optionally multicast)
//   warp 1   : single-thread tcgen05.mma issuer
//                S_j = Q K_j^T   (SS, M128 N128 K16 x16: an in-kernel clock64 trace showed that N=64 MMAs issue at
//                                 half rate, ~67 cycles each, so tiles are 128 keys)  -> TMEM S[j&1] (2 x 128 columns)
//                O  += P_j V_j   (TS, A = bf16 P_j in TMEM aliasing S[j&1], B = V^T tile, M128 N256 K16 x8)
//              S_{j+1} is issued before P_j is waited on, so the tensor pipe overlaps the softmax of tile j;
//              operand stages are released with a commit multicast to BOTH CTAs' empty barriers
//   warps 2-9: softmax, 256 threads = 2 threads per query row (columns 0-63 / 64-127 of the tile; the row max is
//              exchanged through shared memory), running max in the log2 domain, refreshed only when it grows by
//              more than 8 (lazy rescale of each thread's half of O in TMEM), exp2 with the 1/sqrt(d)*log2(e)
//              scale folded in, P written back to TMEM as packed bf16.
// With KV splits, partial (O, m, l) go to a workspace and attn_combine_kernel merges them.
// Balanced ("stream-K") mode, used when a fixed split would leave SMs idle (B=1: 32 query tiles x 4 splits = 128 of 148
// SMs): the (query tile, key tile) units of the launch are dealt out evenly to one persistent CTA per SM; a CTA then
// works on up to two segments (tail of one query tile's keys, head of the next one's), reloading Q in between, and
// attn_combine_bal_kernel merges the 5-6 partials of each query tile.  Cross-attention launch 112.6 -> 108.7 us.
// Value dimension DV: 256 (self-attention: V^T [256][Nk], K-major) or 64 (memory cross-attention, r2): softmax rows sum
// to one, so softmax(QK^T)(mem Wv^T + bv) = (softmax(QK^T) mem) Wv^T + bv -- the kernel attends over the raw 64-d memory
// rows and Wo.Wv is folded into the output projection (memory_attention.py:66-81, sam/transformer.py:311-360).  PV drops
// from M128 N256 to M128 N64 MMAs, the V tile from 64 KB to 16 KB (which pays for a SECOND K stage: with one stage the
// 64 KB K load of tile j+2 could only start when S(j+1) had completed and the ncu capture showed ~1100 idle tensor
// cycles per tile), O needs 64 TMEM columns, and the per-frame V^T projection GEMM (59 MB of writes) is gone.  With
// VMN the value tile is TMA-loaded as the bank stores it ([key][64 channels] rows) and consumed as an MN-major B operand.
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int BM = 128;
constexpr int BN = 128;          // keys per tile: S = Q K^T runs as M128 N128 K16 MMAs (N=64 issues at half rate)
constexpr int D = 256;
constexpr int Q_BYTES = BM * D * 2;
constexpr int K_BYTES = BN * D * 2;
constexpr int THREADS = 352;           // K-producer warp + MMA warp + 8 softmax warps + V-producer warp
constexpr int SOFTMAX_THREADS = 256;
constexpr int XCHG_BYTES = 3 * 2 * BM * 4;  // row-max exchange [tile parity][half][row] + row-sum exchange [half][row]
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_O = 0;
// per value dimension: DV=256: Q 64 KB + K 64 KB + V^T 64 KB (one stage each); DV=64: Q 64 KB + 2 x K 64 KB + V 16 KB
template <int DV>
struct ACfg {
  static constexpr int KST = DV == 64 ? 2 : 1;      // K stages
  static constexpr int V_BYTES = DV * BN * 2;       // one V stage
  static constexpr int SMEM = Q_BYTES + KST * K_BYTES + V_BYTES + XCHG_BYTES + 256 + 1024;
  static constexpr uint32_t TM_S = DV == 64 ? 128 : 256;   // 2 x 128 columns (f32 scores, overwritten in place by bf16 P)
  static constexpr int OC = DV / 2;                 // O columns owned by each of the two softmax threads of a row
};
constexpr float RESCALE_THRESHOLD = 8.0f;
constexpr uint16_t PAIR_MASK = 0x3;

struct AttnParams {
  int Nq, Nk, splits;
  int qtiles, ntiles;    // query tiles per batch element, key tiles per query tile
  long long units;       // balanced mode: B * qtiles * ntiles (query tile, key tile) work units over gridDim.x CTAs
  float scale_log2;
  bf16* O;
  long long ldo, o_bstride;
  float* part_o;
  float* part_ml;
  long long* trace;  // optional dev trace: [role 0..2][tile][8] clock64 stamps of CTA (0,0,0) (vls_attention_trace)
};

#define VLS_TRACE(role, tile, slot)                                                                         \
  do {                                                                                                      \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 64)                    \
      p.trace[((role) * 64 + (tile)) * 8 + (slot)] = clock64();                                             \
  } while (0)

constexpr int MAX_SEGS = 4;   // balanced mode: segments per CTA (needs query tiles <= (MAX_SEGS - 1) * CTAs)
// One piece of work of a CTA: `n` consecutive key tiles starting at tile t0 of query tile (bz, q0).
struct Seg {
  int q0, bz, t0, n, slot;
};

// CL : CTAs per cluster (1: private loads, 2: each CTA multicasts half of every K / V^T tile)
// BAL: balanced ("stream-K") mode.  The (query tile, key tile) units of the whole launch are dealt out evenly to
//      gridDim.x = #SMs persistent CTAs, so a CTA processes up to two SEGMENTS (the tail of one query tile's keys and the
//      head of the next one's) and every segment leaves a partial (O, m, l) in slot 2*cta + segment.  With 32 query
//      tiles x 4 fixed KV splits only 128 of the 148 SMs had work.  Barrier phases simply keep counting across segments.
// DV : value dimension (256: V^T K-major [256][Nk]; 64: see the file header)
// VMN: DV == 64 only -- V is given as rows [Nk][64] and consumed as an MN-major B operand (no transposed copy at all)
template <int CL, bool BAL, int DV, bool VMN>
__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // the dynamic smem base has the same offset in both CTAs of the pair, so this alignment is identical too
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + Q_BYTES;
  using C_ = ACfg<DV>;
  constexpr int KST = C_::KST;
  constexpr int V_BYTES = C_::V_BYTES;
  constexpr uint32_t TM_S = C_::TM_S;
  constexpr int OC = C_::OC;
  static_assert(DV == 256 || DV == 64, "value dimension must be 256 or 64");
  static_assert(!VMN || DV == 64, "row-major V needs DV == 64 (one 128-byte swizzle atom per key row)");
  uint8_t* sV = sK + KST * K_BYTES;
  float* xchg = reinterpret_cast<float*>(sV + V_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xchg) + XCHG_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 4;
  uint64_t* v_full = bars + 7;
  uint64_t* v_empty = bars + 10;
  uint64_t* s_full = bars + 13;
  uint64_t* p_ready = bars + 15;
  uint64_t* pv_done = bars + 17;
  uint64_t* o_done = bars + 18;   // one phase per segment: all PV MMAs of the segment have completed
  uint64_t* q_empty = bars + 19;  // BAL: every S MMA of the segment has read Q (the next query tile may be loaded)
  uint64_t* o_free = bars + 20;   // BAL: the epilogue has read O out of TMEM (the next segment may overwrite it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  // work of this CTA: one segment (fixed KV splits) or up to MAX_SEGS (balanced mode: its unit range may end one query
  // tile, cover whole ones and begin another)
  Seg segs[MAX_SEGS];
  int nseg = 1;
  if (BAL) {
    long long u = p.units * blockIdx.x / gridDim.x;
    const long long u1 = p.units * (blockIdx.x + 1) / gridDim.x;
    nseg = 0;
#pragma unroll
    for (int k = 0; k < MAX_SEGS; ++k) {
      segs[k] = Seg{0, 0, 0, 0, 0};
      if (u < u1) {
        const int qt = (int)(u / p.ntiles);
        Seg& S = segs[k];
        S.t0 = (int)(u - (long long)qt * p.ntiles);
        S.n = (int)min((long long)(p.ntiles - S.t0), u1 - u);
        S.bz = qt / p.qtiles;
        S.q0 = (qt - S.bz * p.qtiles) * BM;
        S.slot = MAX_SEGS * blockIdx.x + k;
        u += S.n;
        nseg = k + 1;
      }
    }
  } else {
    const int split = blockIdx.y;
    segs[0].q0 = blockIdx.x * BM;
    segs[0].bz = blockIdx.z;
    segs[0].t0 = (int)((long long)p.ntiles * split / p.splits);
    segs[0].n = (int)((long long)p.ntiles * (split + 1) / p.splits) - segs[0].t0;
    segs[0].slot = 0;
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], CL);  // released by the MMA commits of every CTA of the cluster
    }
    mbar_init(&v_full[0], 1);
    mbar_init(&v_empty[0], CL);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], SOFTMAX_THREADS);
    }
    mbar_init(pv_done, 1);
    mbar_init(o_done, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_free, SOFTMAX_THREADS);
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
  if (CL > 1) cluster_sync_all();  // the peer's barriers must be initialised before any multicast can signal them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_enter();   // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail; global memory from here on

  if (warp == 0) {
    if (elect_one()) {   // elect.sync: a lane test makes the compiler wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop
      int jg = 0;   // tiles issued so far over all segments: barrier phases keep counting
      for (int sg = 0; sg < nseg; ++sg) {
        const Seg S = segs[sg];
        if (sg > 0) mbar_wait(q_empty, (sg - 1) & 1);   // the previous segment's S MMAs are done with Q
        mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) tma_load_3d(sQ + kp * (BM * 128), &tmQ, q_full, kp * 64, S.q0, S.bz);
        for (int j = 0; j < S.n; ++j, ++jg) {
          const int st = jg % KST;
          const uint32_t ph = (jg / KST) & 1;
          const int kv0 = (S.t0 + j) * BN;
          mbar_wait(&k_empty[st], ph ^ 1);
          VLS_TRACE(0, jg, 0);
          mbar_expect_tx(&k_full[st], K_BYTES);
          if (CL > 1) {
            // K tile: this CTA fetches key rows [rank*64, rank*64+64) of each 64-channel panel for both CTAs
#pragma unroll
            for (int kp = 0; kp < 4; ++kp)
              tma_load_3d_mc(sK + st * K_BYTES + kp * (BN * 128) + rank * (64 * 128), &tmK, &k_full[st], kp * 64,
                             kv0 + (int)rank * 64, S.bz, PAIR_MASK);
          } else {
#pragma unroll
            for (int kp = 0; kp < 4; ++kp)
              tma_load_3d(sK + st * K_BYTES + kp * (BN * 128), &tmK, &k_full[st], kp * 64, kv0, S.bz);
          }
        }
      }
    }
  } else if (warp == 10) {
    // V^T producer on its own warp: with one stage per operand an in-order K,V,K,V producer would hold K(j+1)
    // back until PV(j-1) has released the V buffer
    if (elect_one()) {   // elect.sync: a lane test makes the compiler wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop
      int jg = 0;
      for (int sg = 0; sg < nseg; ++sg) {
        const Seg S = segs[sg];
        for (int j = 0; j < S.n; ++j, ++jg) {
          const uint32_t ph = jg & 1;   // one V stage
          const int kv0 = (S.t0 + j) * BN;
          mbar_wait(&v_empty[0], ph ^ 1);
          VLS_TRACE(0, jg, 1);
          mbar_expect_tx(&v_full[0], V_BYTES);
          if (VMN) {
            // V tile as stored: 128 key rows of 64 channels (128 B each), one box
            tma_load_3d(sV, &tmV, &v_full[0], 0, kv0, S.bz);
          } else {
            // V^T tile = two 64-key panels of [DV channels x 128 B]
#pragma unroll
            for (int vp = 0; vp < 2; ++vp) {
              if (CL > 1)  // this CTA fetches channel rows [rank*DV/2, (rank+1)*DV/2) of each panel for both CTAs
                tma_load_3d_mc(sV + vp * (DV * 128) + rank * (DV / 2 * 128), &tmV, &v_full[0], kv0 + vp * 64,
                               (int)rank * (DV / 2), S.bz, PAIR_MASK);
              else
                tma_load_3d(sV + vp * (DV * 128), &tmV, &v_full[0], kv0 + vp * 64, 0, S.bz);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // elect.sync: a lane test makes the compiler wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop
      constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BM, DV) | (VMN ? (1u << 16) : 0u);   // bit 16: B is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      int jg0 = 0;   // global index of the segment's first tile
      for (int sg = 0; sg < nseg; ++sg) {
        const int n = segs[sg].n;
        if (n <= 0) continue;
        mbar_wait(q_full, sg & 1);
        auto issue_s = [&](int j) {   // j: tile inside the segment; jg: global tile count (stage / phase bookkeeping)
          const int jg = jg0 + j;
          const int st = jg % KST;
          mbar_wait(&k_full[st], (jg / KST) & 1);
          VLS_TRACE(1, jg, 0);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(sK + st * K_BYTES);
          const uint32_t d_s = tmem + TM_S + uint32_t(jg & 1) * BN;
#pragma unroll
          for (int kp = 0; kp < 4; ++kp) {
            const uint64_t qd = make_desc_sw128(q_addr + kp * (BM * 128));
            const uint64_t kd = make_desc_sw128(k_addr + kp * (BN * 128));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_ss(d_s, qd + 2 * kk, kd + 2 * kk, idesc_s, (kp | kk) != 0 ? 1u : 0u);
          }
          if (CL > 1) umma_commit_mc(&k_empty[st], PAIR_MASK); else umma_commit(&k_empty[st]);
          umma_commit(&s_full[jg & 1]);
          if (BAL && j == n - 1) umma_commit(q_empty);   // last S of the segment: Q may be replaced once it completes
          VLS_TRACE(1, jg, 1);
        };
        issue_s(0);
        for (int j = 0; j < n; ++j) {
          if (j + 1 < n) issue_s(j + 1);
          const int jg = jg0 + j;
          mbar_wait(&p_ready[jg & 1], (jg >> 1) & 1);
          VLS_TRACE(1, jg, 2);
          mbar_wait(&v_full[0], jg & 1);
          VLS_TRACE(1, jg, 3);
          if (BAL && sg > 0 && j == 0) mbar_wait(o_free, (sg - 1) & 1);   // the previous segment's O has been read out
          tc_fence_after();
          const uint32_t v_addr = smem_u32(sV);
          const uint32_t a_p = tmem + TM_S + uint32_t(jg & 1) * BN;
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks) {
            // K-major V^T: 16 keys = 32 B inside the 128-byte rows of a 64-key panel; MN-major V rows: 16 keys = two
            // 8-row groups of 1024 B (SBO), the 64 channels of a key are one 128-byte swizzle atom
            const uint64_t vd = VMN ? make_desc_sw128(v_addr + ks * 2048)
                                    : make_desc_sw128(v_addr + (ks >> 2) * (DV * 128)) + 2 * (ks & 3);
            umma_ts(tmem + TM_O, a_p + ks * 8, vd, idesc_pv, (j | ks) != 0 ? 1u : 0u);
          }
          if (CL > 1) umma_commit_mc(&v_empty[0], PAIR_MASK); else umma_commit(&v_empty[0]);
          umma_commit(pv_done);
          VLS_TRACE(1, jg, 4);
        }
        umma_commit(o_done);
        jg0 += n;
      }
    }
  } else if (warp < 10) {
    const int q = warp & 3;            // TMEM lane quarter (two warps share each quarter)
    const int half = (warp - 2) >> 2;  // 0: columns 0-31 of the S tile / 0-127 of O, 1: the other halves
    const int rl = q * 32 + lane;      // row inside the tile
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    int jg0 = 0;
    for (int sg = 0; sg < nseg; ++sg) {
    const Seg S = segs[sg];
    const int n = S.n, t0 = S.t0, bz = S.bz;
    const int row = S.q0 + rl;
    float m_used = -INFINITY;
    float l = 0.0f;
    for (int jl = 0; jl < n; ++jl) {
      const int j = jg0 + jl;          // global tile index: TMEM buffer / barrier phase bookkeeping
      const int b = j & 1;
      mbar_wait(&s_full[b], (j >> 1) & 1);
      if (threadIdx.x == 64) VLS_TRACE(2, j, 0);
      tc_fence_after();
      uint32_t r[64];
      tmem_ld32(tmem + lane_off + TM_S + b * BN + half * 64, r);
      tmem_ld32(tmem + lane_off + TM_S + b * BN + half * 64 + 32, r + 32);
      tc_wait_ld();
      const int kv0 = (t0 + jl) * BN + half * 64;
      const int valid = p.Nk - kv0;  // may be <= 0 for the upper half of the last tile
      float mx = -INFINITY;
      if (valid < 64) {  // only the last key tile of the sequence is ragged
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) r[i] = 0xff800000u;  // -inf
      }
#pragma unroll
      for (int i = 0; i < 64; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
      // both threads of a row need the max over all 128 columns
      float* xb = xchg + b * (2 * BM);
      xb[half * BM + rl] = mx;
      if (threadIdx.x == 64) VLS_TRACE(2, j, 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) VLS_TRACE(2, j, 2);
      mx = fmaxf(mx, xb[(half ^ 1) * BM + rl]);
      const float m_new = fmaxf(m_used, mx * p.scale_log2);
      const bool need = m_new > m_used + RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        float alpha = 1.0f;
        if (need) {
          alpha = ex2_approx(m_used - m_new);
          m_used = m_new;
        }
        l *= alpha;
        if (jl > 0) {
          mbar_wait(pv_done, (j - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < OC / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem + lane_off + TM_O + half * OC + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem + lane_off + TM_O + half * OC + c * 32, o);
          }
          tc_wait_st();
        }
      }
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x0 = __uint_as_float(r[2 * i]) * p.scale_log2 - m_used;
        const float x1 = __uint_as_float(r[2 * i + 1]) * p.scale_log2 - m_used;
        const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
        l += p0 + p1;
        pk[i] = pack_bf16x2(p0, p1);
      }
      if (threadIdx.x == 64) VLS_TRACE(2, j, 3);
      tmem_st32(tmem + lane_off + TM_S + b * BN + half * 32, pk);
      tc_wait_st();
      if (threadIdx.x == 64) VLS_TRACE(2, j, 4);
      tc_fence_before();
      mbar_arrive(&p_ready[b]);
      if (threadIdx.x == 64) VLS_TRACE(2, j, 5);
    }
    if (n > 0) {
      // NOT pv_done: its parity can alias here.  A softmax thread only knows that S_{n-1} has completed, i.e. that
      // PV_{n-3} has, so pv_done may be one OR two phases behind and a parity wait for phase n-1 would pass in the
      // latter case (caught by the bitwise-determinism test: O read before the last two PVs had landed).  The
      // per-tile rescale wait above is safe: S_j complete => PV_{j-2} complete => at most one phase behind.
      mbar_wait(o_done, sg & 1);
      tc_fence_after();
    }
    // total row sum = sum of the two halves (same m_used on both)
    float* lx = xchg + 2 * (2 * BM);  // separate region: a partner may still be reading the last tile's max
    lx[half * BM + rl] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += lx[(half ^ 1) * BM + rl];
    const bool row_ok = row < p.Nq;
    if (!BAL && p.splits == 1) {
      const float inv = l > 0.0f ? 1.0f / l : 0.0f;
      bf16* out = p.O + (long long)bz * p.o_bstride + (long long)row * p.ldo + half * OC;
#pragma unroll 1
      for (int c = 0; c < OC / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(tmem + lane_off + TM_O + half * OC + c * 32, o);
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
      // partial row: balanced mode -> [slot][row in tile]; fixed splits -> [batch][split][query row]
      const long long prow = BAL ? (long long)S.slot * BM + rl : ((long long)bz * p.splits + blockIdx.y) * p.Nq + row;
      if (!BAL && CL == 1) {
        // fixed splits: the f32 partial tile leaves through the TMA engine.  thread = row stores to global memory touch 32
        // cache lines per warp instruction (8 k wavefronts for this 128 KB tile); here each thread drops its row into the dead
        // Q / K buffers in the layout of [128 rows x 32 floats] 128B-swizzled boxes and one thread issues DV / 32 tensor stores
        // (every MMA of the CTA has completed -- o_done -- and every TMA load has been consumed by one)
#pragma unroll 1
        for (int c = 0; c < OC / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem + lane_off + TM_O + half * OC + c * 32, o);
          uint8_t* st = smem + (half * (OC / 32) + c) * (BM * 128) + rl * 128;
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(st + ((i ^ (rl & 7)) << 4)) =
                make_float4(__uint_as_float(o[4 * i]), __uint_as_float(o[4 * i + 1]), __uint_as_float(o[4 * i + 2]), __uint_as_float(o[4 * i + 3]));
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) {
#pragma unroll 1
          for (int bx = 0; bx < DV / 32; ++bx) tma_store_3d(smem + bx * (BM * 128), &tmP, bx * 32, S.q0, bz * p.splits + (int)blockIdx.y);
          tma_store_commit();
          tma_store_wait_read();
        }
      } else {
      float* po = p.part_o + prow * DV + half * OC;
#pragma unroll 1
      for (int c = 0; c < OC / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(tmem + lane_off + TM_O + half * OC + c * 32, o);
        tc_wait_ld();
        if (row_ok) {
          float4* o4 = reinterpret_cast<float4*>(po + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o4[i] = make_float4(__uint_as_float(o[4 * i]), __uint_as_float(o[4 * i + 1]), __uint_as_float(o[4 * i + 2]),
                                __uint_as_float(o[4 * i + 3]));
        }
      }
      }
      if (row_ok && half == 0) {
        p.part_ml[prow * 2] = m_used;
        p.part_ml[prow * 2 + 1] = l;
      }
    }
    if (BAL) {               // O has left TMEM: the MMA warp may start the next segment's PV accumulation
      tc_fence_before();
      mbar_arrive(o_free);
    }
    jg0 += n;
    }  // segments
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // do not exit while the peer may still multicast into this CTA
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// Merge partials: one warp per query row, lane owns DV/32 channels (8 or 2).  The partial rows of a query row (at most
// 8) are addressed first, then ALL their (m, l) pairs and ALL their O vectors are loaded before any arithmetic: two L2
// round trips per row instead of two per partial (the loops over partials used to be latency chains: 9.8 us for 6).
constexpr int MAX_PARTS = 8;
template <int DV>
__device__ __forceinline__ void combine_row(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                            const long long (&prow)[MAX_PARTS], int cnt, int lane, bf16* __restrict__ dst) {
  constexpr int CPL = DV / 32;   // channels per lane
  float2 ml[MAX_PARTS];
#pragma unroll
  for (int k = 0; k < MAX_PARTS; ++k)
    ml[k] = k < cnt ? *reinterpret_cast<const float2*>(part_ml + prow[k] * 2) : make_float2(-INFINITY, 0.f);
  float pv[MAX_PARTS][CPL];
#pragma unroll
  for (int k = 0; k < MAX_PARTS; ++k) {
    if (k < cnt) {
      const float* src = part_o + prow[k] * DV + lane * CPL;
      if constexpr (CPL == 8) {
        const float4 a = reinterpret_cast<const float4*>(src)[0], d = reinterpret_cast<const float4*>(src)[1];
        pv[k][0] = a.x; pv[k][1] = a.y; pv[k][2] = a.z; pv[k][3] = a.w;
        pv[k][4] = d.x; pv[k][5] = d.y; pv[k][6] = d.z; pv[k][7] = d.w;
      } else {
        const float2 a = *reinterpret_cast<const float2*>(src);
        pv[k][0] = a.x; pv[k][1] = a.y;
      }
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < MAX_PARTS; ++k) m = fmaxf(m, ml[k].x);
  float acc[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) acc[i] = 0.f;
  float l = 0.0f;
#pragma unroll
  for (int k = 0; k < MAX_PARTS; ++k) {
    if (k < cnt) {
      const float w = exp2f(ml[k].x - m);
      l += w * ml[k].y;
#pragma unroll
      for (int i = 0; i < CPL; ++i) acc[i] += w * pv[k][i];
    }
  }
  const float inv = l > 0.0f ? 1.0f / l : 0.0f;
  if constexpr (CPL == 8) {
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                   pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
  } else {
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(acc[0] * inv, acc[1] * inv);
  }
}

template <int DV>
__global__ void attn_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, int B, int Nq,
                                    int splits, bf16* __restrict__ O, long long ldo, long long o_bstride) {
  pdl_enter();
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (long long)B * Nq) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(gw / Nq);
  const int row = (int)(gw % Nq);
  long long prow[MAX_PARTS];
#pragma unroll
  for (int k = 0; k < MAX_PARTS; ++k) prow[k] = ((long long)b * splits + (k < splits ? k : 0)) * Nq + row;
  combine_row<DV>(part_o, part_ml, prow, splits, lane,
                  O + (long long)b * o_bstride + (long long)row * ldo + lane * (DV / 32));
}

// Balanced mode: query tile qt received one partial from every CTA whose unit range overlaps [qt*ntiles, (qt+1)*ntiles);
// CTA c owns units [c*U/G, (c+1)*U/G) and its partial for qt is its segment number (qt - first query tile of c).
// The launcher works that out on the host: entry qt = first CTA | count << 8 | (2-bit segment number per partial) << 16
// (the first device version did it with 64-bit divisions per row: 940 instructions per row in ncu).
struct BalTable {
  uint32_t e[3 * 148];
};
template <int DV>
__global__ void attn_combine_bal_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, int B, int Nq,
                                        int qtiles, const BalTable tab, bf16* __restrict__ O, long long ldo,
                                        long long o_bstride) {
  pdl_enter();
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= B * Nq) return;
  const int lane = threadIdx.x & 31;
  const int b = gw / Nq;
  const int row = gw - b * Nq;
  const int rl = row % BM;
  const uint32_t e = tab.e[b * qtiles + row / BM];
  const int c_first = e & 0xff, cnt = (e >> 8) & 0xff;
  long long prow[MAX_PARTS];
#pragma unroll
  for (int k = 0; k < MAX_PARTS; ++k) {
    const int kk = k < cnt ? k : 0;
    prow[k] = (long long)(MAX_SEGS * (c_first + kk) + ((e >> (16 + 2 * kk)) & 3)) * BM + rl;
  }
  combine_row<DV>(part_o, part_ml, prow, cnt, lane,
                  O + (long long)b * o_bstride + (long long)row * ldo + lane * (DV / 32));
}

constexpr int BAL_CTAS = 148;   // one persistent CTA per SM of a B200

}  // namespace

long long* g_attn_trace = nullptr;  // dev-only timeline buffer (3*64*8 int64), see tools/trace_attention.py
int g_attn_balanced = 1;  // 1: balanced mode may be picked by attn_pick_splits (vls_set_tuning "attn_balanced")
int g_attn_v_rows = 1;   // memory cross-attention: 1 = V read as bank rows (MN-major operand), 0 = transposed copy (K-major)
int g_attn_cluster = 1;  // 1: private K/V loads; 2: pairs of query tiles multicast K/V (vls_set_tuning "attn_cluster")

// balanced mode is possible when every CTA gets at least one unit, at most MAX_SEGS segments, a query tile at most
// MAX_PARTS partials, and the combine table has room
static bool bal_ok(int qt, int ntiles) {
  const long long U = (long long)qt * ntiles;
  if (U < BAL_CTAS || qt > 3 * BAL_CTAS) return false;
  const long long per = (U + BAL_CTAS - 1) / BAL_CTAS;          // units per CTA (upper bound)
  if ((per + ntiles - 1) / ntiles + 1 > MAX_SEGS) return false;  // segments per CTA
  if (ntiles / (U / BAL_CTAS) + 2 > MAX_PARTS) return false;     // partials per query tile
  return true;
}

// splits == 0 selects the balanced ("stream-K") mode: MAX_SEGS partial slots per persistent CTA
size_t attn_workspace_bytes(int B, int Nq, int splits, int dv) {
  if (splits == 0) return align256((size_t)BAL_CTAS * MAX_SEGS * BM * dv * 4) + align256((size_t)BAL_CTAS * MAX_SEGS * BM * 2 * 4);
  if (splits <= 1) return 0;
  return align256((size_t)B * splits * Nq * dv * 4) + align256((size_t)B * splits * Nq * 2 * 4);
}

size_t attn_part_ml_offset(int B, int Nq, int splits, int dv) {   // byte offset of the (m, l) partials inside the workspace
  if (splits == 0) return align256((size_t)BAL_CTAS * MAX_SEGS * BM * dv * 4);
  return align256((size_t)B * splits * Nq * dv * 4);
}

int g_attn_bal_min_tiles = 64;   // balanced mode only for at least this many key tiles (vls_set_tuning "attn_bal_min_tiles")

int attn_pick_splits(int B, int Nq, int Nk) {
  const int qtiles = (Nq + BM - 1) / BM;
  const int ntiles = (Nk + BN - 1) / BN;
  int s = 148 / (qtiles * B > 0 ? qtiles * B : 1);
  if (s < 1) s = 1;
  if (s > 8) s = 8;
  while (s > 1 && ntiles / s < 4) --s;  // keep at least 4 KV tiles per split
  // long key sequences whose fixed split leaves SMs idle (B=1: 32 query tiles x 4 splits = 128 of 148; B=8: 256 CTAs =
  // 1.73 waves): deal the (query tile, key tile) units out evenly instead.
  // Only where the fixed path needs KV splits (and therefore partials + a combine) anyway: with many query tiles
  // (s == 1) it writes bf16 outputs directly, and trading that for f32 partials cost 17 % at B=8 (measured).
  if (g_attn_balanced && s >= 2 && ntiles >= g_attn_bal_min_tiles && bal_ok(qtiles * B, ntiles)) {
    const long long ctas = (long long)qtiles * B * s;
    const long long waves = (ctas + BAL_CTAS - 1) / BAL_CTAS;
    if (ctas * 100 < waves * BAL_CTAS * 94) return 0;     // the fixed split would leave > 6 % of the SM-waves idle
  }
  return s;
}

int attn_pick_splits_for(int B, int Nq, int Nk, int dv, int v_rows) {
  if (dv == 64 && v_rows && g_attn_x2) return attn_x2_pick_splits(B, Nq, Nk);
  return attn_pick_splits(B, Nq, Nk);
}

namespace {

template <int CL, bool BAL, int DV, bool VMN>
int launch_variant(const AttnArgs& a, const AttnParams& p, int qtiles, cudaStream_t stream) {
  using C_ = ACfg<DV>;
  CUtensorMap tmQ, tmK, tmV;
  VLS_TRY(make_tmap_bf16(&tmQ, a.Q, D, a.Nq, a.B, a.ldq, a.q_bstride, BM));
  VLS_TRY(make_tmap_bf16(&tmK, a.K, D, a.Nk, a.B, a.ldk, a.k_bstride, CL > 1 ? BN / 2 : BN));
  if (VMN) {
    VLS_TRY(make_tmap_bf16(&tmV, a.Vt, DV, a.Nk, a.B, a.ldvt, a.vt_bstride, BN));                     // rows [Nk][64]
  } else {
    VLS_TRY(make_tmap_bf16(&tmV, a.Vt, a.Nk, DV, a.B, a.ldvt, a.vt_bstride, CL > 1 ? DV / 2 : DV));   // V^T [DV][Nk]
  }
  CUtensorMap tmP = tmQ;   // fixed splits: partial tiles f32 [B * splits][Nq][DV], stored by TMA (rows beyond Nq are clipped)
  if (!BAL && CL == 1 && a.splits > 1)
    VLS_TRY(make_tmap_f32(&tmP, a.part_o, DV, a.Nq, (uint64_t)a.B * a.splits, DV, (long long)a.Nq * DV, BM));
  static unsigned long long attr_set = 0;   // one flag word per instantiation
  if (first_use_on_device(&attr_set)) {
    auto kern = attn_fwd_kernel<CL, BAL, DV, VMN>;
    VLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = BAL ? dim3(BAL_CTAS, 1, 1) : dim3(qtiles, a.splits, a.B);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = C_::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  auto kern = attn_fwd_kernel<CL, BAL, DV, VMN>;
  VLS_CUDA(cudaLaunchKernelEx(&cfg, kern, tmQ, tmK, tmV, tmP, p));
  VLS_POST_LAUNCH(1);
  return 0;
}

template <int DV>
int launch_combine(const AttnArgs& a, int qtiles, int nt, bool bal, cudaStream_t stream) {
  const long long rows = (long long)a.B * a.Nq;
  const int wpb = 8;
  VLS_REQUIRE(rows < (1ll << 31), "attention: too many query rows");
  const dim3 grid((unsigned)((rows + wpb - 1) / wpb));
  if (bal) {
    BalTable tab;
    const long long U = (long long)a.B * qtiles * nt;
    for (int qt = 0; qt < a.B * qtiles; ++qt) {
      const long long first_u = (long long)qt * nt, last_u = first_u + nt - 1;
      const int c_first = (int)(((first_u + 1) * BAL_CTAS - 1) / U), c_last = (int)(((last_u + 1) * BAL_CTAS - 1) / U);
      VLS_REQUIRE(c_last - c_first + 1 <= MAX_PARTS, "attention: too many partials per query tile");
      uint32_t e = (uint32_t)c_first | ((uint32_t)(c_last - c_first + 1) << 8);
      for (int c = c_first; c <= c_last; ++c) {
        const int seg = qt - (int)((U * c / BAL_CTAS) / nt);   // segment number inside CTA c = tiles since its first one
        VLS_REQUIRE(seg >= 0 && seg < MAX_SEGS, "attention: segment index out of range");
        e |= (uint32_t)seg << (16 + 2 * (c - c_first));
      }
      tab.e[qt] = e;
    }
    VLS_CUDA(launch_k(attn_combine_bal_kernel<DV>, grid, dim3(wpb * 32), 0, stream, a.part_o, a.part_ml, a.B, a.Nq, qtiles,
                      tab, reinterpret_cast<bf16*>(a.O), a.ldo, a.o_bstride));
  } else {
    VLS_CUDA(launch_k(attn_combine_kernel<DV>, grid, dim3(wpb * 32), 0, stream, a.part_o, a.part_ml, a.B, a.Nq, a.splits,
                      reinterpret_cast<bf16*>(a.O), a.ldo, a.o_bstride));
  }
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace

int launch_attention(const AttnArgs& a, cudaStream_t stream) {
  VLS_REQUIRE(a.Q && a.K && a.Vt && a.O, "attention: null operand");
  VLS_REQUIRE(a.Nq > 0 && a.Nk > 0 && a.B > 0 && a.splits >= 0, "attention: bad shape");
  VLS_REQUIRE(a.dv == 256 || a.dv == 64, "attention: value dimension must be 256 or 64 (got %d)", a.dv);
  VLS_REQUIRE(!a.v_rows || a.dv == 64, "attention: row-major V needs dv == 64");
  VLS_REQUIRE(a.ldo % 8 == 0 && a.ldvt % 8 == 0, "attention: ldo / ldv must be multiples of 8");
  const int nt = (a.Nk + BN - 1) / BN;
  VLS_REQUIRE(a.splits <= nt, "attention: more KV splits (%d) than KV tiles (%d)", a.splits, nt);
  const bool x2 = a.dv == 64 && a.v_rows && g_attn_x2 && a.splits >= 1;
  VLS_REQUIRE(x2 || a.splits <= MAX_PARTS, "attention: at most %d KV splits", MAX_PARTS);
  VLS_REQUIRE(a.splits == 1 || (a.part_o && a.part_ml), "attention: split workspace missing");
  const int qtiles = (a.Nq + BM - 1) / BM;
  const bool bal = a.splits == 0;
  VLS_REQUIRE(!bal || bal_ok(qtiles * a.B, nt), "attention: shape not supported by the balanced mode");
  const int cl = (!bal && a.dv == 256 && qtiles % 2 == 0 && g_attn_cluster > 1) ? 2 : 1;
  AttnParams p;
  p.Nq = a.Nq; p.Nk = a.Nk; p.splits = a.splits;
  p.qtiles = qtiles; p.ntiles = nt; p.units = (long long)a.B * qtiles * nt;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.O = reinterpret_cast<bf16*>(a.O); p.ldo = a.ldo; p.o_bstride = a.o_bstride;
  p.part_o = a.part_o; p.part_ml = a.part_ml;
  p.trace = g_attn_trace;
  // the profiling bracket covers the attention kernel AND its combine launch (r1 verdict: the combine was outside)
  const int slot = a.Nk > a.Nq ? PROF_ATTN_CROSS : PROF_ATTN_SELF;
  prof_begin(slot, stream);
  int rc;
  if (a.dv == 64 && a.v_rows && g_attn_x2 && a.splits >= 1) {   // two query tiles per CTA, fixed KV splits (attn_x2.cu)
    rc = launch_attention_x2(a, stream);
    if (rc != 0) return rc;
    prof_end(slot, stream);
    return 0;
  }
  if (a.dv == 64 && a.v_rows)
    rc = bal ? launch_variant<1, true, 64, true>(a, p, qtiles, stream) : launch_variant<1, false, 64, true>(a, p, qtiles, stream);
  else if (a.dv == 64)
    rc = bal ? launch_variant<1, true, 64, false>(a, p, qtiles, stream) : launch_variant<1, false, 64, false>(a, p, qtiles, stream);
  else if (bal) rc = launch_variant<1, true, 256, false>(a, p, qtiles, stream);
  else if (cl > 1) rc = launch_variant<2, false, 256, false>(a, p, qtiles, stream);
  else rc = launch_variant<1, false, 256, false>(a, p, qtiles, stream);
  if (rc != 0) return rc;
  if (bal || a.splits > 1) {
    rc = a.dv == 64 ? launch_combine<64>(a, qtiles, nt, bal, stream) : launch_combine<256>(a, qtiles, nt, bal, stream);
    if (rc != 0) return rc;
  }
  prof_end(slot, stream);
  return 0;
}

}  // namespace vls
