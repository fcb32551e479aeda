// Token side of a two-way-transformer layer of the mask decoder in ONE thread-block-cluster kernel
// (sam/transformer.py:183-210 -- TwoWayAttentionBlock.forward -- and :127-132, the final token->image attention).
//
// r1 ran the <= 16 token rows through a chain of ~17 kernels per layer (small linears, LayerNorms, 8-token attentions,
// the token->image attention) of 2-13 us each: latency, not work.  Here a cluster of 8 CTAs (one per attention head) owns
// one batch element and walks the whole chain with cluster barriers between the stages:
//
//   SELF   q/k/v projection of head r (CTA r computes exactly the 3 x 32 columns its head needs: no exchange)
//          -> 8-head token self-attention -> all-gather -> output projection (+ residual) -> all-gather -> LayerNorm1
//   CROSS  q projection of head r -> token->image attention of head r over all T image tokens (keys / values are read
//          from head-major "planes" [head][T][16] bf16 written by the image-side projection GEMM: 32 contiguous bytes per
//          key, so the loads of a warp are 256 contiguous bytes) -> all-gather -> output projection + residual -> all-gather
//          -> LayerNorm2 (or the decoder's final LayerNorm)
//   MLP    256 -> 2048 (ReLU) -> 256 with the hidden units split over the CTAs (K-split second GEMM, reduce-scatter of the
//          partial sums over distributed shared memory) + residual -> all-gather -> LayerNorm3 -> k / v projections of the
//          image->token attention
//
// Every contraction is a warp-level tensor-core MMA (mma.sync.m16n8k16, bf16 x bf16 -> f32): the WEIGHT tile is the M x K
// operand (16 output columns), the token rows are the N dimension (8 per tile), so 8 tokens cost one MMA column block and
// the weights stream L2 -> registers exactly once per cluster (16-byte loads, two 64-byte row pieces per thread and k-step:
// the k index is permuted identically for both operands so that a thread's fragment is contiguous in memory).  Activations
// stay f32; as MMA operands they are split into bf16 hi + lo parts (two MMAs), which keeps the f32 x bf16 product exact to
// 2^-17.  The token->image attention is a flash-attention loop on the same instruction: S = Q K^T with the 16 (padded)
// token rows as M, P = exp2(S - max) packed to bf16 straight from the accumulator fragment into the A fragment of P V, V
// fragments transposed in registers with movmatrix (no shared-memory staging at all).
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int CLD = 8;            // CTAs per cluster = attention heads
constexpr int DT_THREADS = 512;
constexpr int DT_WARPS = DT_THREADS / 32;
constexpr int ROWS = 16;          // token rows held (Nt <= 16)
constexpr int OPS = 256 + 32;     // operand row pitch in bf16: 576 B = 64 (mod 128) -> conflict-free 16-byte fragment loads

struct DecTokParams {
  int Nt, T, flags;               // flags: DEC_TOK_*
  float eps;
  float* queries;                 // f32 [B][Nt][256], in / out
  const float* pe;                // f32 [B][Nt][256] (the initial tokens = query PE)
  const bf16 *sq_w, *sk_w, *sv_w, *so_w;
  const float *sq_b, *sk_b, *sv_b, *so_b, *n1_w, *n1_b;
  const bf16 *tq_w, *to_w;
  const float *tq_b, *to_b, *n2_w, *n2_b;
  const bf16* planes;             // bf16 [B][planes][T][16]
  long long planes_bstride;
  int kplane, vplane;             // K of head h = plane kplane + h, V = plane vplane + h
  const bf16 *m1_w, *m2_w;
  const float *m1_b, *m2_b, *n3_w, *n3_b;
  const bf16 *ik_w, *iv_w;
  const float *ik_b, *iv_b;
  float* kt;                      // f32 [B][Nt][128]
  float* vt;
  long long* trace;               // optional dev trace: 24 clock64 stamps of thread 0 of CTA (0, 0)
};

#define DT_TRACE(slot)                                                        \
  do {                                                                        \
    if (p.trace && tid == 0 && rank == 0 && blockIdx.y == 0) p.trace[slot] = clock64(); \
  } while (0)

struct __align__(16) DecTokSmem {
  float xs[ROWS][256];            // current token rows (replicated in every CTA)
  float pes[ROWS][256];
  float ybuf[ROWS][256];          // all-gather landing zone of the pre-LayerNorm rows
  bf16 opA[2][ROWS][OPS];         // MMA operand (hi, lo): x + pe
  bf16 opB[2][ROWS][OPS];         // x
  bf16 opC[2][ROWS][OPS];         // attention outputs (all-gather landing zone)
  bf16 opH[2][ROWS][OPS];         // this CTA's 256 hidden units
  float qkv[3][ROWS][33];         // q, k, v of this CTA's self-attention head
  float tqs[ROWS][16];            // scaled q of this CTA's token->image head
  float part[CLD][ROWS][32];      // partial sums of the second MLP GEMM: [source CTA][row][this CTA's 32 columns]
  float tpart[DT_WARPS][ROWS][20];  // per-warp flash-attention partials: 16 channels, max, sum
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// v = hi + lo with hi, lo bf16 (|v - hi - lo| <= 2^-17 |v|)
__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint16_t bf16_bits(bf16 v) { return *reinterpret_cast<uint16_t*>(&v); }

// ---- weight tiles.  A tile = 16 weight rows (output columns) x K; fragment element e of token tile nt -> (row, token):
//   e = 0: (g, 8 nt + 2t)   1: (g, 8 nt + 2t + 1)   2: (g + 8, 8 nt + 2t)   3: (g + 8, 8 nt + 2t + 1)
// The k index of a 32-wide step is permuted (thread t owns physical k = 8t .. 8t+7: 4 for each of the two MMAs), so that
// the weight fragment is ONE 16-byte load per row and the activation fragment one 16-byte shared-memory load.
// Loads and MMAs are separate calls: the weights of the NEXT stage are requested before the barrier that precedes it.
template <int K>
struct WTile {
  uint4 a0[K / 32], a1[K / 32];
};
template <int K>
__device__ __forceinline__ void load_w(WTile<K>& w, const bf16* __restrict__ W, long long ldw, int g, int t) {
  const uint4* wa = reinterpret_cast<const uint4*>(W + (long long)g * ldw + 8 * t);
  const uint4* wb = reinterpret_cast<const uint4*>(W + (long long)(g + 8) * ldw + 8 * t);
#pragma unroll
  for (int i = 0; i < K / 32; ++i) {
    w.a0[i] = __ldg(wa + 4 * i);
    w.a1[i] = __ldg(wb + 4 * i);
  }
}
// K0: first column of the operand rows this tile multiplies (K-split stages pass their slice offset)
template <int K, int NT8>
__device__ __forceinline__ void mma_w(const WTile<K>& w, const bf16 (*op)[ROWS][OPS], int k0, float (&out)[NT8][4], int g,
                                      int t) {
  float acc[4][NT8][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[c][nt][e] = 0.f;
#pragma unroll
  for (int i = 0; i < K / 32; ++i) {
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const uint4 bh = *reinterpret_cast<const uint4*>(&op[0][nt * 8 + g][k0 + 32 * i + 8 * t]);
      const uint4 bl = *reinterpret_cast<const uint4*>(&op[1][nt * 8 + g][k0 + 32 * i + 8 * t]);
      mma_bf16(acc[0][nt], w.a0[i].x, w.a1[i].x, w.a0[i].y, w.a1[i].y, bh.x, bh.y);
      mma_bf16(acc[1][nt], w.a0[i].z, w.a1[i].z, w.a0[i].w, w.a1[i].w, bh.z, bh.w);
      mma_bf16(acc[2][nt], w.a0[i].x, w.a1[i].x, w.a0[i].y, w.a1[i].y, bl.x, bl.y);
      mma_bf16(acc[3][nt], w.a0[i].z, w.a1[i].z, w.a0[i].w, w.a1[i].w, bl.z, bl.w);
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) out[nt][e] = (acc[0][nt][e] + acc[1][nt][e]) + (acc[2][nt][e] + acc[3][nt][e]);
}

__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bf16 h0, l0, h1, l1;
    split_bf16(v[2 * j], h0, l0);
    split_bf16(v[2 * j + 1], h1, l1);
    h[j] = uint32_t(bf16_bits(h0)) | (uint32_t(bf16_bits(h1)) << 16);
    l[j] = uint32_t(bf16_bits(l0)) | (uint32_t(bf16_bits(l1)) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct KV4 {   // 32 keys of one head: fragment pieces of 4 tiles of 8 keys
  uint2 k[4], v[4];
};

template <int NT8>
__global__ void __cluster_dims__(CLD, 1, 1) __launch_bounds__(DT_THREADS, 1) dec_tok_kernel(const DecTokParams p) {
  extern __shared__ uint8_t dt_smem_raw[];
  DecTokSmem& s = *reinterpret_cast<DecTokSmem*>((reinterpret_cast<uintptr_t>(dt_smem_raw) + 15) & ~uintptr_t(15));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.y, Nt = p.Nt;
  const int flags = p.flags;
  const bool first = (flags & DEC_TOK_FIRST) != 0;
  pdl_enter();
  DT_TRACE(0);
  if (p.trace && tid == 0 && rank == 0 && blockIdx.y == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[22] = (long long)gt;
  }

  // ---- stage 0: everything that does not depend on the token rows is requested first
  float* qglob = p.queries + (long long)b * Nt * 256;
  const int r0 = tid >> 5, c0 = (tid & 31) * 8;   // this thread's piece of the token rows: row r0, columns c0 .. c0 + 7
  float4 xin[2], pin[2];
  xin[0] = xin[1] = pin[0] = pin[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r0 < Nt) {
    const float4* xq = reinterpret_cast<const float4*>(qglob + r0 * 256 + c0);
    const float4* pq = reinterpret_cast<const float4*>(p.pe + ((long long)b * Nt + r0) * 256 + c0);
    xin[0] = xq[0]; xin[1] = xq[1];
    pin[0] = __ldg(pq); pin[1] = __ldg(pq + 1);
  }
  WTile<256> wt;     // full-K tile of the next "one tile per warp" stage (qkv / mlp1 / mlp2)
  WTile<32> ws;      // K-slice of the next K-split stage (o-proj, q, i2t k/v)
  KV4 kv0, kv1;      // token->image attention: double-buffered key / value fragments
  const int groups = p.T >> 5;
  const bf16* Kp = nullptr;
  const bf16* Vp = nullptr;
  auto attn_load = [&](KV4& buf, int grp) {
    const long long key0 = (long long)grp * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      buf.k[j] = __ldg(reinterpret_cast<const uint2*>(Kp + (key0 + 8 * j + g) * 16 + 4 * t));
      buf.v[j] = __ldg(reinterpret_cast<const uint2*>(Vp + (key0 + 8 * j + g) * 16 + 4 * t));
    }
  };
  if (flags & DEC_TOK_SELF) {
    if (warp < 6) {   // q, k, v columns [32 r, 32 r + 32) = head r: two 16-column tiles each
      const int which = warp >> 1, half = warp & 1;
      load_w<256>(wt, (which == 0 ? p.sq_w : which == 1 ? p.sk_w : p.sv_w) + (long long)(32 * rank + 16 * half) * 256, 256, g, t);
    }
  } else if (flags & DEC_TOK_CROSS) {
    if (warp < 8) load_w<32>(ws, p.tq_w + (long long)(16 * rank) * 256 + 32 * warp, 256, g, t);
    Kp = p.planes + (long long)b * p.planes_bstride + (long long)(p.kplane + (int)rank) * p.T * 16;
    Vp = p.planes + (long long)b * p.planes_bstride + (long long)(p.vplane + (int)rank) * p.T * 16;
    if (warp < groups) attn_load(kv0, warp);
  }
  {   // rows >= Nt of the gathered operands stay zero
    uint32_t* zc = reinterpret_cast<uint32_t*>(s.opC);
    uint32_t* zh = reinterpret_cast<uint32_t*>(s.opH);
    for (int i = tid; i < ROWS * OPS; i += DT_THREADS) { zc[i] = 0u; zh[i] = 0u; }
  }
  // xs / pes / operands of (row r0, columns c0..c0+7): opB = x, opA = x + pe (x alone for the first layer's self-attention)
  auto put_rows = [&](int row, int col, const float (&x)[8], const float (&pe)[8], bool with_pe) {
    uint4 hi, lo;
    split8(x, hi, lo);
    *reinterpret_cast<uint4*>(&s.opB[0][row][col]) = hi;
    *reinterpret_cast<uint4*>(&s.opB[1][row][col]) = lo;
    if (with_pe) {
      float xp[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xp[j] = x[j] + pe[j];
      split8(xp, hi, lo);
    }
    *reinterpret_cast<uint4*>(&s.opA[0][row][col]) = hi;
    *reinterpret_cast<uint4*>(&s.opA[1][row][col]) = lo;
  };
  {
    const float x[8] = {xin[0].x, xin[0].y, xin[0].z, xin[0].w, xin[1].x, xin[1].y, xin[1].z, xin[1].w};
    const float pe[8] = {pin[0].x, pin[0].y, pin[0].z, pin[0].w, pin[1].x, pin[1].y, pin[1].z, pin[1].w};
    *reinterpret_cast<float4*>(&s.xs[r0][c0]) = xin[0];
    *reinterpret_cast<float4*>(&s.xs[r0][c0 + 4]) = xin[1];
    *reinterpret_cast<float4*>(&s.pes[r0][c0]) = pin[0];
    *reinterpret_cast<float4*>(&s.pes[r0][c0 + 4]) = pin[1];
    put_rows(r0, c0, x, pe, !((flags & DEC_TOK_SELF) && first));
  }
  __syncthreads();
  DT_TRACE(1);
  // split-phase: every CTA of the cluster must be running (landing zones zeroed) before the FIRST remote store; the wait
  // sits right in front of it
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

  auto bcast_f32 = [&](float* local, float v) {
    const uint32_t a = smem_u32(local);
#pragma unroll
    for (int r = 0; r < CLD; ++r) st_cluster_f32(mapa_u32(a, (uint32_t)r), v);
  };
  // all-gather of an attention output as MMA operand: the even lane of a pair stores the two hi halves, the odd lane the two
  // lo halves (columns col & ~1, col | 1 of `row`)
  auto bcast_op_pair = [&](bf16 (*op)[ROWS][OPS], int row, int col, float v, bool active) {
    bf16 hi, lo;
    split_bf16(v, hi, lo);
    const uint32_t mine = (lane & 1) ? bf16_bits(lo) : bf16_bits(hi);       // what this lane contributes to its own word
    const uint32_t give = (lane & 1) ? bf16_bits(hi) : bf16_bits(lo);       // ... and to the neighbour's word
    const uint32_t got = __shfl_xor_sync(0xffffffffu, give, 1);
    const uint32_t word = (lane & 1) ? (got | (mine << 16)) : (mine | (got << 16));
    if (active) {
      const uint32_t a = smem_u32(&op[lane & 1][row][col & ~1]);
#pragma unroll
      for (int r = 0; r < CLD; ++r) asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(mapa_u32(a, (uint32_t)r)), "r"(word) : "memory");
    }
  };
  // xs = LayerNorm(ybuf) (warp = row, lane = 8 consecutive columns); optionally the MMA operands and the global copy
  auto layer_norm = [&](const float* w, const float* bb, bool ops, bool to_global) {
    if (warp < Nt) {
      const int c = 8 * lane;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bb + c)), b1 = __ldg(reinterpret_cast<const float4*>(bb + c + 4));
      const float4 y0 = *reinterpret_cast<const float4*>(&s.ybuf[warp][c]), y1 = *reinterpret_cast<const float4*>(&s.ybuf[warp][c + 4]);
      float v[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += v[i];
      const float mean = warp_sum(sum) * (1.0f / 256.0f);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] -= mean; q += v[i] * v[i]; }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256.0f) + p.eps);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] * rstd * wv[i] + bv[i];
      const float4 o0 = make_float4(v[0], v[1], v[2], v[3]), o1 = make_float4(v[4], v[5], v[6], v[7]);
      *reinterpret_cast<float4*>(&s.xs[warp][c]) = o0;
      *reinterpret_cast<float4*>(&s.xs[warp][c + 4]) = o1;
      if (ops) {
        const float4 p0 = *reinterpret_cast<const float4*>(&s.pes[warp][c]), p1 = *reinterpret_cast<const float4*>(&s.pes[warp][c + 4]);
        const float pe[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        put_rows(warp, c, v, pe, true);
      }
      if (to_global && (int)rank == (warp & 7)) {
        *reinterpret_cast<float4*>(qglob + warp * 256 + c) = o0;
        *reinterpret_cast<float4*>(qglob + warp * 256 + c + 4) = o1;
      }
    }
  };
  // K-split stage: warp w owns k-slice (w % NSL) of tile (w / NSL); partial sums go through shared memory
  // (scr = tpart, [warp][token][20]: conflict-free fragment stores) and are summed in slice order by thread (tile, token, col)
  float (*scr)[ROWS][20] = s.tpart;
  auto ks_mma = [&](const bf16 (*op)[ROWS][OPS], int kslice) {
    float o[NT8][4];
    mma_w<32, NT8>(ws, op, 32 * kslice, o, g, t);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) scr[warp][nt * 8 + 2 * t + (e & 1)][g + 8 * (e >> 1)] = o[nt][e];
  };
  auto ks_sum = [&](int tile, int nsl, int tok, int col) {
    float v = scr[tile * nsl][tok][col];
    for (int k = 1; k < nsl; ++k) v += scr[tile * nsl + k][tok][col];
    return v;
  };
  const int e_col = tid & 15, e_tok = (tid >> 4) & 15, e_tile = tid >> 8;   // thread -> element of a K-split stage's output
  bool waited0 = false;
  auto wait_cluster_start = [&]() {   // uniform across the CTA
    if (!waited0) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    waited0 = true;
  };

  // ================================================================== token self attention (sam/transformer.py:183-191)
  if (flags & DEC_TOK_SELF) {
    if (warp < 6) {
      const int which = warp >> 1, half = warp & 1;
      const float* bias = (which == 0 ? p.sq_b : which == 1 ? p.sk_b : p.sv_b) + 32 * rank + 16 * half;
      const float bz0 = __ldg(bias + g), bz1 = __ldg(bias + g + 8);
      float o[NT8][4];
      mma_w<256, NT8>(wt, which == 2 ? s.opB : s.opA, 0, o, g, t);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          s.qkv[which][nt * 8 + 2 * t + (e & 1)][16 * half + g + 8 * (e >> 1)] = o[nt][e] + ((e >> 1) ? bz1 : bz0);
    }
    // output projection, columns [32 r, 32 r + 32): 2 tiles x 8 k-slices; requested now, used after the all-gather
    load_w<32>(ws, p.so_w + (long long)(32 * rank + 16 * (warp >> 3)) * 256 + 32 * (warp & 7), 256, g, t);
    __syncthreads();
    DT_TRACE(2);
    float sa = 0.f;
    if (warp < Nt) {   // warp = query row, lane = key row for the scores, = channel for the output
      float sc = -INFINITY;
      if (lane < Nt) {
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          d0 += s.qkv[0][warp][c] * s.qkv[1][lane][c];
          d1 += s.qkv[0][warp][c + 1] * s.qkv[1][lane][c + 1];
        }
        sc = (d0 + d1) * 0.17677669529663687f;   // 1 / sqrt(32)
      }
      const float m = warp_max(sc);
      const float pr = lane < Nt ? __expf(sc - m) : 0.f;
      const float l = warp_sum(pr);
      for (int j = 0; j < Nt; ++j) sa += __shfl_sync(0xffffffffu, pr, j) * s.qkv[2][j][lane];
      sa /= l;
    }
    wait_cluster_start();
    bcast_op_pair(s.opC, warp < Nt ? warp : 0, 32 * (int)rank + lane, sa, warp < Nt);
    DT_TRACE(3);
    cluster_sync_all();
    DT_TRACE(4);
    ks_mma(s.opC, warp & 7);
    __syncthreads();
    if (e_tok < Nt) {
      const int col = 32 * (int)rank + 16 * e_tile + e_col;
      bcast_f32(&s.ybuf[e_tok][col], ks_sum(e_tile, 8, e_tok, e_col) + __ldg(p.so_b + col) + (first ? 0.f : s.xs[e_tok][col]));
    }
    DT_TRACE(5);
    cluster_sync_all();
    DT_TRACE(6);
    const bool more = (flags & (DEC_TOK_CROSS | DEC_TOK_MLP)) != 0;
    if (more && warp < 8) load_w<32>(ws, p.tq_w + (long long)(16 * rank) * 256 + 32 * warp, 256, g, t);
    layer_norm(p.n1_w, p.n1_b, more, !more);
    if (more && (flags & DEC_TOK_CROSS)) {
      Kp = p.planes + (long long)b * p.planes_bstride + (long long)(p.kplane + (int)rank) * p.T * 16;
      Vp = p.planes + (long long)b * p.planes_bstride + (long long)(p.vplane + (int)rank) * p.T * 16;
      if (warp < groups) attn_load(kv0, warp);
    }
    __syncthreads();
    DT_TRACE(7);
  }

  // ================================================================== tokens -> image attention (:193-198 / :127-132)
  if (flags & DEC_TOK_CROSS) {
    // q columns [16 r, 16 r + 16) = head r (1 tile x 8 k-slices), pre-scaled by 1/sqrt(16) * log2(e)
    if (warp < 8) ks_mma(s.opA, warp);
    if (warp < 8) load_w<32>(ws, p.to_w + (long long)(32 * rank + 16 * (warp >> 2)) * 128 + 32 * (warp & 3), 128, g, t);   // o-proj slices
    __syncthreads();
    if (tid < 256) s.tqs[e_tok][e_col] = (ks_sum(0, 8, e_tok, e_col) + __ldg(p.tq_b + 16 * rank + e_col)) * (0.25f * 1.4426950408889634f);
    __syncthreads();
    DT_TRACE(8);
    {
      // A fragments of Q (hi / lo): rows g, g + 8; physical channels 4t .. 4t+3
      uint32_t qh[4], ql[4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int row = g + 8 * rr;
        bf16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16((row < Nt) ? s.tqs[row][4 * t + j] : 0.f, h[j], l[j]);
        qh[rr] = uint32_t(bf16_bits(h[0])) | (uint32_t(bf16_bits(h[1])) << 16);
        qh[rr + 2] = uint32_t(bf16_bits(h[2])) | (uint32_t(bf16_bits(h[3])) << 16);
        ql[rr] = uint32_t(bf16_bits(l[0])) | (uint32_t(bf16_bits(l[1])) << 16);
        ql[rr + 2] = uint32_t(bf16_bits(l[2])) | (uint32_t(bf16_bits(l[3])) << 16);
      }
      float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
      float o[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[c][e] = 0.f;
      auto attn_compute = [&](const KV4& buf) {
        float sc[4][4];
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
          mma_bf16(sc[j], qh[0], qh[1], qh[2], qh[3], buf.k[j].x, buf.k[j].y);
          mma_bf16(sc[j], ql[0], ql[1], ql[2], ql[3], buf.k[j].x, buf.k[j].y);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
          if (NT8 == 2) mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        const float mn0 = fmaxf(m[0], mx0);
        const float corr0 = ex2_approx(m[0] - mn0);
        m[0] = mn0;
        l[0] *= corr0;
        o[0][0] *= corr0; o[0][1] *= corr0; o[1][0] *= corr0; o[1][1] *= corr0;
        if (NT8 == 2) {
          mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
          mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
          const float mn1 = fmaxf(m[1], mx1);
          const float corr1 = ex2_approx(m[1] - mn1);
          m[1] = mn1;
          l[1] *= corr1;
          o[0][2] *= corr1; o[0][3] *= corr1; o[1][2] *= corr1; o[1][3] *= corr1;
        }
#pragma unroll
        for (int st = 0; st < 2; ++st) {
          const int ja = 2 * st, jb = 2 * st + 1;
          const float pa0 = ex2_approx(sc[ja][0] - m[0]), pa1 = ex2_approx(sc[ja][1] - m[0]);
          const float pb0 = ex2_approx(sc[jb][0] - m[0]), pb1 = ex2_approx(sc[jb][1] - m[0]);
          l[0] += (pa0 + pa1) + (pb0 + pb1);
          float pa2 = 0.f, pa3 = 0.f, pb2 = 0.f, pb3 = 0.f;
          if (NT8 == 2) {
            pa2 = ex2_approx(sc[ja][2] - m[1]); pa3 = ex2_approx(sc[ja][3] - m[1]);
            pb2 = ex2_approx(sc[jb][2] - m[1]); pb3 = ex2_approx(sc[jb][3] - m[1]);
            l[1] += (pa2 + pa3) + (pb2 + pb3);
          }
          const uint32_t a0 = pack_bf16x2(pa0, pa1), a1 = pack_bf16x2(pa2, pa3);
          const uint32_t a2 = pack_bf16x2(pb0, pb1), a3 = pack_bf16x2(pb2, pb3);
          mma_bf16(o[0], a0, a1, a2, a3, movmatrix_trans(buf.v[ja].x), movmatrix_trans(buf.v[jb].x));
          mma_bf16(o[1], a0, a1, a2, a3, movmatrix_trans(buf.v[ja].y), movmatrix_trans(buf.v[jb].y));
        }
      };
#pragma unroll 1
      for (int grp = warp; grp < groups; grp += 2 * DT_WARPS) {   // kv0 of the first group was requested in stage 0
        const int g1 = grp + DT_WARPS, g2 = grp + 2 * DT_WARPS;
        if (g1 < groups) attn_load(kv1, g1);
        attn_compute(kv0);
        if (g1 < groups) {
          if (g2 < groups) attn_load(kv0, g2);
          attn_compute(kv1);
        }
      }
      // per-warp partial -> shared memory.  o[c][0..1]: row g, channels 4t + 2c + {0, 1}; o[c][2..3]: row g + 8
#pragma unroll
      for (int rr = 0; rr < NT8; ++rr) {
        float lr = l[rr];
        lr += __shfl_xor_sync(0xffffffffu, lr, 1);
        lr += __shfl_xor_sync(0xffffffffu, lr, 2);
        const int row = g + 8 * rr;
        *reinterpret_cast<float4*>(&s.tpart[warp][row][4 * t]) = make_float4(o[0][2 * rr], o[0][2 * rr + 1], o[1][2 * rr], o[1][2 * rr + 1]);
        if (t == 0) { s.tpart[warp][row][16] = m[rr]; s.tpart[warp][row][17] = lr; }
      }
    }
    __syncthreads();
    DT_TRACE(9);
    wait_cluster_start();
    {   // merge the 16 warps' partials; thread = (row, channel) for tid < 256
      const int row = tid >> 4, ch = tid & 15;
      float val = 0.f;
      const bool act = tid < 256 && row < Nt;
      if (act) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < DT_WARPS; ++w) M = fmaxf(M, s.tpart[w][row][16]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < DT_WARPS; ++w) {
          const float f = ex2_approx(s.tpart[w][row][16] - M);
          num += s.tpart[w][row][ch] * f;
          den += s.tpart[w][row][17] * f;
        }
        val = num / den;
      }
      if (tid < 256) bcast_op_pair(s.opC, act ? row : 0, 16 * (int)rank + ch, val, act);
    }
    if (flags & DEC_TOK_MLP) load_w<256>(wt, p.m1_w + (long long)(256 * (int)rank + 16 * warp) * 256, 256, g, t);   // mlp1 tile of this warp
    DT_TRACE(10);
    cluster_sync_all();
    DT_TRACE(11);
    if (warp < 8) ks_mma(s.opC, warp & 3);   // output projection (K = 128): 2 tiles x 4 k-slices
    __syncthreads();
    if (e_tok < Nt) {
      const int col = 32 * (int)rank + 16 * e_tile + e_col;
      bcast_f32(&s.ybuf[e_tok][col], ks_sum(e_tile, 4, e_tok, e_col) + __ldg(p.to_b + col) + s.xs[e_tok][col]);
    }
    DT_TRACE(12);
    cluster_sync_all();
    DT_TRACE(13);
    layer_norm(p.n2_w, p.n2_b, (flags & DEC_TOK_MLP) != 0, !(flags & DEC_TOK_MLP));
    __syncthreads();
    DT_TRACE(14);
  }

  // ================================================================== token MLP (:200-203) + i2t k / v projections (:205-208)
  if (flags & DEC_TOK_MLP) {
    {   // hidden units [256 r + 16 w, + 16)
      const int h0 = 256 * (int)rank + 16 * warp;
      const float bz0 = __ldg(p.m1_b + h0 + g), bz1 = __ldg(p.m1_b + h0 + g + 8);
      float o[NT8][4];
      mma_w<256, NT8>(wt, s.opB, 0, o, g, t);
      // second GEMM over this CTA's hidden slice: output columns [16 w, 16 w + 16), requested while the hidden units are stored
      load_w<256>(wt, p.m2_w + (long long)(16 * warp) * 2048 + 256 * (int)rank, 2048, g, t);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = g + 8 * (e >> 1), tok = nt * 8 + 2 * t + (e & 1);
          bf16 hi, lo;
          split_bf16(tok < Nt ? fmaxf(o[nt][e] + ((e >> 1) ? bz1 : bz0), 0.f) : 0.f, hi, lo);
          s.opH[0][tok][16 * warp + col] = hi;
          s.opH[1][tok][16 * warp + col] = lo;
        }
    }
    __syncthreads();
    DT_TRACE(15);
    {   // partial sums -> owner CTA w / 2
      float o[NT8][4];
      mma_w<256, NT8>(wt, s.opH, 0, o, g, t);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = 16 * (warp & 1) + g + 8 * (e >> 1), tok = nt * 8 + 2 * t + (e & 1);
          if (tok < Nt) st_cluster_f32(mapa_u32(smem_u32(&s.part[rank][tok][col]), (uint32_t)(warp >> 1)), o[nt][e]);
        }
    }
    // image->token attention k (x + pe) and v (x) columns [16 r, 16 r + 16): 2 tiles x 8 k-slices
    load_w<32>(ws, (warp < 8 ? p.ik_w : p.iv_w) + (long long)(16 * rank) * 256 + 32 * (warp & 7), 256, g, t);
    DT_TRACE(16);
    cluster_sync_all();
    DT_TRACE(17);
    {   // reduce this CTA's 32 columns in fixed source order, + bias + residual, all-gather
      const int tok = tid >> 5, col = tid & 31, cg = 32 * (int)rank + col;
      if (tok < Nt) {
        float v = s.part[0][tok][col];
#pragma unroll
        for (int r = 1; r < CLD; ++r) v += s.part[r][tok][col];
        bcast_f32(&s.ybuf[tok][cg], v + __ldg(p.m2_b + cg) + s.xs[tok][cg]);
      }
    }
    DT_TRACE(18);
    cluster_sync_all();
    DT_TRACE(19);
    layer_norm(p.n3_w, p.n3_b, true, true);
    __syncthreads();
    ks_mma(warp < 8 ? s.opA : s.opB, warp & 7);
    __syncthreads();
    if (e_tok < Nt) {
      const float* bias = (e_tile == 0 ? p.ik_b : p.iv_b) + 16 * rank;
      float* dst = (e_tile == 0 ? p.kt : p.vt) + (long long)b * Nt * 128 + 16 * rank;
      dst[e_tok * 128 + e_col] = ks_sum(e_tile, 8, e_tok, e_col) + __ldg(bias + e_col);
    }
    DT_TRACE(20);
  }
  wait_cluster_start();   // (a launch without any remote store still has to consume its arrive)
  if (p.trace && tid == 0 && rank == 0 && blockIdx.y == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[23] = (long long)gt;
    p.trace[21] = clock64();
  }
}

}  // namespace

long long* g_dec_trace = nullptr;   // dev-only: 24 int64 clock64 stamps per launch, consecutive launches (tools/trace_dec.py)
int g_dec_fused = 1;   // mask decoder: 1 = token side of a layer as cluster kernels, 0 = the chain of small kernels

bool dec_tok_supported(int Nt, int T) { return g_dec_fused && Nt >= 1 && Nt <= ROWS && T >= 32 && T % 32 == 0; }

int launch_dec_tok(const DecTokArgs& a, cudaStream_t stream) {
  VLS_REQUIRE(a.queries && a.pe && a.B > 0 && a.Nt >= 1 && a.Nt <= ROWS, "dec_tok: bad arguments (Nt = %d)", a.Nt);
  VLS_REQUIRE(!(a.flags & DEC_TOK_MLP) || (a.flags & DEC_TOK_CROSS), "dec_tok: the MLP stage runs behind the CROSS stage");
  VLS_REQUIRE(!(a.flags & DEC_TOK_CROSS) || (a.planes && a.T >= 32 && a.T % 32 == 0), "dec_tok: T must be a multiple of 32");
  DecTokParams p = {};
  p.Nt = a.Nt; p.T = a.T; p.flags = a.flags; p.eps = a.eps;
  p.queries = a.queries; p.pe = a.pe;
  auto h = [](const void* q) { return reinterpret_cast<const bf16*>(q); };
  if (a.flags & DEC_TOK_SELF) {
    VLS_REQUIRE(a.self_attn && a.n1_w && a.n1_b, "dec_tok: self-attention weights missing");
    p.sq_w = h(a.self_attn->q_w); p.sk_w = h(a.self_attn->k_w); p.sv_w = h(a.self_attn->v_w); p.so_w = h(a.self_attn->o_w);
    p.sq_b = a.self_attn->q_b; p.sk_b = a.self_attn->k_b; p.sv_b = a.self_attn->v_b; p.so_b = a.self_attn->o_b;
    p.n1_w = a.n1_w; p.n1_b = a.n1_b;
  }
  if (a.flags & DEC_TOK_CROSS) {
    VLS_REQUIRE(a.t2i && a.n2_w && a.n2_b, "dec_tok: token->image attention weights missing");
    p.tq_w = h(a.t2i->q_w); p.tq_b = a.t2i->q_b; p.to_w = h(a.t2i->o_w); p.to_b = a.t2i->o_b;
    p.n2_w = a.n2_w; p.n2_b = a.n2_b;
    p.planes = h(a.planes); p.planes_bstride = a.planes_bstride; p.kplane = a.kplane; p.vplane = a.vplane;
  }
  if (a.flags & DEC_TOK_MLP) {
    VLS_REQUIRE(a.m1_w && a.m1_b && a.m2_w && a.m2_b && a.n3_w && a.n3_b && a.i2t && a.kt && a.vt, "dec_tok: MLP weights missing");
    p.m1_w = h(a.m1_w); p.m1_b = a.m1_b; p.m2_w = h(a.m2_w); p.m2_b = a.m2_b; p.n3_w = a.n3_w; p.n3_b = a.n3_b;
    p.ik_w = h(a.i2t->k_w); p.ik_b = a.i2t->k_b; p.iv_w = h(a.i2t->v_w); p.iv_b = a.i2t->v_b;
    p.kt = a.kt; p.vt = a.vt;
  }
  if (g_dec_trace) { p.trace = g_dec_trace; g_dec_trace += 24; }
  constexpr size_t SMEM = sizeof(DecTokSmem) + 16;
  static_assert(SMEM <= 232448, "dec_tok: shared memory");
  static unsigned long long attr1 = 0, attr2 = 0;
  if (a.Nt <= 8) {
    if (first_use_on_device(&attr1)) VLS_CUDA(cudaFuncSetAttribute(dec_tok_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    VLS_CUDA(launch_k(dec_tok_kernel<1>, dim3(CLD, a.B), dim3(DT_THREADS), SMEM, stream, p));
  } else {
    if (first_use_on_device(&attr2)) VLS_CUDA(cudaFuncSetAttribute(dec_tok_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    VLS_CUDA(launch_k(dec_tok_kernel<2>, dim3(CLD, a.B), dim3(DT_THREADS), SMEM, stream, p));
  }
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
