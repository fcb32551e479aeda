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

// parameter block staged in shared memory at kernel start (a dependent global load costs ~330 cycles each time)
enum { P_N1W = 0, P_N1B = 256, P_N2W = 512, P_N2B = 768, P_N3W = 1024, P_N3B = 1280, P_SO = 1536, P_TO = 1568, P_M2 = 1600,
       P_TQ = 1632, P_IK = 1648, P_IV = 1664, P_M1 = 1680, P_SQ = 1936, P_SK = 1968, P_SV = 2000, P_END = 2032 };

struct __align__(16) DecTokSmem {
  float xs[ROWS][256];            // current token rows (replicated in every CTA)
  float pes[ROWS][256];
  float ybuf[ROWS][256];          // all-gather landing zone of the pre-LayerNorm rows (+ local staging of the MLP partials)
  bf16 opA[2][ROWS][OPS];         // MMA operand (hi, lo): x + pe
  bf16 opB[2][ROWS][OPS];         // x
  bf16 opC[2][ROWS][OPS];         // attention outputs (all-gather landing zone)
  bf16 opH[2][ROWS][OPS];         // this CTA's 256 hidden units
  float qkv[3][ROWS][33];         // q, k, v of this CTA's self-attention head (q rows are overwritten by the head's output)
  float tqs[ROWS][16];            // scaled q of this CTA's token->image head (then the head's output)
  float part[CLD][ROWS][32];      // partial sums of the second MLP GEMM: [source CTA][row][this CTA's 32 columns]
  float tpart[DT_WARPS][ROWS][20];  // per-warp flash-attention partials: 16 channels, max, sum; also the K-split scratch
  float prm[P_END];               // LayerNorm weights and this CTA's bias slices
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
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// one 16-byte piece to the same shared-memory offset of every CTA of the cluster (remote stores cost ~10 cycles of issue
// per warp instruction whatever their width: measured, tools/micro/cluster_latency.cu)
__device__ __forceinline__ void bcast_v4(const void* local, uint4 v) {
  const uint32_t a = smem_u32(local);
#pragma unroll
  for (int r = 0; r < CLD; ++r) st_cluster_v4(mapa_u32(a, (uint32_t)r), v);
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
// mma.sync issues at 16 cycles per instruction and SM sub-partition on B200 (measured), so the MMA count matters: XLO = false
// drops the lo half of the activations (used where they are bf16-rounded anyway).
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
// k0: first column of the operand rows this tile multiplies (K-split stages pass their slice offset)
template <int K, int NT8, bool XLO>
__device__ __forceinline__ void mma_w(const WTile<K>& w, const bf16 (*op)[ROWS][OPS], int k0, float (&out)[NT8][4], int g,
                                      int t) {
  constexpr int CH = XLO ? 4 : 2;
  float acc[CH][NT8][4];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[c][nt][e] = 0.f;
#pragma unroll
  for (int i = 0; i < K / 32; ++i) {
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const uint4 bh = *reinterpret_cast<const uint4*>(&op[0][nt * 8 + g][k0 + 32 * i + 8 * t]);
      mma_bf16(acc[0][nt], w.a0[i].x, w.a1[i].x, w.a0[i].y, w.a1[i].y, bh.x, bh.y);
      mma_bf16(acc[1][nt], w.a0[i].z, w.a1[i].z, w.a0[i].w, w.a1[i].w, bh.z, bh.w);
      if (XLO) {
        const uint4 bl = *reinterpret_cast<const uint4*>(&op[1][nt * 8 + g][k0 + 32 * i + 8 * t]);
        mma_bf16(acc[CH - 2][nt], w.a0[i].x, w.a1[i].x, w.a0[i].y, w.a1[i].y, bl.x, bl.y);
        mma_bf16(acc[CH - 1][nt], w.a0[i].z, w.a1[i].z, w.a0[i].w, w.a1[i].w, bl.z, bl.w);
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = acc[0][nt][e] + acc[1][nt][e];
      if (XLO) v += acc[CH - 2][nt][e] + acc[CH - 1][nt][e];
      out[nt][e] = v;
    }
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

struct KV2 {   // 32 keys of one head = 2 tiles of 16 keys: k[j] / v[j] = (key 16 j' + g (j even) or + g + 8 (j odd), channels 4t..4t+3)
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
  float4 prm_in = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    const int i = 4 * tid;
    const float* src = nullptr;
    if (i < P_N1B) src = p.n1_w ? p.n1_w + i : nullptr;
    else if (i < P_N2W) src = p.n1_b ? p.n1_b + (i - P_N1B) : nullptr;
    else if (i < P_N2B) src = p.n2_w ? p.n2_w + (i - P_N2W) : nullptr;
    else if (i < P_N3W) src = p.n2_b ? p.n2_b + (i - P_N2B) : nullptr;
    else if (i < P_N3B) src = p.n3_w ? p.n3_w + (i - P_N3W) : nullptr;
    else if (i < P_SO) src = p.n3_b ? p.n3_b + (i - P_N3B) : nullptr;
    else if (i < P_TO) src = p.so_b ? p.so_b + 32 * rank + (i - P_SO) : nullptr;
    else if (i < P_M2) src = p.to_b ? p.to_b + 32 * rank + (i - P_TO) : nullptr;
    else if (i < P_TQ) src = p.m2_b ? p.m2_b + 32 * rank + (i - P_M2) : nullptr;
    else if (i < P_IK) src = p.tq_b ? p.tq_b + 16 * rank + (i - P_TQ) : nullptr;
    else if (i < P_IV) src = p.ik_b ? p.ik_b + 16 * rank + (i - P_IK) : nullptr;
    else if (i < P_M1) src = p.iv_b ? p.iv_b + 16 * rank + (i - P_IV) : nullptr;
    else if (i < P_SQ) src = p.m1_b ? p.m1_b + 256 * rank + (i - P_M1) : nullptr;
    else if (i < P_SK) src = p.sq_b ? p.sq_b + 32 * rank + (i - P_SQ) : nullptr;
    else if (i < P_SV) src = p.sk_b ? p.sk_b + 32 * rank + (i - P_SK) : nullptr;
    else if (i < P_END) src = p.sv_b ? p.sv_b + 32 * rank + (i - P_SV) : nullptr;
    if (src) prm_in = __ldg(reinterpret_cast<const float4*>(src));
  }
  WTile<256> wt;     // full-K tile of the next "one tile per warp" stage (qkv / mlp1 / mlp2)
  WTile<32> ws;      // K-slice of the next K-split stage (o-proj, q, i2t k/v)
  KV2 kv0, kv1;      // token->image attention: double-buffered key / value fragments
  const int groups = p.T >> 5;
  const bf16* Kp = nullptr;
  const bf16* Vp = nullptr;
  auto attn_load = [&](KV2& buf, int grp) {
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
    if (4 * tid < P_END) *reinterpret_cast<float4*>(&s.prm[4 * tid]) = prm_in;
  }
  __syncthreads();
  DT_TRACE(1);
  // split-phase: every CTA of the cluster must be running (landing zones zeroed) before the FIRST remote store; the wait
  // sits right in front of it
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

  // all-gather of 8 consecutive columns of an attention output as MMA operand (hi + lo): two 16-byte stores per CTA
  auto bcast_op8 = [&](bf16 (*op)[ROWS][OPS], int row, int col, const float (&v)[8]) {
    uint4 hi, lo;
    split8(v, hi, lo);
    bcast_v4(&op[0][row][col], hi);
    bcast_v4(&op[1][row][col], lo);
  };
  // xs = LayerNorm(ybuf) (warp = row, lane = 8 consecutive columns); optionally the MMA operands and the global copy
  auto layer_norm = [&](int pw, int pb, bool ops, bool to_global) {
    if (warp < Nt) {
      const int c = 8 * lane;
      const float4 w0 = *reinterpret_cast<const float4*>(&s.prm[pw + c]), w1 = *reinterpret_cast<const float4*>(&s.prm[pw + c + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&s.prm[pb + c]), b1 = *reinterpret_cast<const float4*>(&s.prm[pb + c + 4]);
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
  // (scr = tpart, [warp][token][20]: conflict-free fragment stores) and are summed in slice order by thread
  // (tile, token, 4 columns) -- tid < 128
  float (*scr)[ROWS][20] = s.tpart;
  auto ks_mma = [&](const bf16 (*op)[ROWS][OPS], int kslice) {
    float o[NT8][4];
    mma_w<32, NT8, true>(ws, op, 32 * kslice, o, g, t);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) scr[warp][nt * 8 + 2 * t + (e & 1)][g + 8 * (e >> 1)] = o[nt][e];
  };
  const int e_cg = (tid & 3) * 4, e_tok = (tid >> 2) & 15, e_tile = (tid >> 6) & 1;
  auto ks_sum4 = [&](int tile, int nsl) {
    float4 v = *reinterpret_cast<const float4*>(&scr[tile * nsl][e_tok][e_cg]);
    for (int k = 1; k < nsl; ++k) {
      const float4 u = *reinterpret_cast<const float4*>(&scr[tile * nsl + k][e_tok][e_cg]);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    return v;
  };
  // pre-LayerNorm rows: columns [32 r + 16 tile + cg, + 4) of token e_tok = K-split sum + bias + residual -> every CTA's ybuf
  auto ks_bcast_rows = [&](int nsl, int pbias, bool residual) {
    if (tid < 128 && e_tok < Nt) {
      float4 v = ks_sum4(e_tile, nsl);
      const int cl = 16 * e_tile + e_cg, col = 32 * (int)rank + cl;
      const float4 bz = *reinterpret_cast<const float4*>(&s.prm[pbias + cl]);
      v.x += bz.x; v.y += bz.y; v.z += bz.z; v.w += bz.w;
      if (residual) {
        const float4 xr = *reinterpret_cast<const float4*>(&s.xs[e_tok][col]);
        v.x += xr.x; v.y += xr.y; v.z += xr.z; v.w += xr.w;
      }
      bcast_v4(&s.ybuf[e_tok][col], make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
    }
  };
  bool waited0 = false;
  auto wait_cluster_start = [&]() {   // uniform across the CTA
    if (!waited0) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    waited0 = true;
  };

  // ================================================================== token self attention (sam/transformer.py:183-191)
  if (flags & DEC_TOK_SELF) {
    if (warp < 6) {
      const int which = warp >> 1, half = warp & 1;
      const float* bias = &s.prm[P_SQ + 32 * which + 16 * half];
      const float bz0 = bias[g], bz1 = bias[g + 8];
      float o[NT8][4];
      mma_w<256, NT8, true>(wt, which == 2 ? s.opB : s.opA, 0, o, g, t);
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
      float sa = 0.f;
      for (int j = 0; j < Nt; ++j) sa += __shfl_sync(0xffffffffu, pr, j) * s.qkv[2][j][lane];
      __syncwarp();
      s.qkv[0][warp][lane] = sa / l;   // only this warp reads q row `warp`
    }
    __syncthreads();
    wait_cluster_start();
    if (tid < 4 * Nt) {   // (row, 8 columns) -> operand pieces in every CTA
      const int row = tid >> 2, c8 = (tid & 3) * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = s.qkv[0][row][c8 + j];
      bcast_op8(s.opC, row, 32 * (int)rank + c8, v);
    }
    DT_TRACE(3);
    cluster_sync_all();
    DT_TRACE(4);
    ks_mma(s.opC, warp & 7);
    __syncthreads();
    ks_bcast_rows(8, P_SO, !first);
    DT_TRACE(5);
    cluster_sync_all();
    DT_TRACE(6);
    const bool more = (flags & (DEC_TOK_CROSS | DEC_TOK_MLP)) != 0;
    if (more && warp < 8) load_w<32>(ws, p.tq_w + (long long)(16 * rank) * 256 + 32 * warp, 256, g, t);
    layer_norm(P_N1W, P_N1B, more, !more);
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
    if (tid < 64) {
      float4 v = ks_sum4(0, 8);
      const float4 bz = *reinterpret_cast<const float4*>(&s.prm[P_TQ + e_cg]);
      const float sc = 0.25f * 1.4426950408889634f;
      *reinterpret_cast<float4*>(&s.tqs[e_tok][e_cg]) = make_float4((v.x + bz.x) * sc, (v.y + bz.y) * sc, (v.z + bz.z) * sc, (v.w + bz.w) * sc);
    }
    __syncthreads();
    DT_TRACE(8);
    {
      // Flash attention with the KEYS as the M dimension: S^T[16 keys][8 tokens] = K_h Q_h^T (one MMA per 16 keys and token
      // tile, + one for the lo half of q), P^T = exp2(S^T - max) transposed in registers (movmatrix) into the B fragment of
      // O^T[16 channels][8 tokens] += V_h^T P^T, whose A fragment is V transposed the same way.  Channels are permuted:
      // thread t loads physical channels 4t..4t+3 of a key (one 8-byte load).
      // B fragments of Q^T (hi / lo): token n = g (+ 8 per token tile), physical channels 4t .. 4t+3
      uint32_t qh[NT8][2], ql[NT8][2];
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        const int row = nt * 8 + g;
        bf16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16((row < Nt) ? s.tqs[row][4 * t + j] : 0.f, h[j], l[j]);
        qh[nt][0] = uint32_t(bf16_bits(h[0])) | (uint32_t(bf16_bits(h[1])) << 16);
        qh[nt][1] = uint32_t(bf16_bits(h[2])) | (uint32_t(bf16_bits(h[3])) << 16);
        ql[nt][0] = uint32_t(bf16_bits(l[0])) | (uint32_t(bf16_bits(l[1])) << 16);
        ql[nt][1] = uint32_t(bf16_bits(l[2])) | (uint32_t(bf16_bits(l[3])) << 16);
      }
      // per thread: tokens 8 nt + 2t, 8 nt + 2t + 1 (columns of S^T and O^T)
      float m[NT8][2], l[NT8][2], o[NT8][4];
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        m[nt][0] = m[nt][1] = -INFINITY;
        l[nt][0] = l[nt][1] = 0.f;
        o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
      }
      auto attn_compute = [&](const KV2& buf) {
        float sc[2][NT8][4];   // [key tile][token tile]: (key g, tok 2t) (key g, tok 2t+1) (key g+8, tok 2t) (key g+8, tok 2t+1)
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
          for (int nt = 0; nt < NT8; ++nt) {
            sc[kt][nt][0] = sc[kt][nt][1] = sc[kt][nt][2] = sc[kt][nt][3] = 0.f;
            mma_bf16(sc[kt][nt], buf.k[2 * kt].x, buf.k[2 * kt + 1].x, buf.k[2 * kt].y, buf.k[2 * kt + 1].y, qh[nt][0], qh[nt][1]);
            mma_bf16(sc[kt][nt], buf.k[2 * kt].x, buf.k[2 * kt + 1].x, buf.k[2 * kt].y, buf.k[2 * kt + 1].y, ql[nt][0], ql[nt][1]);
          }
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
          float mx0 = fmaxf(fmaxf(sc[0][nt][0], sc[0][nt][2]), fmaxf(sc[1][nt][0], sc[1][nt][2]));
          float mx1 = fmaxf(fmaxf(sc[0][nt][1], sc[0][nt][3]), fmaxf(sc[1][nt][1], sc[1][nt][3]));
#pragma unroll
          for (int sh = 4; sh < 32; sh <<= 1) {   // over the 8 key lanes g
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, sh));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, sh));
          }
          const float mn0 = fmaxf(m[nt][0], mx0), mn1 = fmaxf(m[nt][1], mx1);
          const float cr0 = ex2_approx(m[nt][0] - mn0), cr1 = ex2_approx(m[nt][1] - mn1);
          m[nt][0] = mn0; m[nt][1] = mn1;
          l[nt][0] *= cr0; l[nt][1] *= cr1;
          o[nt][0] *= cr0; o[nt][1] *= cr1; o[nt][2] *= cr0; o[nt][3] *= cr1;
        }
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          // A = V^T: (channel g, keys 2t, 2t+1) (channel g + 8, ...) (channel g, keys 2t + 8, 2t + 9) (channel g + 8, ...)
          const uint32_t a0 = movmatrix_trans(buf.v[2 * kt].x), a1 = movmatrix_trans(buf.v[2 * kt].y);
          const uint32_t a2 = movmatrix_trans(buf.v[2 * kt + 1].x), a3 = movmatrix_trans(buf.v[2 * kt + 1].y);
#pragma unroll
          for (int nt = 0; nt < NT8; ++nt) {
            const float p0 = ex2_approx(sc[kt][nt][0] - m[nt][0]), p1 = ex2_approx(sc[kt][nt][1] - m[nt][1]);
            const float p2 = ex2_approx(sc[kt][nt][2] - m[nt][0]), p3 = ex2_approx(sc[kt][nt][3] - m[nt][1]);
            l[nt][0] += p0 + p2;
            l[nt][1] += p1 + p3;
            mma_bf16(o[nt], a0, a1, a2, a3, movmatrix_trans(pack_bf16x2(p0, p1)), movmatrix_trans(pack_bf16x2(p2, p3)));
          }
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
      // per-warp partial -> shared memory.  O^T rows: virtual channel g -> physical 4 (g >> 1) + (g & 1), g + 8 -> that + 2
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        float l0 = l[nt][0], l1 = l[nt][1];
#pragma unroll
        for (int sh = 4; sh < 32; sh <<= 1) {
          l0 += __shfl_xor_sync(0xffffffffu, l0, sh);
          l1 += __shfl_xor_sync(0xffffffffu, l1, sh);
        }
        const int ch = 4 * (g >> 1) + (g & 1), tok = nt * 8 + 2 * t;
        s.tpart[warp][tok][ch] = o[nt][0];
        s.tpart[warp][tok + 1][ch] = o[nt][1];
        s.tpart[warp][tok][ch + 2] = o[nt][2];
        s.tpart[warp][tok + 1][ch + 2] = o[nt][3];
        if (g == 0) {
          s.tpart[warp][tok][16] = m[nt][0]; s.tpart[warp][tok][17] = l0;
          s.tpart[warp][tok + 1][16] = m[nt][1]; s.tpart[warp][tok + 1][17] = l1;
        }
      }
    }
    __syncthreads();
    DT_TRACE(9);
    if (tid < 256) {   // merge the 16 warps' partials; thread = (row, channel)
      const int row = tid >> 4, ch = tid & 15;
      if (row < Nt) {
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
        s.tqs[row][ch] = num / den;
      }
    }
    if (flags & DEC_TOK_MLP) load_w<256>(wt, p.m1_w + (long long)(256 * (int)rank + 16 * warp) * 256, 256, g, t);   // mlp1 tile of this warp
    __syncthreads();
    wait_cluster_start();
    if (tid < 2 * Nt) {
      const int row = tid >> 1, c8 = (tid & 1) * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = s.tqs[row][c8 + j];
      bcast_op8(s.opC, row, 16 * (int)rank + c8, v);
    }
    DT_TRACE(10);
    cluster_sync_all();
    DT_TRACE(11);
    if (warp < 8) ks_mma(s.opC, warp & 3);   // output projection (K = 128): 2 tiles x 4 k-slices
    __syncthreads();
    ks_bcast_rows(4, P_TO, true);
    DT_TRACE(12);
    cluster_sync_all();
    DT_TRACE(13);
    layer_norm(P_N2W, P_N2B, (flags & DEC_TOK_MLP) != 0, !(flags & DEC_TOK_MLP));
    __syncthreads();
    DT_TRACE(14);
  }

  // ================================================================== token MLP (:200-203) + i2t k / v projections (:205-208)
  if (flags & DEC_TOK_MLP) {
    {   // hidden units [256 r + 16 w, + 16)
      const float bz0 = s.prm[P_M1 + 16 * warp + g], bz1 = s.prm[P_M1 + 16 * warp + g + 8];
      float o[NT8][4];
      mma_w<256, NT8, true>(wt, s.opB, 0, o, g, t);
      // second GEMM over this CTA's hidden slice: output columns [16 w, 16 w + 16), requested while the hidden units are stored
      load_w<256>(wt, p.m2_w + (long long)(16 * warp) * 2048 + 256 * (int)rank, 2048, g, t);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = g + 8 * (e >> 1), tok = nt * 8 + 2 * t + (e & 1);
          s.opH[0][tok][16 * warp + col] = __float2bfloat16_rn(tok < Nt ? fmaxf(o[nt][e] + ((e >> 1) ? bz1 : bz0), 0.f) : 0.f);
        }
    }
    __syncthreads();
    DT_TRACE(15);
    {   // partial sums of all 256 output columns (hidden activations bf16: no lo half) -> local staging -> owner CTAs
      float o[NT8][4];
      mma_w<256, NT8, false>(wt, s.opH, 0, o, g, t);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s.ybuf[nt * 8 + 2 * t + (e & 1)][16 * warp + g + 8 * (e >> 1)] = o[nt][e];
    }
    // image->token attention k (x + pe) and v (x) columns [16 r, 16 r + 16): 2 tiles x 8 k-slices
    load_w<32>(ws, (warp < 8 ? p.ik_w : p.iv_w) + (long long)(16 * rank) * 256 + 32 * (warp & 7), 256, g, t);
    __syncthreads();
#pragma unroll
    for (int i = tid; i < 8 * Nt * 8; i += DT_THREADS) {   // (token, owner, 4 columns): 16 bytes each to the owner's `part`
      const int cg = (i & 7) * 4, owner = (i >> 3) & 7, tok = i >> 6;
      const uint4 v = *reinterpret_cast<const uint4*>(&s.ybuf[tok][32 * owner + cg]);
      st_cluster_v4(mapa_u32(smem_u32(&s.part[rank][tok][cg]), (uint32_t)owner), v);
    }
    DT_TRACE(16);
    cluster_sync_all();
    DT_TRACE(17);
    if (tid < 8 * Nt) {   // reduce this CTA's 32 columns in fixed source order, + bias + residual, all-gather
      const int tok = tid >> 3, cg = (tid & 7) * 4, col = 32 * (int)rank + cg;
      float4 v = *reinterpret_cast<const float4*>(&s.part[0][tok][cg]);
#pragma unroll
      for (int r = 1; r < CLD; ++r) {
        const float4 u = *reinterpret_cast<const float4*>(&s.part[r][tok][cg]);
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
      }
      const float4 bz = *reinterpret_cast<const float4*>(&s.prm[P_M2 + cg]);
      const float4 xr = *reinterpret_cast<const float4*>(&s.xs[tok][col]);
      v.x += bz.x + xr.x; v.y += bz.y + xr.y; v.z += bz.z + xr.z; v.w += bz.w + xr.w;
      bcast_v4(&s.ybuf[tok][col], make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
    }
    DT_TRACE(18);
    cluster_sync_all();
    DT_TRACE(19);
    layer_norm(P_N3W, P_N3B, true, true);
    __syncthreads();
    ks_mma(warp < 8 ? s.opA : s.opB, warp & 7);
    __syncthreads();
    if (tid < 128 && e_tok < Nt) {
      const float4 v = ks_sum4(e_tile, 8);
      const float4 bz = *reinterpret_cast<const float4*>(&s.prm[(e_tile == 0 ? P_IK : P_IV) + e_cg]);
      float* dst = (e_tile == 0 ? p.kt : p.vt) + (long long)b * Nt * 128 + 16 * rank;
      *reinterpret_cast<float4*>(dst + e_tok * 128 + e_cg) = make_float4(v.x + bz.x, v.y + bz.y, v.z + bz.z, v.w + bz.w);
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
