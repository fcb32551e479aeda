// Mask-decoder kernels that are not plain GEMMs (sam/transformer.py:181-286, sam/mask_decoder.py:168-245,
// sam2_base.py:359-390): token self-attention, token->image and image->token attention with 8 heads of
// 32 / 16 channels, the two ConvTranspose2d(k2,s2) upscaling epilogues (the first one is a tcgen05
// GEMM + this LayerNorm2d/GELU pass, the second is fused with the hyper-network mask product so the
// [32,256,256] upscaled embedding never touches HBM), and the best-IoU / object-gate selection.
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

// ------------------------------------------------------------------ token self attention (8 heads x 32)
// q,k,v: f32 [B][Nt][256]; out f32 [B][Nt][256].  block = (head, b), one warp per query row.
__global__ void tok_self_attn_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                     const float* __restrict__ v, int Nt, float* __restrict__ out) {
  pdl_enter();
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  for (int i = warp; i < Nt; i += nw) {
    const float* qi = q + ((long long)b * Nt + i) * 256 + h * 32;
    float s = -INFINITY;
    if (lane < Nt) {
      const float* kj = k + ((long long)b * Nt + lane) * 256 + h * 32;
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) d += qi[c] * kj[c];
      s = d * scale;
    }
    const float m = warp_max(s);
    const float p = lane < Nt ? __expf(s - m) : 0.f;
    const float l = warp_sum(p);
    float acc = 0.f;
    for (int j = 0; j < Nt; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      acc += pj * v[((long long)b * Nt + j) * 256 + h * 32 + lane];
    }
    out[((long long)b * Nt + i) * 256 + h * 32 + lane] = acc / l;
  }
}

// ------------------------------------------------------------------ token -> image attention (8 heads x 16)
// q: f32 [B][Nt][128]; K, V: bf16 rows [B][T][ld] at column offsets koff/voff (+ head*16).
// block = (token, head, b), 256 threads stride over the T image tokens.
__global__ void __launch_bounds__(256)
t2i_attn_kernel(const float* __restrict__ q, const bf16* __restrict__ kv, long long ld, long long kv_sb, int koff,
                int voff, int Nt, int T, float* __restrict__ out) {
  pdl_enter();
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float red[8][18];
  float qv[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) qv[c] = q[((long long)b * Nt + i) * 128 + h * 16 + c] * 0.25f;  // 1/sqrt(16)
  const bf16* base = kv + (long long)b * kv_sb;
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  // four keys per step: their eight 16-byte K / V loads are issued back to back before any arithmetic (the rows are
  // `ld` elements apart, so every load is an L2 round trip; serialised they were most of the kernel's 22 us)
  for (int t0 = threadIdx.x; t0 < T; t0 += 4 * 256) {
    uint4 kq[4][2], vq[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * 256;
      if (t < T) {
        const uint4* kp = reinterpret_cast<const uint4*>(base + (long long)t * ld + koff + h * 16);
        const uint4* vp = reinterpret_cast<const uint4*>(base + (long long)t * ld + voff + h * 16);
        kq[u][0] = kp[0]; kq[u][1] = kp[1];
        vq[u][0] = vp[0]; vq[u][1] = vp[1];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t0 + u * 256 >= T) break;
      const uint32_t ku[8] = {kq[u][0].x, kq[u][0].y, kq[u][0].z, kq[u][0].w, kq[u][1].x, kq[u][1].y, kq[u][1].z, kq[u][1].w};
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const __nv_bfloat162 k2 = *reinterpret_cast<const __nv_bfloat162*>(&ku[c]);
        s += qv[2 * c] * __low2float(k2) + qv[2 * c + 1] * __high2float(k2);
      }
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn), p = __expf(s - mn);
      const uint32_t vu[8] = {vq[u][0].x, vq[u][0].y, vq[u][0].z, vq[u][0].w, vq[u][1].x, vq[u][1].y, vq[u][1].z, vq[u][1].w};
      l = l * corr + p;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(&vu[c]);
        acc[2 * c] = acc[2 * c] * corr + p * __low2float(v2);
        acc[2 * c + 1] = acc[2 * c + 1] * corr + p * __high2float(v2);
      }
      m = mn;
    }
  }
  // block combine
  const float wm = warp_max(m);
  const float wc = (m == -INFINITY) ? 0.f : __expf(m - wm);
  l = warp_sum(l * wc);
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = warp_sum(acc[c] * wc);
  if (lane == 0) {
    red[warp][16] = wm;
    red[warp][17] = l;
#pragma unroll
    for (int c = 0; c < 16; ++c) red[warp][c] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float gm = -INFINITY;
    for (int w = 0; w < 8; ++w) gm = fmaxf(gm, red[w][16]);
    float gl = 0.f, ga = 0.f;
    for (int w = 0; w < 8; ++w) {
      const float f = (red[w][16] == -INFINITY) ? 0.f : __expf(red[w][16] - gm);
      gl += red[w][17] * f;
      ga += red[w][threadIdx.x] * f;
    }
    out[((long long)b * Nt + i) * 128 + h * 16 + threadIdx.x] = ga / gl;
  }
}

// ------------------------------------------------------------------ image -> token attention (8 heads x 16)
// q: bf16 rows [B][T][ld] at column qoff; k_tok, v_tok f32 [B][Nt][128]; out bf16 [B][T][128].
__global__ void __launch_bounds__(256)
i2t_attn_kernel(const bf16* __restrict__ qrows, long long ld, long long q_sb, int qoff, const float* __restrict__ ktok,
                const float* __restrict__ vtok, int Nt, int T, bf16* __restrict__ out, int planes) {
  pdl_enter();
  extern __shared__ float sm_i2t[];
  float* sk = sm_i2t;
  float* sv = sm_i2t + Nt * 128;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < Nt * 128; i += 256) {
    sk[i] = ktok[(long long)b * Nt * 128 + i];
    sv[i] = vtok[(long long)b * Nt * 128 + i];
  }
  __syncthreads();
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  if (id >= (long long)T * 8) return;
  // rows: q of (t, h) at t * ld + qoff + 16 h, thread = (t, h);  planes: head-major [qoff + h][T][16], thread = (h, t)
  const int t = planes ? (int)(id % T) : (int)(id >> 3), h = planes ? (int)(id / T) : (int)(id & 7);
  const uint4* qp = reinterpret_cast<const uint4*>(qrows + (long long)b * q_sb +
                                                   (planes ? ((long long)(qoff + h) * T + t) * 16 : (long long)t * ld + qoff + h * 16));
  const uint4 q0 = qp[0], q1 = qp[1];
  const uint32_t qu[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
  float qv[16];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(&qu[c]);
    qv[2 * c] = __low2float(q2) * 0.25f;
    qv[2 * c + 1] = __high2float(q2) * 0.25f;
  }
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  for (int j = 0; j < Nt; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) s += qv[c] * sk[j * 128 + h * 16 + c];
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn), p = __expf(s - mn);
    l = l * corr + p;
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = acc[c] * corr + p * sv[j * 128 + h * 16 + c];
    m = mn;
  }
  const float inv = 1.0f / l;
  uint4* op = reinterpret_cast<uint4*>(out + ((long long)b * T + t) * 128 + h * 16);
  op[0] = make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                     pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
  op[1] = make_uint4(pack_bf16x2(acc[8] * inv, acc[9] * inv), pack_bf16x2(acc[10] * inv, acc[11] * inv),
                     pack_bf16x2(acc[12] * inv, acc[13] * inv), pack_bf16x2(acc[14] * inv, acc[15] * inv));
}

// ------------------------------------------------------------------ upscaling stage 1 epilogue
// g: f32 [B][h*w][4*64] = ConvTranspose2d(256->64,k2,s2) as a GEMM, column (dy*2+dx)*64 + co, bias included.
// out[b][(2y+dy)*2w + 2x+dx][co] = GELU(LN2d_64(g + feat_s1[b or 0][co][2y+dy][2x+dx]))   (bf16 rows)
// block = 32 consecutive output pixels of one output row; feat_s1 tile staged through smem.
__global__ void __launch_bounds__(256)
up1_post_kernel(const float* __restrict__ g, const void* __restrict__ feat, int feat_bf16, long long feat_sb, int h, int w,
                const float* __restrict__ lnw, const float* __restrict__ lnb, float eps, bf16* __restrict__ out) {
  pdl_enter();
  __shared__ float tile[64][33];
  const int b = blockIdx.z, Y = blockIdx.y, X0 = blockIdx.x * 32;
  const int W2 = 2 * w, H2 = 2 * h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < 64; c += 8) {
    const int X = X0 + lane;
    float v = 0.f;
    if (X < W2) {
      const long long idx = (long long)b * feat_sb + ((long long)c * H2 + Y) * W2 + X;
      v = feat_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(feat)[idx]) : reinterpret_cast<const float*>(feat)[idx];
    }
    tile[c][lane] = v;
  }
  __syncthreads();
  const float w0 = lnw[lane], w1 = lnw[lane + 32], b0 = lnb[lane], b1 = lnb[lane + 32];
  for (int px = warp; px < 32; px += 8) {
    const int X = X0 + px;
    if (X >= W2) continue;
    const int y = Y >> 1, x = X >> 1, blk = ((Y & 1) << 1) | (X & 1);
    const float* src = g + (((long long)b * h + y) * w + x) * 256 + blk * 64;
    float v0 = src[lane] + tile[lane][px];
    float v1 = src[lane + 32] + tile[lane + 32][px];
    const float mean = warp_sum(v0 + v1) * (1.0f / 64.0f);
    v0 -= mean;
    v1 -= mean;
    const float rstd = rsqrtf(warp_sum(v0 * v0 + v1 * v1) * (1.0f / 64.0f) + eps);
    v0 = gelu_erf(v0 * rstd * w0 + b0);
    v1 = gelu_erf(v1 * rstd * w1 + b1);
    bf16* o = out + (((long long)b * H2 + Y) * W2 + X) * 64;
    o[lane] = __float2bfloat16_rn(v0);
    o[lane + 32] = __float2bfloat16_rn(v1);
  }
}

// ------------------------------------------------------------------ upscaling stage 2 + hyper-network product
// u: bf16 rows [B][h2*w2][64]; w2t: f32 [4 pos][64 ci][32 co]; bias f32 [32]; feat_s0 [B or 1][32][2h2][2w2];
// hyper f32 [B][M][32]; masks f32 [B][M][2h2][2w2] = sum_co hyper[m][co] * GELU(convT(u) + bias + feat_s0).
// block = 32 consecutive input pixels of one input row, 128 threads = (pos, x).
template <int M>
__global__ void __launch_bounds__(128)
up2_masks_kernel(const bf16* __restrict__ u, const float* __restrict__ w2t, const float* __restrict__ bias,
                 const void* __restrict__ feat, int feat_bf16, long long feat_sb, const float* __restrict__ hyper, int h2,
                 int w2, float* __restrict__ masks) {
  pdl_enter();
  extern __shared__ float sm_up2[];
  float* sw = sm_up2;                 // [4 pos][64 ci][32 co]
  float* su = sw + 4 * 32 * 64;       // [64 ci][33]
  float* sf = su + 64 * 33;           // [2 dy][32 co][65]
  float* sh = sf + 2 * 32 * 65;       // [M][32]
  float* sb = sh + M * 32;            // [32]
  const int b = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * 32;
  const int H4 = 2 * h2, W4 = 2 * w2;
  // all fills are 128-bit and unrolled (they used to be ~110 serialised scalar round trips to L2 per thread)
#pragma unroll 8
  for (int i = threadIdx.x; i < 4 * 32 * 64 / 4; i += 128) reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(w2t)[i];
  for (int i = threadIdx.x; i < M * 32; i += 128) sh[i] = hyper[(long long)b * M * 32 + i];
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
#pragma unroll
  for (int i = threadIdx.x; i < 32 * 8; i += 128) {            // 32 pixels x 8 vectors of 8 channels
    const int px = i >> 3, c8 = (i & 7) * 8;
    const int x = x0 + px;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (x < w2) v = *reinterpret_cast<const uint4*>(u + (((long long)b * h2 + y) * w2 + x) * 64 + c8);
    const uint32_t vu[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 p2 = *reinterpret_cast<const __nv_bfloat162*>(&vu[j]);
      su[(c8 + 2 * j) * 33 + px] = __low2float(p2);
      su[(c8 + 2 * j + 1) * 33 + px] = __high2float(p2);
    }
  }
  if (!feat_bf16 && (W4 & 3) == 0 && (feat_sb & 3) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0) {
#pragma unroll 8
    for (int i = threadIdx.x; i < 2 * 32 * 16; i += 128) {     // (dy, co) rows of 64 floats = 16 float4
      const int X = (i & 15) * 4, co = (i >> 4) & 31, dy = i >> 9;
      const int Xg = 2 * x0 + X;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (Xg < W4)
        v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(feat) + (long long)b * feat_sb +
                                             ((long long)co * H4 + 2 * y + dy) * W4 + Xg);
      float* d = sf + (dy * 32 + co) * 65 + X;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  } else {
    for (int i = threadIdx.x; i < 2 * 32 * 64; i += 128) {
      const int X = i & 63, co = (i >> 6) & 31, dy = i >> 11;
      const int Xg = 2 * x0 + X;
      float v = 0.f;
      if (Xg < W4) {
        const long long idx = (long long)b * feat_sb + ((long long)co * H4 + 2 * y + dy) * W4 + Xg;
        v = feat_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(feat)[idx]) : reinterpret_cast<const float*>(feat)[idx];
      }
      sf[(dy * 32 + co) * 65 + X] = v;
    }
  }
  __syncthreads();
  const int px = threadIdx.x & 31, pos = threadIdx.x >> 5;  // pos = dy*2+dx, warp-uniform
  const int dy = pos >> 1, dx = pos & 1;
  float z[32];
#pragma unroll
  for (int co = 0; co < 32; ++co) z[co] = 0.f;
  const float4* wp = reinterpret_cast<const float4*>(sw + pos * 64 * 32);
#pragma unroll 2
  for (int ci = 0; ci < 64; ++ci) {
    const float uv = su[ci * 33 + px];
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 wv = wp[ci * 8 + c4];  // warp-uniform address: one broadcast LDS.128 per 4 output channels
      z[4 * c4] += uv * wv.x;
      z[4 * c4 + 1] += uv * wv.y;
      z[4 * c4 + 2] += uv * wv.z;
      z[4 * c4 + 3] += uv * wv.w;
    }
  }
  float mk[M];
#pragma unroll
  for (int m = 0; m < M; ++m) mk[m] = 0.f;
#pragma unroll
  for (int co = 0; co < 32; ++co) {
    const float a = gelu_erf(z[co] + sb[co] + sf[(dy * 32 + co) * 65 + 2 * px + dx]);
#pragma unroll
    for (int m = 0; m < M; ++m) mk[m] += sh[m * 32 + co] * a;
  }
  const int x = x0 + px;
  if (x < w2) {
#pragma unroll
    for (int m = 0; m < M; ++m)
      masks[(((long long)b * M + m) * H4 + 2 * y + dy) * W4 + 2 * x + dx] = mk[m];
  }
}

// ------------------------------------------------------------------ upscaling stage 2 + hyper-network product, tensor cores
// Same result as up2_masks_kernel; the ConvTranspose2d(64 -> 32, k2, s2) is a tcgen05 GEMM: a CTA owns 128 consecutive
// input pixels (one row of the 128 x 128 stage-1 map), D[128 pixels][128 = (dy, dx, co)] = U[128][64] . W^T in tensor
// memory.  The weights stay f32-exact: W = Wh + Wl (two bf16 matrices), 4 + 4 MMAs of K = 16 into the same accumulator.
// Epilogue: thread = pixel reads its 4 x 32 accumulators, adds bias and the feat_s0 skip (two 8-byte loads per (dy, co),
// 256 contiguous bytes per warp), GELU, dots the 32 channels with the M hyper-network vectors and stores 2 x M x 8 bytes.
// The r1 kernel did the 134 MFMA of the convolution on the FP32 pipe (21.9 us = 12 TFLOP/s).
constexpr int UP2_THREADS = 288;   // warps 0-7: epilogue (thread = (pixel = TMEM lane, dy)), warp 8: TMA + MMA
constexpr int UP2_SMEM = 3 * 128 * 64 * 2 + 1024 + 256 + (4 * 32 + 32) * 4 + 64;

template <int M>
__global__ void __launch_bounds__(UP2_THREADS, 1)
up2_masks_tc_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias,
                    const float* __restrict__ feat, long long feat_sb, const float* __restrict__ hyper, int h2, int w2,
                    float* __restrict__ masks) {
  extern __shared__ uint8_t up2_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(up2_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sU = smem;                       // [128 pixels][64 ci] bf16, SW128
  uint8_t* sW = smem + 128 * 64 * 2;        // [2][128 (pos, co)][64 ci]
  float* sh = reinterpret_cast<float*>(smem + 3 * 128 * 64 * 2);   // [32 co][M] hyper (transposed), then [32] bias
  float* sb = sh + 32 * M;
  uint64_t* full = reinterpret_cast<uint64_t*>(sb + 32);
  uint64_t* done = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, m0 = blockIdx.x * 128;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmU);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_enter();
  if (warp == 8) {
    if (elect_one()) {
      mbar_expect_tx(full, 3 * 128 * 64 * 2);
      tma_load_3d(sU, &tmU, full, 0, m0, b);
      tma_load_3d(sW, &tmW, full, 0, 0, 0);
      tma_load_3d(sW + 128 * 64 * 2, &tmW, full, 0, 128, 0);
      mbar_wait(full, 0);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16(128, 128);
      const uint64_t ad = make_desc_sw128(smem_u32(sU));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t bd = make_desc_sw128(smem_u32(sW + h * 128 * 64 * 2));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem, ad + 2 * k, bd + 2 * k, idesc, (h | k) != 0 ? 1u : 0u);
      }
      umma_commit(done);
    }
  } else {
    for (int i = threadIdx.x; i < 32 * M; i += 256) sh[i] = hyper[(long long)b * M * 32 + (i % M) * 32 + i / M];   // sh[co][m]
    if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
    const int q = warp & 3, dy = warp >> 2;   // TMEM lane quarter, output row parity
    const int pix = m0 + q * 32 + lane;
    const bool ok = pix < h2 * w2;
    const int y = ok ? pix / w2 : 0, x = ok ? pix % w2 : 0;
    const int H4 = 2 * h2, W4 = 2 * w2;
    const float* fbase = feat + (long long)b * feat_sb + (long long)(2 * y + dy) * W4 + 2 * x;
    float2 f[32];   // requested before the accumulator is ready
#pragma unroll
    for (int co = 0; co < 32; ++co)
      f[co] = ok ? __ldg(reinterpret_cast<const float2*>(fbase + (long long)co * H4 * W4)) : make_float2(0.f, 0.f);
    asm volatile("bar.sync 1, 256;" ::: "memory");   // hyper / bias staged (epilogue warps only)
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t lane_off = uint32_t(q * 32) << 16;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_off + (dy * 2 + dx) * 32, r);
      tc_wait_ld();
      float mk[M];
#pragma unroll
      for (int m = 0; m < M; ++m) mk[m] = 0.f;
#pragma unroll
      for (int co = 0; co < 32; ++co) {
        const float a = gelu_fast(__uint_as_float(r[co]) + sb[co] + (dx ? f[co].y : f[co].x));
#pragma unroll
        for (int m = 0; m < M; ++m) mk[m] += sh[co * M + m] * a;
      }
      if (dx == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) f[m].x = mk[m];   // park the dx = 0 results: one 8-byte store per m
      } else if (ok) {
#pragma unroll
        for (int m = 0; m < M; ++m)
          *reinterpret_cast<float2*>(masks + (((long long)b * M + m) * H4 + 2 * y + dy) * W4 + 2 * x) = make_float2(f[m].x, mk[m]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------ best-IoU / object-gate selection
// sam2_base.py:359-390: gate = obj_logit > 0; multimask -> argmax over iou[:,1:]; low_res = gate ? mask : -1024
__global__ void select_best_kernel(const float* __restrict__ masks, const float* __restrict__ iou,
                                   const float* __restrict__ tokens, const float* __restrict__ obj_logits, int M,
                                   int multimask, int HW, float* __restrict__ low_res, float* __restrict__ tok_sel,
                                   int* __restrict__ best_idx, float* __restrict__ is_obj_out,
                                   float* __restrict__ occluded_out, float no_obj_score) {
  pdl_enter();
  const int b = blockIdx.y;
  int best = 0;
  if (multimask) {
    best = 1;
    float bv = iou[b * M + 1];
    for (int m = 2; m < M; ++m)
      if (iou[b * M + m] > bv) {
        bv = iou[b * M + m];
        best = m;
      }
  }
  const bool is_obj = obj_logits[b] > 0.f;
  const float* src = masks + ((long long)b * M + best) * HW;
  float* dst = low_res + (long long)b * HW;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < HW; i += gridDim.x * blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(src + i);
    if (!is_obj) v = make_float4(no_obj_score, no_obj_score, no_obj_score, no_obj_score);
    *reinterpret_cast<float4*>(dst + i) = v;
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < 256; c += blockDim.x) tok_sel[b * 256 + c] = tokens[((long long)b * M + best) * 256 + c];
    if (threadIdx.x == 0) {
      best_idx[b] = best;
      is_obj_out[b] = is_obj ? 1.f : 0.f;
      occluded_out[b] = is_obj ? 0.f : 1.f;
    }
  }
}

// obj_ptr = is_obj ? ptr : no_obj_ptr   (sam2_base.py:394-403 with fixed_no_obj_ptr, hard gate)
__global__ void gate_ptr_kernel(float* __restrict__ ptr, const float* __restrict__ is_obj,
                                const float* __restrict__ no_obj_ptr, int B) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 256) return;
  const float lam = is_obj[i / 256];
  ptr[i] = lam * ptr[i] + (1.f - lam) * no_obj_ptr[i % 256];
}

}  // namespace

int launch_tok_self_attn(const float* q, const float* k, const float* v, int B, int Nt, float* out, cudaStream_t stream) {
  VLS_REQUIRE(Nt >= 1 && Nt <= 32, "decoder: between 1 and 32 tokens are supported (got %d)", Nt);
  VLS_CUDA(launch_k(tok_self_attn_kernel, dim3(dim3(8, B)), dim3(128), 0, stream, q, k, v, Nt, out));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_t2i_attn(const float* q, const void* kv, long long ld, long long kv_sb, int koff, int voff, int B, int Nt,
                    int T, float* out, cudaStream_t stream) {
  VLS_CUDA(launch_k(t2i_attn_kernel, dim3(dim3(Nt, 8, B)), dim3(256), 0, stream, q, reinterpret_cast<const bf16*>(kv), ld, kv_sb, koff, voff, Nt, T, out));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_i2t_attn(const void* qrows, long long ld, long long q_sb, int qoff, const float* ktok, const float* vtok, int B,
                    int Nt, int T, void* out, cudaStream_t stream, int planes) {
  const size_t smem = (size_t)Nt * 128 * 2 * sizeof(float);
  VLS_CUDA(launch_k(i2t_attn_kernel, dim3(dim3((T * 8 + 255) / 256, B)), dim3(256), smem, stream, reinterpret_cast<const bf16*>(qrows), ld, q_sb, qoff, ktok, vtok, Nt, T, reinterpret_cast<bf16*>(out), planes));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_up1_post(const void* g, const void* feat, int feat_bf16, long long feat_sb, int B, int h, int w,
                    const float* lnw, const float* lnb, float eps, void* out, cudaStream_t stream) {
  VLS_CUDA(launch_k(up1_post_kernel, dim3(dim3((2 * w + 31) / 32, 2 * h, B)), dim3(256), 0, stream, reinterpret_cast<const float*>(g), feat, feat_bf16, feat_sb, h, w, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_up2_masks(const void* u, const float* w2t, const float* bias, const void* feat, int feat_bf16,
                     long long feat_sb, const float* hyper, int B, int M, int h2, int w2, float* masks,
                     cudaStream_t stream) {
  VLS_REQUIRE(M == 4, "decoder: num_mask_tokens must be 4");
  const size_t smem = (size_t)(4 * 32 * 64 + 64 * 33 + 2 * 32 * 65 + 4 * 32 + 32) * sizeof(float);
  static unsigned long long attr = 0;
  if (first_use_on_device(&attr)) {
    VLS_CUDA(cudaFuncSetAttribute(up2_masks_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  VLS_CUDA(launch_k(up2_masks_kernel<4>, dim3(dim3((w2 + 31) / 32, h2, B)), dim3(128), smem, stream, reinterpret_cast<const bf16*>(u), w2t, bias, feat, feat_bf16, feat_sb, hyper, h2, w2, masks));
  VLS_POST_LAUNCH(1);
  return 0;
}

int g_up2_tc = 1;   // vls_set_tuning("up2_tc")
// tensor-core path of the fused ConvT#2 + hyper product: f32 skip features, w2 a multiple of 32; wh = bf16 [2][128][64]
int launch_up2_masks_tc(const void* u, const void* wh, const float* bias, const float* feat, long long feat_sb,
                        const float* hyper, int B, int M, int h2, int w2, float* masks, cudaStream_t stream) {
  VLS_REQUIRE((M == 1 || M == 4) && w2 % 32 == 0, "up2_masks_tc: M must be 1 or 4 and the width a multiple of 32");
  CUtensorMap tmU, tmW;
  VLS_TRY(make_tmap_bf16(&tmU, u, 64, (uint64_t)h2 * w2, B, 64, (long long)h2 * w2 * 64, 128));
  VLS_TRY(make_tmap_bf16(&tmW, wh, 64, 256, 1, 64, 256 * 64, 128));
  static unsigned long long attr1 = 0, attr4 = 0;
  const dim3 grid((h2 * w2 + 127) / 128, B);
  if (M == 4) {
    if (first_use_on_device(&attr4)) VLS_CUDA(cudaFuncSetAttribute(up2_masks_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, UP2_SMEM));
    VLS_CUDA(launch_k(up2_masks_tc_kernel<4>, grid, dim3(UP2_THREADS), UP2_SMEM, stream, tmU, tmW, bias, feat, feat_sb, hyper, h2, w2, masks));
  } else {
    if (first_use_on_device(&attr1)) VLS_CUDA(cudaFuncSetAttribute(up2_masks_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UP2_SMEM));
    VLS_CUDA(launch_k(up2_masks_tc_kernel<1>, grid, dim3(UP2_THREADS), UP2_SMEM, stream, tmU, tmW, bias, feat, feat_sb, hyper, h2, w2, masks));
  }
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_select_best(const float* masks, const float* iou, const float* tokens, const float* obj_logits, int B, int M,
                       int multimask, int HW, float* low_res, float* tok_sel, int* best_idx, float* is_obj,
                       float* occluded, cudaStream_t stream) {
  VLS_REQUIRE(HW % 4 == 0, "select_best: H*W must be a multiple of 4");
  VLS_CUDA(launch_k(select_best_kernel, dim3(dim3(16, B)), dim3(256), 0, stream, masks, iou, tokens, obj_logits, M, multimask, HW, low_res, tok_sel, best_idx, is_obj, occluded, -1024.0f));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_gate_ptr(float* ptr, const float* is_obj, const float* no_obj_ptr, int B, cudaStream_t stream) {
  VLS_CUDA(launch_k(gate_ptr_kernel, dim3((B * 256 + 255) / 256), dim3(256), 0, stream, ptr, is_obj, no_obj_ptr, B));
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
