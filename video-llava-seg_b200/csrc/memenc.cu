// Memory-encoder bandwidth kernels (memory_encoder.py:17-117): the mask down-sampler's strided 3x3
// convolutions fused with LayerNorm2d + GELU (and, for the first one, with the x4 bilinear up-sampling
// and sigmoid*20-10 of sam2_base.py:372-378,703-708, so the [B,1,1024,1024] f32 mask never has to be
// materialised), the im2col gather that turns the last 3x3/s2 convolution into a tcgen05 GEMM, and the
// CXBlock depth-wise 7x7 convolution fused with LayerNorm2d.  Activations are NHWC ("token rows").
#include "common.cuh"
#include "kernels.h"

namespace vls {

int g_dwconv_tma = 1;   // vls_set_tuning("dwconv_tma")
int g_dwconv_small = 1; // vls_set_tuning("dwconv_small"): 1 = small batches take the one-CTA-per-8-pixels kernel (default), 0 = the strip kernel

namespace {

// bilinear sample of a low-res map at high-res pixel (Y, X), align_corners=False, integer factor F
__device__ __forceinline__ float bilerp(const float* __restrict__ lo, int h, int w, int Y, int X, float inv_f) {
  const float sy = fmaxf((Y + 0.5f) * inv_f - 0.5f, 0.f), sx = fmaxf((X + 0.5f) * inv_f - 0.5f, 0.f);
  const int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly = sy - y0, lx = sx - x0;
  const float a = lo[y0 * w + x0], b = lo[y0 * w + x1], c = lo[y1 * w + x0], d = lo[y1 * w + x1];
  return (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * c + lx * d);
}

// ------------------------------------------------------------------ stage 1: 1 -> 4 channels, HxW -> H/2 x W/2
// mode 0: src is the high-res mask, used as is        (MemoryEncoder.forward skip_mask_sigmoid=True)
// mode 1: src is the high-res mask, sigmoid(src)*scale + bias   (skip_mask_sigmoid=False: scale 1, bias 0)
// mode 4: src is the high-res mask, (src > 0)*scale + bias
// mode 2: src is the LOW-res logit map [h/F], value = sigmoid(bilinear)*scale + bias   (sam2_base.py:703-708)
// mode 3: src is the LOW-res logit map, value = (bilinear > 0)*scale + bias            (sam2_base.py:698-700)
__global__ void __launch_bounds__(256)
mds1_kernel(const float* __restrict__ src, int mode, int H, int W, int factor, float scale, float bias_v,
            const float* __restrict__ wgt /*[4][9]*/, const float* __restrict__ cb, const float* __restrict__ lnw,
            const float* __restrict__ lnb, float eps, bf16* __restrict__ out) {
  pdl_enter();
  __shared__ float tile[33][34];
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * 16, ox0 = blockIdx.x * 16;
  const int OH = H >> 1, OW = W >> 1;
  const int lh = H / factor, lw = W / factor;
  const bool low = (mode == 2 || mode == 3);
  const float* s = src + (long long)b * (low ? (long long)lh * lw : (long long)H * W);
  const float inv_f = 1.0f / factor;
#pragma unroll 5   // 4.25 elements per thread: their sixteen low-res loads go out together
  for (int i = threadIdx.x; i < 33 * 33; i += 256) {
    const int ty = i / 33, tx = i % 33;
    const int Y = 2 * oy0 - 1 + ty, X = 2 * ox0 - 1 + tx;
    float v = 0.f;
    if (Y >= 0 && Y < H && X >= 0 && X < W) {
      if (mode == 0) v = s[(long long)Y * W + X];
      else if (mode == 1) v = 1.f / (1.f + expf(-s[(long long)Y * W + X])) * scale + bias_v;
      else if (mode == 4) v = (s[(long long)Y * W + X] > 0.f ? 1.f : 0.f) * scale + bias_v;
      else {
        const float hr = bilerp(s, lh, lw, Y, X, inv_f);
        v = (mode == 2 ? 1.f / (1.f + expf(-hr)) : (hr > 0.f ? 1.f : 0.f)) * scale + bias_v;
      }
    }
    tile[ty][tx] = v;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy >= OH || ox >= OW) return;
  float acc[4];
#pragma unroll
  for (int co = 0; co < 4; ++co) acc[co] = cb[co];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float v = tile[2 * ty + ky][2 * tx + kx];
#pragma unroll
      for (int co = 0; co < 4; ++co) acc[co] += v * wgt[co * 9 + ky * 3 + kx];
    }
  const float mean = 0.25f * (acc[0] + acc[1] + acc[2] + acc[3]);
  float var = 0.f;
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    acc[co] -= mean;
    var += acc[co] * acc[co];
  }
  const float rstd = rsqrtf(var * 0.25f + eps);
#pragma unroll
  for (int co = 0; co < 4; ++co) acc[co] = gelu_erf(acc[co] * rstd * lnw[co] + lnb[co]);
  *reinterpret_cast<uint2*>(out + (((long long)b * OH + oy) * OW + ox) * 4) =
      make_uint2(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]));
}

// ------------------------------------------------------------------ stage 2: 4 -> 16 channels (NHWC bf16)
// wgt f32 [9 taps][4 ci][16 co]
__global__ void __launch_bounds__(256)
mds2_kernel(const bf16* __restrict__ in, int H, int W, const float* __restrict__ wgt, const float* __restrict__ cb,
            const float* __restrict__ lnw, const float* __restrict__ lnb, float eps, bf16* __restrict__ out) {
  // A pair of lanes owns one output pixel, 8 of its 16 channels each (LayerNorm statistics through one shuffle): at one
  // thread per pixel the 256 x 256 outputs of the production shape were 256 CTAs = 8-16 warps per SM, each a long chain
  // (9 loads, 576 FMAs, 16 GELUs): 15 us in the frame for 2.9 M warp instructions.
  pdl_enter();
  __shared__ __align__(16) float sw[9 * 4 * 16];
  for (int i = threadIdx.x; i < 9 * 4 * 16; i += 256) sw[i] = wgt[i];
  __syncthreads();
  const int b = blockIdx.z, OH = H >> 1, OW = W >> 1;
  const int lane = threadIdx.x & 31, half = lane & 1;
  const int ox = blockIdx.x * 16 + (lane >> 1), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  const bool ok = ox < OW && oy < OH;       // (no early return: the partner lane takes part in the shuffles)
  float acc[8];
#pragma unroll
  for (int co = 0; co < 8; ++co) acc[co] = cb[8 * half + co];
  const bf16* base = in + (long long)b * H * W * 4;
  // all nine taps are requested before the first multiply (zero outside the image): with a `continue` per tap the loads
  // were nine dependent L2 round trips
  uint2 raw[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int Y = 2 * oy + ky - 1, X = 2 * ox + kx - 1;
      raw[ky * 3 + kx] = (ok && Y >= 0 && Y < H && X >= 0 && X < W) ? *reinterpret_cast<const uint2*>(base + ((long long)Y * W + X) * 4)
                                                                     : make_uint2(0u, 0u);
    }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&raw[t].x);
    const __nv_bfloat162 p1 = *reinterpret_cast<const __nv_bfloat162*>(&raw[t].y);
    const float v[4] = {__low2float(p0), __high2float(p0), __low2float(p1), __high2float(p1)};
    const float4* wp = reinterpret_cast<const float4*>(sw + t * 64) + 2 * half;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci)
#pragma unroll
      for (int c4 = 0; c4 < 2; ++c4) {
        const float4 w4 = wp[ci * 4 + c4];
        acc[4 * c4] += v[ci] * w4.x;
        acc[4 * c4 + 1] += v[ci] * w4.y;
        acc[4 * c4 + 2] += v[ci] * w4.z;
        acc[4 * c4 + 3] += v[ci] * w4.w;
      }
  }
  float mean = 0.f;
#pragma unroll
  for (int co = 0; co < 8; ++co) mean += acc[co];
  mean += __shfl_xor_sync(0xffffffffu, mean, 1);
  mean *= (1.f / 16.f);
  float var = 0.f;
#pragma unroll
  for (int co = 0; co < 8; ++co) {
    acc[co] -= mean;
    var += acc[co] * acc[co];
  }
  var += __shfl_xor_sync(0xffffffffu, var, 1);
  const float rstd = rsqrtf(var * (1.f / 16.f) + eps);
  if (!ok) return;
  uint32_t pk[4];
#pragma unroll
  for (int co = 0; co < 4; ++co)
    pk[co] = pack_bf16x2(gelu_erf(acc[2 * co] * rstd * lnw[8 * half + 2 * co] + lnb[8 * half + 2 * co]),
                         gelu_erf(acc[2 * co + 1] * rstd * lnw[8 * half + 2 * co + 1] + lnb[8 * half + 2 * co + 1]));
  *reinterpret_cast<uint4*>(out + (((long long)b * OH + oy) * OW + ox) * 16 + 8 * half) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// ------------------------------------------------------------------ stage 3: 16 -> 64 channels (NHWC bf16)
// 4 lanes per output-pixel PAIR, 16 output channels each; every weight vector fetched from shared memory feeds both
// pixels (the one-pixel version was bound by its 576 LDS.128 per thread).  wgt f32 [9 taps][16 ci][64 co] in smem.
// CTA = 16 x 8 output pixels.
__global__ void __launch_bounds__(256)
mds3_kernel(const bf16* __restrict__ in, int H, int W, const float* __restrict__ wgt, const float* __restrict__ cb,
            const float* __restrict__ lnw, const float* __restrict__ lnb, float eps, bf16* __restrict__ out) {
  pdl_enter();
  extern __shared__ float sw3[];
#pragma unroll 9
  for (int i = threadIdx.x; i < 9 * 16 * 64 / 4; i += 256)
    reinterpret_cast<float4*>(sw3)[i] = reinterpret_cast<const float4*>(wgt)[i];
  __syncthreads();
  const int b = blockIdx.z, OH = H >> 1, OW = W >> 1;
  const int q = threadIdx.x & 3;
  const int ox = blockIdx.x * 16 + 2 * ((threadIdx.x >> 2) & 7), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  const bool row_ok = oy < OH;
  float acc[2][16];
#pragma unroll
  for (int co = 0; co < 16; ++co) acc[0][co] = acc[1][co] = cb[q * 16 + co];
  const bf16* base = in + (long long)b * H * W * 16;
  if (row_ok && ox < OW) {
    // NOT unrolled over the taps: fully unrolled the kernel is ~9 k straight-line instructions per warp (150 KB of code)
    // and ncu shows it stalled on instruction fetch (stall_no_instruction 5.7 per issue); one tap is ~1 k
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap - 3 * ky;
      const int Y = 2 * oy + ky - 1;
      if (Y < 0 || Y >= H) continue;
      uint4 rr[2][2];                      // the two input columns (one per pixel of the pair) this tap reads
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int X = 2 * (ox + p) + kx - 1;
        const bool inb = X >= 0 && X < W;
        const uint4* ip = reinterpret_cast<const uint4*>(base + ((long long)Y * W + (inb ? X : 0)) * 16);
        rr[p][0] = inb ? ip[0] : make_uint4(0, 0, 0, 0);
        rr[p][1] = inb ? ip[1] : make_uint4(0, 0, 0, 0);
      }
      const float* wp = sw3 + tap * 16 * 64 + q * 16;
#pragma unroll
      for (int c2 = 0; c2 < 8; ++c2) {
        const float4* w0 = reinterpret_cast<const float4*>(wp + (2 * c2) * 64);
        const float4* w1 = reinterpret_cast<const float4*>(wp + (2 * c2 + 1) * 64);
        float4 wa[4], wd[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          wa[c4] = w0[c4];
          wd[c4] = w1[c4];
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const uint4 lo = rr[p][0], hi = rr[p][1];
          const uint32_t ru[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          const __nv_bfloat162 pv = *reinterpret_cast<const __nv_bfloat162*>(&ru[c2]);
          const float v0 = __low2float(pv), v1 = __high2float(pv);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            acc[p][4 * c4] += v0 * wa[c4].x + v1 * wd[c4].x;
            acc[p][4 * c4 + 1] += v0 * wa[c4].y + v1 * wd[c4].y;
            acc[p][4 * c4 + 2] += v0 * wa[c4].z + v1 * wd[c4].z;
            acc[p][4 * c4 + 3] += v0 * wa[c4].w + v1 * wd[c4].w;
          }
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float s = 0.f;
#pragma unroll
    for (int co = 0; co < 16; ++co) s += acc[p][co];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const float mean = s * (1.f / 64.f);
    float var = 0.f;
#pragma unroll
    for (int co = 0; co < 16; ++co) {
      acc[p][co] -= mean;
      var += acc[p][co] * acc[p][co];
    }
    var += __shfl_xor_sync(0xffffffffu, var, 1);
    var += __shfl_xor_sync(0xffffffffu, var, 2);
    const float rstd = rsqrtf(var * (1.f / 64.f) + eps);
    if (!row_ok || ox + p >= OW) continue;
    uint32_t pk[8];
#pragma unroll
    for (int co = 0; co < 8; ++co) {
      const int c = q * 16 + 2 * co;
      pk[co] = pack_bf16x2(gelu_erf(acc[p][2 * co] * rstd * lnw[c] + lnb[c]),
                           gelu_erf(acc[p][2 * co + 1] * rstd * lnw[c + 1] + lnb[c + 1]));
    }
    uint4* o = reinterpret_cast<uint4*>(out + (((long long)b * OH + oy) * OW + ox + p) * 64 + q * 16);
    o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// ------------------------------------------------------------------ im2col for the 64 -> 256, 3x3/s2 convolution
// in NHWC bf16 [B][H*W][C]; out bf16 [B][(H/2)*(W/2)][9*C], column (ky*3+kx)*C + ci. One uint4 (8 ch) per thread.
__global__ void im2col3x3s2_kernel(const bf16* __restrict__ in, int B, int H, int W, int C, bf16* __restrict__ out) {
  pdl_enter();
  const int OH = H >> 1, OW = W >> 1, c8 = C >> 3;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * 9 * c8;
  if (id >= total) return;
  const int v = (int)(id % c8);
  const int tap = (int)((id / c8) % 9);
  const long long pix = id / (9 * c8);
  const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), b = (int)(pix / ((long long)OW * OH));
  const int Y = 2 * oy + tap / 3 - 1, X = 2 * ox + tap % 3 - 1;
  uint4 val = make_uint4(0, 0, 0, 0);
  if (Y >= 0 && Y < H && X >= 0 && X < W)
    val = *reinterpret_cast<const uint4*>(in + (((long long)b * H + Y) * W + X) * C + v * 8);
  *reinterpret_cast<uint4*>(out + pix * 9 * C + tap * C + v * 8) = val;
}

// ------------------------------------------------------------------ CXBlock: depth-wise 7x7 (pad 3) + LayerNorm2d
// x f32 NHWC [B][H*W][256]; wgt f32 [49][256]; out bf16 [B][H*W][256].
// block = 256 threads (one per channel) computes 8 consecutive pixels of a row with a sliding window.
__global__ void __launch_bounds__(256)
dwconv7_ln_kernel(const float* __restrict__ x, int H, int W, const float* __restrict__ wgt, const float* __restrict__ cb,
                  const float* __restrict__ lnw, const float* __restrict__ lnb, float eps, bf16* __restrict__ out) {
  pdl_enter();
  __shared__ float red[8][2][8];
  const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
  const int b = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * 8;
  const float* xb = x + (long long)b * H * W * 256;
  float acc[8];
  const float bias = cb[c];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = bias;
  for (int dy = 0; dy < 7; ++dy) {
    const int Y = y + dy - 3;
    if (Y < 0 || Y >= H) continue;
    float row[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) {
      const int X = x0 + i - 3;
      row[i] = (X >= 0 && X < W) ? xb[((long long)Y * W + X) * 256 + c] : 0.f;
    }
    float wv[7];
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) wv[dx] = wgt[(dy * 7 + dx) * 256 + c];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) acc[j] += wv[dx] * row[j + dx];
  }
  // LayerNorm over the 256 channels (= threads) of each of the 8 pixels
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s = warp_sum(acc[j]);
    if (lane == 0) red[warp][0][j] = s;
  }
  __syncthreads();
  float mean[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][0][j];
    mean[j] = s * (1.f / 256.f);
    acc[j] -= mean[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s = warp_sum(acc[j] * acc[j]);
    if (lane == 0) red[warp][1][j] = s;
  }
  __syncthreads();
  const float g = lnw[c], be = lnb[c];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][1][j];
    const float rstd = rsqrtf(s * (1.f / 256.f) + eps);
    const int X = x0 + j;
    if (X < W) out[(((long long)b * H + y) * W + X) * 256 + c] = __float2bfloat16_rn(acc[j] * rstd * g + be);
  }
}

}  // namespace

int launch_mds1(const float* src, int mode, int B, int H, int W, int factor, float scale, float bias_v, const float* wgt,
                const float* cb, const float* lnw, const float* lnb, float eps, void* out, cudaStream_t stream) {
  VLS_REQUIRE(H % 2 == 0 && W % 2 == 0 && factor >= 1 && H % factor == 0 && W % factor == 0, "mds1: bad shape");
  VLS_CUDA(launch_k(mds1_kernel, dim3(dim3((W / 2 + 15) / 16, (H / 2 + 15) / 16, B)), dim3(256), 0, stream, src, mode, H, W, factor, scale, bias_v, wgt, cb, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_mds2(const void* in, int B, int H, int W, const float* wgt, const float* cb, const float* lnw, const float* lnb,
                float eps, void* out, cudaStream_t stream) {
  VLS_CUDA(launch_k(mds2_kernel, dim3(dim3((W / 2 + 15) / 16, (H / 2 + 7) / 8, B)), dim3(256), 0, stream, reinterpret_cast<const bf16*>(in), H, W, wgt, cb, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_mds3(const void* in, int B, int H, int W, const float* wgt, const float* cb, const float* lnw, const float* lnb,
                float eps, void* out, cudaStream_t stream) {
  const int smem = 9 * 16 * 64 * 4;
  VLS_CUDA(launch_k(mds3_kernel, dim3(dim3((W / 2 + 15) / 16, (H / 2 + 7) / 8, B)), dim3(256), smem, stream, reinterpret_cast<const bf16*>(in), H, W, wgt, cb, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_im2col3x3s2(const void* in, int B, int H, int W, int C, void* out, cudaStream_t stream) {
  VLS_REQUIRE(C % 8 == 0, "im2col: C must be a multiple of 8");
  const long long total = (long long)B * (H / 2) * (W / 2) * 9 * (C / 8);
  VLS_CUDA(launch_k(im2col3x3s2_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, reinterpret_cast<const bf16*>(in), B, H, W, C, reinterpret_cast<bf16*>(out)));
  VLS_POST_LAUNCH(1);
  return 0;
}

// ---- v2: column-strip sliding window with packed FP32 FMAs.
// The op is 49 FMA per 6 bytes: above the FP32 machine balance of plain FFMA, so the kernel is written for the FMA
// pipe first: a thread owns a channel PAIR and every multiply-add is one fma.rn.f32x2 (FFMA2: two FMAs per issue slot).
// CTA = 128 threads = 256 channels; it walks a strip of 4 pixel columns down R output rows.  Each input row is loaded
// ONCE (10 x LDG.64 per thread, 256 B per warp and pixel) and scattered into the seven output rows it contributes to;
// the partial output rows live in registers and rotate statically (row loop unrolled by 8, two input rows per step so
// that one LDS.64 of a depth-wise weight feeds 8 FFMA2).  The LayerNorm statistics of two finished rows (2 x 4 pixels x
// {sum, sum of squares}) are reduced with a 16-shuffle butterfly transpose instead of 80 plain shuffles, one
// __syncthreads per two output rows.  Depth-wise weights sit in shared memory as [tap][channel pair].
constexpr int DW_TX = 4;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

__global__ void __launch_bounds__(128)   // 164 registers -> 3 CTAs/SM; capping at 128 (4 CTAs) spills and is 13 % slower
dwconv7_ln_strip_kernel(const float* __restrict__ x, int H, int W, int R, const float* __restrict__ wgt,
                        const float* __restrict__ cb, const float* __restrict__ lnw, const float* __restrict__ lnb, float eps,
                        bf16* __restrict__ out) {
  pdl_enter();
  extern __shared__ float2 dw_smem[];
  float2* sw = dw_smem;                                            // [49][128]
  float* red = reinterpret_cast<float*>(dw_smem + 49 * 128);       // [2 parity][4 warps][16]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, y0 = blockIdx.y * R, x0 = blockIdx.x * DW_TX;
  for (int i = tid; i < 49 * 128; i += 128) sw[i] = *reinterpret_cast<const float2*>(wgt + (i >> 7) * 256 + 2 * (i & 127));
  const float2 bias = *reinterpret_cast<const float2*>(cb + 2 * tid);
  const float2 g = *reinterpret_cast<const float2*>(lnw + 2 * tid), be = *reinterpret_cast<const float2*>(lnb + 2 * tid);
  const float* xb = x + (long long)b * H * W * 256 + 2 * tid;
  bf16* ob = out + (long long)b * H * W * 256 + 2 * tid;
  __syncthreads();
  // eight partial output rows in registers (slot = output row mod 8 relative to the strip); every step consumes TWO
  // input rows per weight fetch: one LDS.64 feeds 8 FFMA2
  float2 acc[8][DW_TX];
#pragma unroll
  for (int s = 0; s < 8; ++s)
#pragma unroll
    for (int j = 0; j < DW_TX; ++j) acc[s][j] = bias;
  const int y_last = min(y0 + R, H) - 1;          // last output row of this CTA
  uint32_t xmask = 0;                             // which of the strip's 10 input columns exist
#pragma unroll
  for (int i = 0; i < DW_TX + 6; ++i) xmask |= (x0 + i - 3 >= 0 && x0 + i - 3 < W) ? (1u << i) : 0u;
  int parity = 0;
  for (int base = y0 - 3; base <= y_last + 3; base += 8) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const int yi = base + k;                     // input rows yi, yi + 1
      if (yi > y_last + 3) break;
      float2 row[2][DW_TX + 6];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const bool yok = yi + r >= 0 && yi + r < H;
        // one address per input row; the ten column loads use immediate offsets and the column mask hoisted out of the loop
        const float* rp = xb + ((long long)(yi + r) * W + (x0 - 3)) * 256;
#pragma unroll
        for (int i = 0; i < DW_TX + 6; ++i)
          row[r][i] = (yok && ((xmask >> i) & 1u)) ? *reinterpret_cast<const float2*>(rp + i * 256) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int dy = 0; dy < 7; ++dy) {             // input row yi+r, tap row dy -> output row yi + r + 3 - dy
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
          const float2 w2 = sw[(dy * 7 + dx) * 128 + tid];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int slot = (k + r + 3 - dy + 8) % 8;
#pragma unroll
            for (int j = 0; j < DW_TX; ++j) acc[slot][j] = ffma2(w2, row[r][j + dx], acc[slot][j]);
          }
        }
      }
      // output rows yi - 3 and yi - 2 are complete: slots (k + 5) % 8 and (k + 6) % 8
      const int yo = yi - 3;
      const int sa = (k + 5) % 8, sb = (k + 6) % 8;
      if (yo + 1 >= y0 && yo <= y_last) {
        // LayerNorm over the 256 channels of 2 rows x 4 pixels: 16 quantities, butterfly transpose-reduce (16 shuffles)
        float q[16];
#pragma unroll
        for (int j = 0; j < DW_TX; ++j) {
          const float2 v = acc[sa][j], u = acc[sb][j];
          q[j] = v.x + v.y;
          q[4 + j] = v.x * v.x + v.y * v.y;
          q[8 + j] = u.x + u.y;
          q[12 + j] = u.x * u.x + u.y * u.y;
        }
#pragma unroll
        for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
          const bool hi = lane & bit;
#pragma unroll
          for (int i = 0; i < half; ++i) {
            const float send = hi ? q[i] : q[i + half], keep = hi ? q[i + half] : q[i];
            q[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
          }
        }
        q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
        // lane (pairs of lanes) now holds quantity index lane >> 1 summed over the warp
        if ((lane & 1) == 0) red[(parity * 4 + warp) * 16 + (lane >> 1)] = q[0];
        __syncthreads();
        const float4* rp = reinterpret_cast<const float4*>(red + parity * 64);
        float4 t[4] = {rp[0], rp[1], rp[2], rp[3]};
#pragma unroll
        for (int w = 1; w < 4; ++w)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 a = rp[4 * w + c];
            t[c].x += a.x; t[c].y += a.y; t[c].z += a.z; t[c].w += a.w;
          }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int yr = yo + r;
          if (yr < y0 || yr > y_last) continue;
          const float sum[4] = {t[2 * r].x, t[2 * r].y, t[2 * r].z, t[2 * r].w};
          const float sq[4] = {t[2 * r + 1].x, t[2 * r + 1].y, t[2 * r + 1].z, t[2 * r + 1].w};
#pragma unroll
          for (int j = 0; j < DW_TX; ++j) {
            const float mean = sum[j] * (1.f / 256.f);
            const float var = fmaxf(sq[j] * (1.f / 256.f) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + eps);
            const int X = x0 + j;
            if (X < W) {
              const float2 v = r == 0 ? acc[sa][j] : acc[sb][j];
              *reinterpret_cast<uint32_t*>(ob + ((long long)yr * W + X) * 256) =
                  pack_bf16x2((v.x - mean) * rstd * g.x + be.x, (v.y - mean) * rstd * g.y + be.y);
            }
          }
        }
        parity ^= 1;
      }
#pragma unroll
      for (int j = 0; j < DW_TX; ++j) acc[sa][j] = acc[sb][j] = bias;   // these slots start output rows yi + 5, yi + 6
    }
  }
}

// ---- v3: the strip kernel with its input rows staged by TMA.
// What the ncu capture of v2 showed (profiles/r1_dwconv_strip_ncu_full_summary.txt): 100 M of its 170 M warp instructions were
// not FFMA2 -- per two input rows 20 predicated LDG.64 with their zero fills, halo masks and 64-bit address arithmetic -- and the
// issue slots were only half busy because every k-step began with 20 loads whose values the first FFMA2 needs.  Here a 4-D
// tensor map over x [B][H][W][256] delivers each k-step's box (2 rows x 10 columns x 256 channels = 20 KB) into a two-stage
// shared-memory ring: the halo is zero-filled by the TMA unit (coordinates may be negative / beyond the image), the box
// of k-step s+2 is requested as soon as k-step s has been consumed, and a thread reads its channel pair with immediate-offset
// LDS.64.  The depth-wise weights (50 KB) are read through L1 (ld.global.nc, 256 B per warp instruction): one copy per SM
// instead of one per CTA, which leaves 40 KB of shared memory per CTA.  LayerNorm epilogue in packed f32x2 arithmetic.
constexpr int DW_NST = 2;
constexpr int DW_STAGE_FLOATS = 2 * (DW_TX + 6) * 256;
constexpr size_t DW_TMA_SMEM = (size_t)DW_NST * DW_STAGE_FLOATS * 4 + 2 * 4 * 16 * 4 + DW_NST * 8;

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
dwconv7_ln_tma_kernel(const __grid_constant__ CUtensorMap tmX, int H, int W, int R, const float* __restrict__ wgt,
                      const float* __restrict__ cb, const float* __restrict__ lnw, const float* __restrict__ lnb, float eps,
                      bf16* __restrict__ out) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char dwt_smem[];
  float* ring = reinterpret_cast<float*>(dwt_smem);                              // [DW_NST][2 rows][10 columns][256]
  float* red = ring + DW_NST * DW_STAGE_FLOATS;                                  // [2 parity][4 warps][16]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * 4 * 16);                // [DW_NST]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, y0 = blockIdx.y * R, x0 = blockIdx.x * DW_TX;
  const int y_last = min(y0 + R, H) - 1;                   // last output row of this CTA
  const int nsteps = (y_last - y0 + 8) >> 1;               // k-steps: input rows y0 - 3 + 2s and the next one, up to y_last + 3
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < DW_NST; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
#pragma unroll
    for (int s = 0; s < DW_NST; ++s)
      if (s < nsteps) {
        mbar_expect_tx(&full[s], DW_STAGE_FLOATS * 4);
        tma_load_4d(ring + s * DW_STAGE_FLOATS, &tmX, &full[s], 0, x0 - 3, y0 - 3 + 2 * s, b);
      }
  }
  const float2 bias = *reinterpret_cast<const float2*>(cb + 2 * tid);
  const float2 g = *reinterpret_cast<const float2*>(lnw + 2 * tid), be = *reinterpret_cast<const float2*>(lnb + 2 * tid);
  const float2* w2p = reinterpret_cast<const float2*>(wgt) + tid;                // tap t: w2p[t * 128]
  bf16* ob = out + (long long)b * H * W * 256 + 2 * tid;
  __syncthreads();
  float2 acc[8][DW_TX];   // eight partial output rows (slot = output row mod 8 relative to the strip); slots are started by tap (0, 0)
#pragma unroll
  for (int s = 0; s < 8; ++s)
#pragma unroll
    for (int j = 0; j < DW_TX; ++j) acc[s][j] = bias;
  int parity = 0, step = 0;
  for (int base = y0 - 3; base <= y_last + 3; base += 8) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const int yi = base + k;                     // input rows yi, yi + 1
      if (yi > y_last + 3) break;
      const int stage = (k >> 1) % DW_NST;         // base advances by 4 k-steps, DW_NST divides 4
      mbar_wait(&full[stage], (uint32_t)(step / DW_NST) & 1u);
      const float2* rs = reinterpret_cast<const float2*>(ring + stage * DW_STAGE_FLOATS) + tid;
      float2 row[2][DW_TX + 6];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < DW_TX + 6; ++i) row[r][i] = rs[(r * (DW_TX + 6) + i) * 128];
#pragma unroll
      for (int dy = 0; dy < 7; ++dy) {             // input row yi+r, tap row dy -> output row yi + r + 3 - dy
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
          const float2 w2 = __ldg(w2p + (dy * 7 + dx) * 128);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int slot = (k + r + 3 - dy + 8) % 8;
#pragma unroll
            for (int j = 0; j < DW_TX; ++j)   // tap (0, 0) is the first contribution to its output row: it starts from the bias
              acc[slot][j] = ffma2(w2, row[r][j + dx], (dy == 0 && dx == 0) ? bias : acc[slot][j]);
          }
        }
      }
      // output rows yi - 3 and yi - 2 are complete: slots (k + 5) % 8 and (k + 6) % 8
      const int yo = yi - 3;
      const int sa = (k + 5) % 8, sb = (k + 6) % 8;
      const bool do_ln = yo + 1 >= y0 && yo <= y_last;
      if (do_ln) {
        // LayerNorm over the 256 channels of 2 rows x 4 pixels: 16 quantities, butterfly transpose-reduce (16 shuffles)
        float q[16];
#pragma unroll
        for (int j = 0; j < DW_TX; ++j) {
          const float2 v = acc[sa][j], u = acc[sb][j];
          const float2 vv = fmul2(v, v), uu = fmul2(u, u);
          q[j] = v.x + v.y;
          q[4 + j] = vv.x + vv.y;
          q[8 + j] = u.x + u.y;
          q[12 + j] = uu.x + uu.y;
        }
#pragma unroll
        for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
          const bool hi = lane & bit;
#pragma unroll
          for (int i = 0; i < half; ++i) {
            const float send = hi ? q[i] : q[i + half], keep = hi ? q[i + half] : q[i];
            q[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
          }
        }
        q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
        if ((lane & 1) == 0) red[(parity * 4 + warp) * 16 + (lane >> 1)] = q[0];
      }
      __syncthreads();                             // the stage is consumed by everyone; the partial sums are visible
      if (tid == 0 && step + DW_NST < nsteps) {
        mbar_expect_tx(&full[stage], DW_STAGE_FLOATS * 4);
        tma_load_4d(ring + stage * DW_STAGE_FLOATS, &tmX, &full[stage], 0, x0 - 3, y0 - 3 + 2 * (step + DW_NST), b);
      }
      if (do_ln) {
        const float4* rp = reinterpret_cast<const float4*>(red + parity * 64);
        float2 t[8];                               // t[2c], t[2c+1] = float4 number c of the 16 totals
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 a = rp[c];
          t[2 * c] = make_float2(a.x, a.y);
          t[2 * c + 1] = make_float2(a.z, a.w);
        }
#pragma unroll
        for (int w = 1; w < 4; ++w)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 a = rp[4 * w + c];
            t[2 * c] = fadd2(t[2 * c], make_float2(a.x, a.y));
            t[2 * c + 1] = fadd2(t[2 * c + 1], make_float2(a.z, a.w));
          }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int yr = yo + r;
          if (yr < y0 || yr > y_last) continue;
          const float sum[4] = {t[4 * r].x, t[4 * r].y, t[4 * r + 1].x, t[4 * r + 1].y};
          const float sq[4] = {t[4 * r + 2].x, t[4 * r + 2].y, t[4 * r + 3].x, t[4 * r + 3].y};
          bf16* orow = ob + ((long long)yr * W + x0) * 256;
#pragma unroll
          for (int j = 0; j < DW_TX; ++j) {
            const float mean = sum[j] * (1.f / 256.f);
            const float var = fmaxf(sq[j] * (1.f / 256.f) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + eps);
            if (x0 + j < W) {
              const float2 v = r == 0 ? acc[sa][j] : acc[sb][j];
              const float2 a2 = fmul2(g, make_float2(rstd, rstd));
              const float2 b2 = ffma2(make_float2(-mean, -mean), a2, be);
              const float2 o = ffma2(v, a2, b2);
              *reinterpret_cast<uint32_t*>(orow + j * 256) = pack_bf16x2(o.x, o.y);
            }
          }
        }
        parity ^= 1;
      }
      ++step;
    }
  }
}

int launch_dwconv7_ln(const float* x, int B, int H, int W, const float* wgt, const float* cb, const float* lnw,
                      const float* lnb, float eps, void* out, cudaStream_t stream) {
  // Column-strip kernels (FFMA2): a strip of 4 pixel columns is cut into n pieces of R = ceil(H / n) rows; every piece
  // recomputes 6 halo rows and the grid runs in whole waves of 148 SMs x 3 CTAs, so n minimises
  // (R + 6) / R x (waves rounded up / waves): 2 pieces of 32 rows at B = 64, 16 pieces of 4 rows (256 CTAs) at B = 1.
  // Few images (the production frame): one CTA per 8 pixels of a row (512 CTAs at B = 1); the strip kernel at B = 1
  // (vls_set_tuning("dwconv_small", 0)) was measured at the same frame time (0.9437 vs 0.9431 ms), so the default stays.
  const long long strips = (long long)((W + DW_TX - 1) / DW_TX) * B;
  const int tma = g_dwconv_tma;
  const bool tma_ok = tma && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (strips * ((H + 15) / 16) < 1184 && (g_dwconv_small || !tma_ok)) {
    VLS_CUDA(launch_k(dwconv7_ln_kernel, dim3((W + 7) / 8, H, B), dim3(256), 0, stream, x, H, W, wgt, cb, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
    VLS_POST_LAUNCH(1);
    return 0;
  }
  int R = H;
  {
    double best = 1e30;
    for (int n = 1; n <= 16; ++n) {
      const int r = (H + n - 1) / n;
      if (r < 4) break;
      const double waves = (double)strips * ((H + r - 1) / r) / (148.0 * 3.0);
      const double cost = (r + 6.0) / r * ceil(waves) / waves;
      if (cost < best - 1e-9) best = cost, R = r;
    }
  }
  if (tma_ok) {
    CUtensorMap tm;
    VLS_TRY(make_tmap_f32_nhwc(&tm, x, 256, W, H, B, DW_TX + 6, 2));
    static unsigned long long attr = 0;
    if (first_use_on_device(&attr)) {
      VLS_CUDA(cudaFuncSetAttribute(dwconv7_ln_tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_TMA_SMEM));
      VLS_CUDA(cudaFuncSetAttribute(dwconv7_ln_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_TMA_SMEM));
    }
    auto kern = tma == 2 ? dwconv7_ln_tma_kernel<4> : dwconv7_ln_tma_kernel<3>;
    VLS_CUDA(launch_k(kern, dim3((W + DW_TX - 1) / DW_TX, (H + R - 1) / R, B), dim3(128), DW_TMA_SMEM, stream, tm, H, W, R, wgt, cb, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
  } else {
    const size_t smem = 49 * 128 * sizeof(float2) + 2 * 4 * 16 * sizeof(float);
    static unsigned long long attr = 0;
    if (first_use_on_device(&attr)) {
      VLS_CUDA(cudaFuncSetAttribute(dwconv7_ln_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    VLS_CUDA(launch_k(dwconv7_ln_strip_kernel, dim3((W + DW_TX - 1) / DW_TX, (H + R - 1) / R, B), dim3(128), smem, stream, x, H, W, R, wgt, cb, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
  }
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
