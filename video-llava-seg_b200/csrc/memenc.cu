// Memory-encoder bandwidth kernels (memory_encoder.py:17-117): the mask down-sampler's strided 3x3
// convolutions fused with LayerNorm2d + GELU (and, for the first one, with the x4 bilinear up-sampling
// and sigmoid*20-10 of sam2_base.py:372-378,703-708, so the [B,1,1024,1024] f32 mask never has to be
// materialised), the im2col gather that turns the last 3x3/s2 convolution into a tcgen05 GEMM, and the
// CXBlock depth-wise 7x7 convolution fused with LayerNorm2d.  Activations are NHWC ("token rows").
#include "common.cuh"
#include "kernels.h"

namespace vls {

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
  pdl_enter();
  __shared__ float sw[9 * 4 * 16];
  for (int i = threadIdx.x; i < 9 * 4 * 16; i += 256) sw[i] = wgt[i];
  __syncthreads();
  const int b = blockIdx.z, OH = H >> 1, OW = W >> 1;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= OW || oy >= OH) return;
  float acc[16];
#pragma unroll
  for (int co = 0; co < 16; ++co) acc[co] = cb[co];
  const bf16* base = in + (long long)b * H * W * 4;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int Y = 2 * oy + ky - 1;
    if (Y < 0 || Y >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int X = 2 * ox + kx - 1;
      if (X < 0 || X >= W) continue;
      const uint2 raw = *reinterpret_cast<const uint2*>(base + ((long long)Y * W + X) * 4);
      const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
      const __nv_bfloat162 p1 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
      const float v[4] = {__low2float(p0), __high2float(p0), __low2float(p1), __high2float(p1)};
      const float* wp = sw + (ky * 3 + kx) * 64;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
#pragma unroll
        for (int co = 0; co < 16; ++co) acc[co] += v[ci] * wp[ci * 16 + co];
    }
  }
  float mean = 0.f;
#pragma unroll
  for (int co = 0; co < 16; ++co) mean += acc[co];
  mean *= (1.f / 16.f);
  float var = 0.f;
#pragma unroll
  for (int co = 0; co < 16; ++co) {
    acc[co] -= mean;
    var += acc[co] * acc[co];
  }
  const float rstd = rsqrtf(var * (1.f / 16.f) + eps);
  uint32_t pk[8];
#pragma unroll
  for (int co = 0; co < 8; ++co)
    pk[co] = pack_bf16x2(gelu_erf(acc[2 * co] * rstd * lnw[2 * co] + lnb[2 * co]),
                         gelu_erf(acc[2 * co + 1] * rstd * lnw[2 * co + 1] + lnb[2 * co + 1]));
  uint4* o = reinterpret_cast<uint4*>(out + (((long long)b * OH + oy) * OW + ox) * 16);
  o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

// ------------------------------------------------------------------ stage 3: 16 -> 64 channels (NHWC bf16)
// 4 lanes per output pixel, 16 output channels each.  wgt f32 [9 taps][16 ci][64 co] in smem.
__global__ void __launch_bounds__(256)
mds3_kernel(const bf16* __restrict__ in, int H, int W, const float* __restrict__ wgt, const float* __restrict__ cb,
            const float* __restrict__ lnw, const float* __restrict__ lnb, float eps, bf16* __restrict__ out) {
  pdl_enter();
  extern __shared__ float sw3[];
  for (int i = threadIdx.x; i < 9 * 16 * 64; i += 256) sw3[i] = wgt[i];
  __syncthreads();
  const int b = blockIdx.z, OH = H >> 1, OW = W >> 1;
  const int q = threadIdx.x & 3;
  const int ox = blockIdx.x * 16 + ((threadIdx.x >> 2) & 15), oy = blockIdx.y * 4 + (threadIdx.x >> 6);
  const bool ok = ox < OW && oy < OH;
  float acc[16];
#pragma unroll
  for (int co = 0; co < 16; ++co) acc[co] = cb[q * 16 + co];
  const bf16* base = in + (long long)b * H * W * 16;
  if (ok) {
    for (int ky = 0; ky < 3; ++ky) {
      const int Y = 2 * oy + ky - 1;
      if (Y < 0 || Y >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int X = 2 * ox + kx - 1;
        if (X < 0 || X >= W) continue;
        const uint4* ip = reinterpret_cast<const uint4*>(base + ((long long)Y * W + X) * 16);
        const uint4 r0 = ip[0], r1 = ip[1];
        const uint32_t ru[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        const float* wp = sw3 + (ky * 3 + kx) * 16 * 64 + q * 16;
#pragma unroll
        for (int c2 = 0; c2 < 8; ++c2) {
          const __nv_bfloat162 p = *reinterpret_cast<const __nv_bfloat162*>(&ru[c2]);
          const float v0 = __low2float(p), v1 = __high2float(p);
          const float4* w0 = reinterpret_cast<const float4*>(wp + (2 * c2) * 64);
          const float4* w1 = reinterpret_cast<const float4*>(wp + (2 * c2 + 1) * 64);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 a = w0[c4], d = w1[c4];
            acc[4 * c4] += v0 * a.x + v1 * d.x;
            acc[4 * c4 + 1] += v0 * a.y + v1 * d.y;
            acc[4 * c4 + 2] += v0 * a.z + v1 * d.z;
            acc[4 * c4 + 3] += v0 * a.w + v1 * d.w;
          }
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int co = 0; co < 16; ++co) s += acc[co];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  const float mean = s * (1.f / 64.f);
  float var = 0.f;
#pragma unroll
  for (int co = 0; co < 16; ++co) {
    acc[co] -= mean;
    var += acc[co] * acc[co];
  }
  var += __shfl_xor_sync(0xffffffffu, var, 1);
  var += __shfl_xor_sync(0xffffffffu, var, 2);
  const float rstd = rsqrtf(var * (1.f / 64.f) + eps);
  if (!ok) return;
  uint32_t pk[8];
#pragma unroll
  for (int co = 0; co < 8; ++co) {
    const int c = q * 16 + 2 * co;
    pk[co] = pack_bf16x2(gelu_erf(acc[2 * co] * rstd * lnw[c] + lnb[c]),
                         gelu_erf(acc[2 * co + 1] * rstd * lnw[c + 1] + lnb[c + 1]));
  }
  uint4* o = reinterpret_cast<uint4*>(out + (((long long)b * OH + oy) * OW + ox) * 64 + q * 16);
  o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
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
  VLS_CUDA(launch_k(mds2_kernel, dim3(dim3((W / 2 + 31) / 32, (H / 2 + 7) / 8, B)), dim3(256), 0, stream, reinterpret_cast<const bf16*>(in), H, W, wgt, cb, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_mds3(const void* in, int B, int H, int W, const float* wgt, const float* cb, const float* lnw, const float* lnb,
                float eps, void* out, cudaStream_t stream) {
  const int smem = 9 * 16 * 64 * 4;
  VLS_CUDA(launch_k(mds3_kernel, dim3(dim3((W / 2 + 15) / 16, (H / 2 + 3) / 4, B)), dim3(256), smem, stream, reinterpret_cast<const bf16*>(in), H, W, wgt, cb, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
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

int launch_dwconv7_ln(const float* x, int B, int H, int W, const float* wgt, const float* cb, const float* lnw,
                      const float* lnb, float eps, void* out, cudaStream_t stream) {
  VLS_CUDA(launch_k(dwconv7_ln_kernel, dim3(dim3((W + 7) / 8, H, B)), dim3(256), 0, stream, x, H, W, wgt, cb, lnw, lnb, eps, reinterpret_cast<bf16*>(out)));
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
