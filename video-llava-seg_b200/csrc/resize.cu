// Bilinear resize, align_corners=False (F.interpolate semantics used at sam2_base.py:373-378 and
// sam2_video_predictor.py:416-421): one thread per 4 consecutive output pixels, 128-bit stores.
#include "common.cuh"
#include "kernels.h"

namespace vls {
namespace {

// one output sample; shared by both kernels so that (resize > t) and the fused binarisation agree bit for bit
__device__ __forceinline__ float bilerp(const float* __restrict__ src, int h, int w, int y0, int y1, float ly, int X, float sx) {
  const float fx = fmaxf((X + 0.5f) * sx - 0.5f, 0.f);
  const int x0 = min((int)fx, w - 1), x1 = min(x0 + 1, w - 1);
  const float lx = fx - x0;
  const float a = src[y0 * w + x0], b = src[y0 * w + x1], c = src[y1 * w + x0], d = src[y1 * w + x1];
  return (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * c + lx * d);
}

// Output stage of the tracker (sam2_video_predictor.py:404-424 followed by the consumer's `> 0`, e.g.
// llava/inference/utils.py:71-85): bilinear up-sampling + threshold in one pass.  The f32 video-resolution logits
// (4 B/pixel written, then read again by the threshold) never exist; the result is 1 byte/pixel and/or 1 bit/pixel
// (np.packbits order: first pixel = most significant bit).  One thread per 8 consecutive output pixels.
__global__ void resize_binarize_kernel(const float* __restrict__ in, int n, int h, int w, int H, int W, float sy, float sx,
                                       float thresh, uint8_t* __restrict__ out_u8, uint8_t* __restrict__ out_bits) {
  pdl_enter();
  const int W8 = (W + 7) >> 3;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)n * H * W8) return;
  const int x8 = (int)(id % W8), Y = (int)((id / W8) % H), img = (int)(id / ((long long)W8 * H));
  const float* src = in + (long long)img * h * w;
  const float fy = fmaxf((Y + 0.5f) * sy - 0.5f, 0.f);
  const int y0 = min((int)fy, h - 1), y1 = min(y0 + 1, h - 1);
  const float ly = fy - y0;
  uint32_t bits = 0;
  uint8_t px[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int X = x8 * 8 + i;
    const bool on = X < W && bilerp(src, h, w, y0, y1, ly, min(X, W - 1), sx) > thresh;
    px[i] = on ? 1 : 0;
    bits |= (on ? 1u : 0u) << (7 - i);
  }
  if (out_bits) out_bits[((long long)img * H + Y) * W8 + x8] = (uint8_t)bits;
  if (out_u8) {
    uint8_t* o = out_u8 + ((long long)img * H + Y) * W + x8 * 8;
    if ((W & 7) == 0) {
      uint2 v;
      v.x = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
      v.y = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
      *reinterpret_cast<uint2*>(o) = v;
    } else {
      for (int i = 0; i < 8 && x8 * 8 + i < W; ++i) o[i] = px[i];
    }
  }
}

__global__ void resize_bilinear_kernel(const float* __restrict__ in, int n, int h, int w, float* __restrict__ out, int H,
                                       int W, float sy, float sx) {
  pdl_enter();
  const int W4 = (W + 3) >> 2;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)n * H * W4) return;
  const int x4 = (int)(id % W4), Y = (int)((id / W4) % H), img = (int)(id / ((long long)W4 * H));
  const float* src = in + (long long)img * h * w;
  const float fy = fmaxf((Y + 0.5f) * sy - 0.5f, 0.f);
  const int y0 = min((int)fy, h - 1), y1 = min(y0 + 1, h - 1);
  const float ly = fy - y0;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = bilerp(src, h, w, y0, y1, ly, min(x4 * 4 + i, W - 1), sx);
  float* o = out + ((long long)img * H + Y) * W + x4 * 4;
  if ((W & 3) == 0) {
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int i = 0; i < 4 && x4 * 4 + i < W; ++i) o[i] = v[i];
  }
}

}  // namespace

int launch_resize_bilinear(const float* in, int n, int h, int w, float* out, int H, int W, cudaStream_t stream) {
  VLS_REQUIRE(n >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "resize: bad shape");
  if (n == 0) return 0;
  const long long total = (long long)n * H * ((W + 3) / 4);
  VLS_CUDA(launch_k(resize_bilinear_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, in, n, h, w, out, H, W, (float)h / H, (float)w / W));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_resize_binarize(const float* in, int n, int h, int w, int H, int W, float thresh, uint8_t* out_u8,
                           uint8_t* out_bits, cudaStream_t stream) {
  VLS_REQUIRE(n >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "resize_binarize: bad shape");
  VLS_REQUIRE(out_u8 || out_bits, "resize_binarize: no output requested");
  if (n == 0) return 0;
  const long long total = (long long)n * H * ((W + 7) / 8);
  VLS_CUDA(launch_k(resize_binarize_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, in, n, h, w, H, W, (float)h / H, (float)w / W, thresh, out_u8, out_bits));
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
