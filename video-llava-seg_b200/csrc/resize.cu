// Bilinear resize, align_corners=False (F.interpolate semantics used at sam2_base.py:373-378 and
// sam2_video_predictor.py:416-421): one thread per 4 consecutive output pixels, 128-bit stores.
#include "common.cuh"
#include "kernels.h"

namespace vls {
namespace {

__global__ void resize_bilinear_kernel(const float* __restrict__ in, int n, int h, int w, float* __restrict__ out, int H,
                                       int W, float sy, float sx) {
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
  for (int i = 0; i < 4; ++i) {
    const int X = x4 * 4 + i;
    const float fx = fmaxf((X + 0.5f) * sx - 0.5f, 0.f);
    const int x0 = min((int)fx, w - 1), x1 = min(x0 + 1, w - 1);
    const float lx = fx - x0;
    const float a = src[y0 * w + x0], b = src[y0 * w + x1], c = src[y1 * w + x0], d = src[y1 * w + x1];
    v[i] = (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * c + lx * d);
  }
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
  resize_bilinear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, n, h, w, out, H, W, (float)h / H,
                                                                              (float)w / W);
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
