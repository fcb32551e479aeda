// Bandwidth kernels around the GEMMs: LayerNorm over 256 channels (warp per row, 128-bit loads),
// layout/precision preparation (seq-first / NCHW -> batch-major token rows, bf16 rounding, fused
// "a + alpha*b"), and the small token-side linears of the mask decoder (warp per output).
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

__device__ __forceinline__ float ld_any(const void* p, long long i, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}

// ------------------------------------------------------------------ LayerNorm, C = 256
// x: f32 rows [B][T][256] contiguous. One warp per row, lane owns channels [8*lane, 8*lane+8).
__global__ void ln256_kernel(const float* __restrict__ x, int B, int T, const float* __restrict__ w,
                             const float* __restrict__ bsh, float eps, int gelu, float* __restrict__ out_f32,
                             long long f_sb, long long f_st, bf16* __restrict__ out_bf16, long long h_sb,
                             long long h_st) {
  pdl_enter();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * T) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(row / T), t = (int)(row % T);
  const float4* src = reinterpret_cast<const float4*>(x + row * 256 + lane * 8);
  const float4 a = src[0], c = src[1];
  float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / 256.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] -= mean;
    q += v[i] * v[i];
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256.0f) + eps);
  const float4 w0 = reinterpret_cast<const float4*>(w + lane * 8)[0], w1 = reinterpret_cast<const float4*>(w + lane * 8)[1];
  const float4 b0 = reinterpret_cast<const float4*>(bsh + lane * 8)[0], b1 = reinterpret_cast<const float4*>(bsh + lane * 8)[1];
  const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = v[i] * rstd * ww[i] + bb[i];
    if (gelu) v[i] = gelu_erf(v[i]);
  }
  if (out_f32) {
    float4* o = reinterpret_cast<float4*>(out_f32 + b * f_sb + t * f_st + lane * 8);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (out_bf16) {
    *reinterpret_cast<uint4*>(out_bf16 + b * h_sb + t * h_st + lane * 8) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// ------------------------------------------------------------------ out[b][t][c] = a + alpha * p
// a, p: element (t, b, c) at  t*s_t + b*s_b + c  (f32 or bf16); C % 4 == 0.  out rows contiguous.
// four consecutive channels at once (VEC: 8- / 16-byte aligned bases, strides multiples of 4 elements)
template <bool VEC>
__device__ __forceinline__ void ld4_any(const void* p, long long i, int is_bf16, float v[4]) {
  if (VEC) {
    if (is_bf16) {
      const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p) + i);
      const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
    } else {
      const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld_any(p, i + k, is_bf16);
  }
}

template <bool VEC>
__global__ void axpy_rows_kernel(const void* __restrict__ a, int a_bf16, long long a_st, long long a_sb,
                                 const void* __restrict__ p, int p_bf16, long long p_st, long long p_sb, float alpha,
                                 int B, int T, int C, float* __restrict__ out_f32, bf16* __restrict__ out_bf16,
                                 long long out_sb) {
  pdl_enter();
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total4 = (long long)B * T * C / 4;
  if (i4 >= total4) return;
  const long long e_in = i4 * 4;
  const int c = (int)(e_in % C);
  const long long bt = e_in / C;
  const int t = (int)(bt % T), b = (int)(bt / T);
  const long long e = (long long)b * out_sb + (long long)t * C + c;   // out_sb = T * C unless the rows are a slice of a taller matrix
  float v[4];
  ld4_any<VEC>(a, t * a_st + b * a_sb + c, a_bf16, v);
  if (p) {
    float q[4];
    ld4_any<VEC>(p, t * p_st + b * p_sb + c, p_bf16, q);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += alpha * q[i];
  }
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + e) = make_float4(v[0], v[1], v[2], v[3]);
  if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + e) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

// ------------------------------------------------------------------ LayerNorm2d over 64 channels + GELU (mask down-sampler)
// x: f32 rows [rows][64]; half a warp per row, lane owns 4 channels; out bf16 rows.
__global__ void ln64_gelu_kernel(const float* __restrict__ x, long long rows, const float* __restrict__ w,
                                 const float* __restrict__ bsh, float eps, bf16* __restrict__ out) {
  pdl_enter();
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int c = (threadIdx.x & 15) * 4;
  const bool ok = row < rows;
  float4 v = ok ? *reinterpret_cast<const float4*>(x + row * 64 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 64.0f);
  v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
  float q = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / 64.0f) + eps);
  const float4 g = *reinterpret_cast<const float4*>(w + c), b = *reinterpret_cast<const float4*>(bsh + c);
  if (ok)
    *reinterpret_cast<uint2*>(out + row * 64 + c) =
        make_uint2(pack_bf16x2(gelu_erf(v.x * rstd * g.x + b.x), gelu_erf(v.y * rstd * g.y + b.y)),
                   pack_bf16x2(gelu_erf(v.z * rstd * g.z + b.z), gelu_erf(v.w * rstd * g.w + b.w)));
}

// ------------------------------------------------------------------ NCHW (+ optional addend) -> token rows
// in element (b, c, y, x) at b*sb + c*sc + y*sh + x*sw (any of them may be 0 for expanded views).
struct Strides4 { long long sb, sc, sh, sw; };
__global__ void nchw_to_rows_kernel(const void* __restrict__ in, int in_bf16, Strides4 si, const void* __restrict__ add,
                                    int add_bf16, Strides4 sa, int C, int H, int W, float* __restrict__ out_f32,
                                    bf16* __restrict__ out_bf16) {
  pdl_enter();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int T = H * W;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    float v = 0.f;
    if (c < C && t < T) {
      const int y = t / W, x = t % W;
      v = ld_any(in, b * si.sb + c * si.sc + y * si.sh + x * si.sw, in_bf16);
      if (add) v += ld_any(add, b * sa.sb + c * sa.sc + y * sa.sh + x * sa.sw, add_bf16);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < C) {
      const float v = tile[threadIdx.x][i];
      const long long o = ((long long)b * T + t) * C + c;
      if (out_f32) out_f32[o] = v;
      if (out_bf16) out_bf16[o] = __float2bfloat16_rn(v);
    }
  }
}

// Channel-contiguous input (sc == 1: the "NCHW" tensor is a permuted view of token rows, which is how the memory attention
// hands pix_feat_with_mem to the decoder): no transpose, a thread moves 4 channels of one token.  The generic kernel above
// reads such a view with 4 useful bytes per 32-byte sector (7.6 us for 4 MB).
__global__ void chlast_to_rows_kernel(const void* __restrict__ in, int in_bf16, Strides4 si, const void* __restrict__ add,
                                      int add_bf16, Strides4 sa, int C, int H, int W, float* __restrict__ out_f32,
                                      bf16* __restrict__ out_bf16, long long total4) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = C >> 2, T = H * W;
  const int c = (int)(i % c4) * 4;
  const int t = (int)((i / c4) % T), b = (int)(i / ((long long)c4 * T));
  const int y = t / W, x = t % W;
  const long long ib = b * si.sb + y * si.sh + x * si.sw + c;
  float v[4];
  if (in_bf16) {
    const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(in) + ib);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.x), d = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(d); v[3] = __high2float(d);
  } else {
    const float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + ib);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  if (add) {
    const long long ab = b * sa.sb + y * sa.sh + x * sa.sw;
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] += ld_any(add, ab + (c + k) * sa.sc, add_bf16);
  }
  const long long o = ((long long)b * T + t) * C + c;
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(v[0], v[1], v[2], v[3]);
  if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + o) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

// ------------------------------------------------------------------ token rows -> NCHW f32/bf16 (+ gated channel vector)
// out[b][c][t] = in[b][t][c] + gate[b] * vec[c]
__global__ void rows_to_nchw_kernel(const float* __restrict__ in, int C, int T, const float* __restrict__ gate,
                                    const float* __restrict__ vec, float* __restrict__ out_f32,
                                    bf16* __restrict__ out_bf16) {
  pdl_enter();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && c < C) ? in[((long long)b * T + t) * C + c] : 0.f;
  }
  __syncthreads();
  const float g = gate ? gate[b] : 0.f;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (t < T && c < C) {
      const float v = tile[threadIdx.x][i] + (vec ? g * vec[c] : 0.f);
      const long long o = ((long long)b * C + c) * T + t;
      if (out_f32) out_f32[o] = v;
      if (out_bf16) out_bf16[o] = __float2bfloat16_rn(v);
    }
  }
}

// ------------------------------------------------------------------ small linear (token side)
// out[g][r][n] = act( sum_k (x[g][r][k] + xadd[g][r][k]) * W[g][n][k] + bias[g][n] ) + res[g][r][n]
// one warp per output element; x f32, W bf16. K % 8 == 0.
struct SmallLin {
  const float* x; long long x_sg, x_sr;
  const float* xadd; long long xa_sg, xa_sr;
  const bf16* W; long long w_sg;
  const float* bias; long long b_sg;
  const float* res; long long r_sg, r_sr;
  float* out; long long o_sg, o_sr;
  int G, R, N, K, act;  // act: 0 none, 1 relu, 3 sigmoid
};
__device__ __forceinline__ void small_linear_warp(const SmallLin& p, long long gw, int lane) {
  const int n = (int)(gw % p.N);
  const int r = (int)((gw / p.N) % p.R);
  const int g = (int)(gw / ((long long)p.N * p.R));
  const float* x = p.x + g * p.x_sg + r * p.x_sr;
  const float* xa = p.xadd ? p.xadd + g * p.xa_sg + r * p.xa_sr : nullptr;
  const bf16* w = p.W + g * p.w_sg + (long long)n * p.K;
  float acc = 0.f;
  for (int k = lane * 8; k < p.K; k += 256) {
    const uint4 wv = *reinterpret_cast<const uint4*>(w + k);
    const float4 x0 = *reinterpret_cast<const float4*>(x + k), x1 = *reinterpret_cast<const float4*>(x + k + 4);
    float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    if (xa) {
      const float4 a0 = *reinterpret_cast<const float4*>(xa + k), a1 = *reinterpret_cast<const float4*>(xa + k + 4);
      xv[0] += a0.x; xv[1] += a0.y; xv[2] += a0.z; xv[3] += a0.w;
      xv[4] += a1.x; xv[5] += a1.y; xv[6] += a1.z; xv[7] += a1.w;
    }
    const uint32_t wu[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 w2 = *reinterpret_cast<const __nv_bfloat162*>(&wu[i]);
      acc += xv[2 * i] * __low2float(w2) + xv[2 * i + 1] * __high2float(w2);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (p.bias) acc += p.bias[g * p.b_sg + n];
    if (p.act == 1) acc = fmaxf(acc, 0.f);
    if (p.act == 3) acc = 1.0f / (1.0f + __expf(-acc));
    if (p.res) acc += p.res[g * p.r_sg + r * p.r_sr + n];
    p.out[g * p.o_sg + r * p.o_sr + n] = acc;
  }
}

__global__ void small_linear_kernel(const SmallLin p) {
  pdl_enter();
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (long long)p.G * p.R * p.N) return;
  small_linear_warp(p, gw, threadIdx.x & 31);
}

// Several INDEPENDENT small linears in one launch (the token side of the mask decoder is a chain of 2-3 us kernels:
// q/k/v projections, the first layers of the hyper-network / IoU / object-score heads, ... run side by side).
constexpr int SMALL_LIN_MAX = 4;
struct SmallLinMulti {
  SmallLin p[SMALL_LIN_MAX];
  long long first[SMALL_LIN_MAX + 1];   // prefix sums of G*R*N
};
__global__ void small_linear_multi_kernel(const SmallLinMulti m, int count) {
  pdl_enter();
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= m.first[count]) return;
  int i = 0;
#pragma unroll
  for (int k = 1; k < SMALL_LIN_MAX; ++k) i += (k < count && gw >= m.first[k]) ? 1 : 0;
  small_linear_warp(m.p[i], gw - m.first[i], threadIdx.x & 31);
}

// LayerNorm over 256 for a handful of token rows (f32 in/out), optional residual-free in-place.
__global__ void ln256_small_kernel(const float* __restrict__ x, long long x_sr, int rows, const float* __restrict__ w,
                                   const float* __restrict__ b, float eps, float* __restrict__ out, long long o_sr) {
  pdl_enter();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = x[row * x_sr + lane + 32 * i];
    s += v[i];
  }
  const float mean = warp_sum(s) * (1.0f / 256.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] -= mean;
    q += v[i] * v[i];
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256.0f) + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[row * o_sr + lane + 32 * i] = v[i] * rstd * w[lane + 32 * i] + b[lane + 32 * i];
}

// tokens[b][i][:] = i < n_out ? out_tokens[i][:] : sparse[b][i - n_out][:]      (mask_decoder.py:179-197)
__global__ void build_tokens_kernel(const float* __restrict__ out_tokens, int n_out, const float* __restrict__ sparse,
                                    int Ns, int B, float* __restrict__ tok_a, float* __restrict__ tok_b) {
  pdl_enter();
  const int Nt = n_out + Ns;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Nt * 256) return;
  const int c = i & 255, r = (i >> 8) % Nt, b = i / (Nt * 256);
  const float v = r < n_out ? out_tokens[r * 256 + c] : sparse[((long long)b * Ns + (r - n_out)) * 256 + c];
  tok_a[i] = v;
  tok_b[i] = v;
}

// out_bf16[b][t][c] = in[b][t][c] + gate[b] * vec[c]
__global__ void rows_gate_cast_kernel(const float* __restrict__ in, int B, int T, int C, const float* __restrict__ gate,
                                      const float* __restrict__ vec, bf16* __restrict__ out) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T * C) return;
  const int c = (int)(i % C), b = (int)(i / ((long long)T * C));
  out[i] = __float2bfloat16_rn(in[i] + ((gate && vec) ? gate[b] * vec[c] : 0.f));
}

// dst[g][r][:n] = src[g*sg + r*sr + :n]
__global__ void gather_rows_kernel(const float* __restrict__ src, long long sg, long long sr, int G, int R, int n,
                                   float* __restrict__ dst) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * R * n) return;
  const int c = i % n, r = (i / n) % R, g = i / (n * R);
  dst[i] = src[g * sg + r * sr + c];
}

// ------------------------------------------------------------------ device memory bank (graphed.py, sam2_base.py:533-646)
// bank [B][n_mem*HW + n_ptr*k][64] bf16 in key order [cond | t-6 .. t-1 | ptr(cond), ptr(t-1) .. ptr(t-(n_ptr-1))].
// Advance one frame IN PLACE: memories 1..n_mem-2 <- 2..n_mem-1, memory n_mem-1 <- new_rows; pointers 2..n_ptr-1 <-
// 1..n_ptr-2, pointer 1 <- new_ptr (f32 -> bf16).  A thread owns one 16-byte column of a memory slot (or one element
// pair of a pointer slot) and walks the chain of slots itself, so every element is read before the same thread
// overwrites it: no staging buffer, no grid barrier, one launch instead of six copy nodes.
__global__ void bank_shift_kernel(bf16* __restrict__ bank, int B, int HW, int n_mem, int n_ptr, int k,
                                  const bf16* __restrict__ new_rows, const float* __restrict__ new_ptr) {
  pdl_enter();
  const long long slot_vec = (long long)HW * 64 / 8;            // uint4 vectors per memory slot
  const long long Nk = (long long)n_mem * HW + (long long)n_ptr * k;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ptr_vec = k * 64 / 8;                                // uint4 vectors per pointer slot
  if (id < (long long)B * slot_vec) {
    const int b = (int)(id / slot_vec);
    const long long off = id % slot_vec;
    uint4* base = reinterpret_cast<uint4*>(bank + (long long)b * Nk * 64) + off;
    uint4 v[8];                                  // all loads first (independent, in flight together), then the stores
#pragma unroll
    for (int j = 1; j < 8; ++j)
      if (j + 1 < n_mem) v[j] = base[(j + 1) * slot_vec];
    const uint4 fresh = reinterpret_cast<const uint4*>(new_rows + (long long)b * HW * 64)[off];
#pragma unroll
    for (int j = 1; j < 8; ++j)
      if (j + 1 < n_mem) base[j * slot_vec] = v[j];
    base[(long long)(n_mem - 1) * slot_vec] = fresh;
  } else if (id < (long long)B * slot_vec + (long long)B * ptr_vec) {
    const long long pid = id - (long long)B * slot_vec;
    const int b = (int)(pid / ptr_vec), off = (int)(pid % ptr_vec);
    uint4* base = reinterpret_cast<uint4*>(bank + ((long long)b * Nk + (long long)n_mem * HW) * 64) + off;
    // all loads first: written as "slot j <- slot j-1" in a loop the copies are one chain of dependent L2 round trips (the
    // compiler must assume that a store may feed the next load): 14 x ~0.4 us made this kernel 8.7 us long
    for (int j = n_ptr - 1; j >= 17; --j) base[j * ptr_vec] = base[(j - 1) * ptr_vec];   // (more than 17 pointer slots: never in use)
    uint4 pv[16];
#pragma unroll
    for (int j = 1; j < 16; ++j)
      if (j + 1 < n_ptr) pv[j] = base[j * ptr_vec];
#pragma unroll
    for (int j = 1; j < 16; ++j)
      if (j + 1 < n_ptr) base[(j + 1) * ptr_vec] = pv[j];
    const float* src = new_ptr + (long long)b * k * 64 + off * 8;
    uint4 v;
    v.x = pack_bf16x2(src[0], src[1]); v.y = pack_bf16x2(src[2], src[3]);
    v.z = pack_bf16x2(src[4], src[5]); v.w = pack_bf16x2(src[6], src[7]);
    base[ptr_vec] = v;
  }
}

// up to 8 device-to-device copies in one launch (per-frame snapshots of the graph's static outputs).  A copy is done in
// 16-byte vectors, or -- small tensors whose size / address is not a multiple of 16, e.g. the [B,1] object scores -- in
// bytes (unit = 1): such a tensor used to fall back to a cudaMemcpy of its own between two graph replays.
struct MultiCopy {
  const char* src[8];
  char* dst[8];
  long long items[8];    // work items per copy: 16-byte vectors or bytes
  int unit[8];           // 16 or 1
  long long first[9];    // prefix sums of items
};
__global__ void multi_copy_kernel(const MultiCopy m, int n) {
  pdl_enter();
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= m.first[n]) return;
  int c = 0;
#pragma unroll
  for (int i = 1; i < 8; ++i) c += (i < n && id >= m.first[i]) ? 1 : 0;
  const long long k = id - m.first[c];
  if (m.unit[c] == 16) reinterpret_cast<uint4*>(m.dst[c])[k] = reinterpret_cast<const uint4*>(m.src[c])[k];
  else m.dst[c][k] = m.src[c][k];
}

}  // namespace

// rows [B][T][64] bf16 -> transposed [B][64][ld] bf16 (ld >= T): the K-major V^T operand of the attention kernel when the
// row-major (MN-major) value path is switched off (vls_set_tuning "attn_v_rows" 0).  64 x 64 tiles through shared memory.
__global__ void transpose_rows64_kernel(const bf16* __restrict__ in, int T, bf16* __restrict__ out, long long ld) {
  pdl_enter();
  __shared__ bf16 tile[64][66];
  const int b = blockIdx.y, t0 = blockIdx.x * 64;
  const bf16* src = in + ((long long)b * T + t0) * 64;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i >> 6, c = i & 63;
    tile[r][c] = (t0 + r < T) ? src[(long long)r * 64 + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  bf16* dst = out + (long long)b * 64 * ld + t0;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int c = i >> 6, r = i & 63;
    if (t0 + r < ld) dst[(long long)c * ld + r] = tile[r][c];
  }
}

int launch_transpose_rows64(const void* in, int B, int T, void* out, long long ld, cudaStream_t stream) {
  VLS_REQUIRE(in && out && B > 0 && T > 0 && ld >= T, "transpose_rows64: bad arguments");
  VLS_CUDA(launch_k(transpose_rows64_kernel, dim3((unsigned)((ld + 63) / 64), B), dim3(256), 0, stream,
                    reinterpret_cast<const bf16*>(in), T, reinterpret_cast<bf16*>(out), ld));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_bank_shift(void* bank, int B, int HW, int n_mem, int n_ptr, int k, const void* new_rows, const float* new_ptr,
                      cudaStream_t stream) {
  VLS_REQUIRE(bank && new_rows && new_ptr, "bank_shift: null argument");
  VLS_REQUIRE(B >= 1 && HW >= 1 && n_mem >= 2 && n_mem <= 9 && n_ptr >= 2 && k >= 1 && (HW * 64) % 8 == 0, "bank_shift: bad shape");
  const long long total = (long long)B * HW * 8 + (long long)B * k * 8;
  VLS_CUDA(launch_k(bank_shift_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream,
                    reinterpret_cast<bf16*>(bank), B, HW, n_mem, n_ptr, k, reinterpret_cast<const bf16*>(new_rows), new_ptr));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_multi_copy(const void* const* src, void* const* dst, const size_t* bytes, int n, cudaStream_t stream) {
  VLS_REQUIRE(n >= 0 && n <= 8, "multi_copy: between 0 and 8 copies per launch");
  if (n == 0) return 0;
  MultiCopy m;
  m.first[0] = 0;
  for (int i = 0; i < 8; ++i) {
    m.src[i] = nullptr; m.dst[i] = nullptr; m.items[i] = 0; m.unit[i] = 16;
    if (i < n) {
      VLS_REQUIRE(src[i] && dst[i], "multi_copy: copy %d has a null pointer", i);
      const bool vec = bytes[i] % 16 == 0 && ((uintptr_t)src[i] % 16) == 0 && ((uintptr_t)dst[i] % 16) == 0;
      VLS_REQUIRE(vec || bytes[i] <= 4096, "multi_copy: copy %d must be 16-byte aligned and a multiple of 16 bytes (or at most 4096 bytes)", i);
      m.src[i] = reinterpret_cast<const char*>(src[i]);
      m.dst[i] = reinterpret_cast<char*>(dst[i]);
      m.unit[i] = vec ? 16 : 1;
      m.items[i] = vec ? (long long)(bytes[i] / 16) : (long long)bytes[i];
    }
    m.first[i + 1] = m.first[i] + m.items[i];
  }
  if (m.first[n] == 0) return 0;
  VLS_CUDA(launch_k(multi_copy_kernel, dim3((unsigned)((m.first[n] + 255) / 256)), dim3(256), 0, stream, m, n));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_build_tokens(const float* out_tokens, int n_out, const float* sparse, int Ns, int B, float* tok_a, float* tok_b,
                        cudaStream_t stream) {
  const int total = B * (n_out + Ns) * 256;
  VLS_CUDA(launch_k(build_tokens_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, out_tokens, n_out, sparse, Ns, B, tok_a, tok_b));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_rows_gate_cast(const float* in, int B, int T, int C, const float* gate, const float* vec, void* out,
                          cudaStream_t stream) {
  const long long total = (long long)B * T * C;
  VLS_CUDA(launch_k(rows_gate_cast_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, in, B, T, C, gate, vec, reinterpret_cast<bf16*>(out)));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_gather_rows(const float* src, long long sg, long long sr, int G, int R, int n, float* dst, cudaStream_t stream) {
  const int total = G * R * n;
  if (total == 0) return 0;
  VLS_CUDA(launch_k(gather_rows_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, src, sg, sr, G, R, n, dst));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_ln256(const float* x, int B, int T, const float* w, const float* b, float eps, int gelu, float* out_f32,
                 long long f_sb, long long f_st, void* out_bf16, long long h_sb, long long h_st, cudaStream_t stream) {
  const long long rows = (long long)B * T;
  if (rows == 0) return 0;
  const int wpb = 8;
  VLS_CUDA(launch_k(ln256_kernel, dim3((unsigned)((rows + wpb - 1) / wpb)), dim3(wpb * 32), 0, stream,  x, B, T, w, b, eps, gelu, out_f32, f_sb, f_st, reinterpret_cast<bf16*>(out_bf16), h_sb, h_st));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_axpy_rows_strided(const void* a, int a_bf16, long long a_st, long long a_sb, const void* p, int p_bf16,
                             long long p_st, long long p_sb, float alpha, int B, int T, int C, float* out_f32, void* out_bf16,
                             long long out_sb, cudaStream_t stream) {
  VLS_REQUIRE(C % 4 == 0 && out_sb % 4 == 0, "axpy_rows: C and the output batch stride must be multiples of 4");
  const long long total4 = (long long)B * T * C / 4;
  if (total4 == 0) return 0;
  auto vec_ok = [](const void* q, int is_bf16, long long st, long long sb) {
    return q == nullptr || ((reinterpret_cast<uintptr_t>(q) % (is_bf16 ? 8 : 16)) == 0 && st % 4 == 0 && sb % 4 == 0);
  };
  // a + alpha * p is evaluated identically on both paths (same operations in the same order per element)
  auto kern = (vec_ok(a, a_bf16, a_st, a_sb) && vec_ok(p, p_bf16, p_st, p_sb)) ? axpy_rows_kernel<true> : axpy_rows_kernel<false>;
  VLS_CUDA(launch_k(kern, dim3((unsigned)((total4 + 255) / 256)), dim3(256), 0, stream,  a, a_bf16, a_st, a_sb, p, p_bf16, p_st, p_sb, alpha, B, T, C, out_f32, reinterpret_cast<bf16*>(out_bf16), out_sb));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_axpy_rows(const void* a, int a_bf16, long long a_st, long long a_sb, const void* p, int p_bf16,
                     long long p_st, long long p_sb, float alpha, int B, int T, int C, float* out_f32, void* out_bf16,
                     cudaStream_t stream) {
  return launch_axpy_rows_strided(a, a_bf16, a_st, a_sb, p, p_bf16, p_st, p_sb, alpha, B, T, C, out_f32, out_bf16,
                                  (long long)T * C, stream);
}

int launch_nchw_to_rows(const void* in, int in_bf16, const long long si[4], const void* add, int add_bf16,
                        const long long sa[4], int B, int C, int H, int W, float* out_f32, void* out_bf16,
                        cudaStream_t stream) {
  Strides4 s1{si[0], si[1], si[2], si[3]};
  Strides4 s2{0, 0, 0, 0};
  if (add) s2 = Strides4{sa[0], sa[1], sa[2], sa[3]};
  const int esz = in_bf16 ? 2 : 4;
  const bool chlast = si[1] == 1 && C % 4 == 0 && si[0] % 4 == 0 && si[2] % 4 == 0 && si[3] % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(in) % (4 * esz)) == 0;
  if (chlast) {   // channel-contiguous view of token rows: plain vectorised copy (+ addend)
    const long long total4 = (long long)B * H * W * (C / 4);
    VLS_CUDA(launch_k(chlast_to_rows_kernel, dim3((unsigned)((total4 + 255) / 256)), dim3(256), 0, stream, in, in_bf16, s1, add,
                      add_bf16, s2, C, H, W, out_f32, reinterpret_cast<bf16*>(out_bf16), total4));
    VLS_POST_LAUNCH(1);
    return 0;
  }
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, B), blk(32, 8);
  VLS_CUDA(launch_k(nchw_to_rows_kernel, dim3(grid), dim3(blk), 0, stream, in, in_bf16, s1, add, add_bf16, s2, C, H, W, out_f32, reinterpret_cast<bf16*>(out_bf16)));
  VLS_POST_LAUNCH(1);
  return 0;
}

int g_mds3_tc = 1;   // mask down-sampler stage 3 as im2col + tcgen05 GEMM (vls_set_tuning "mds3_tc")

int launch_ln64_gelu(const float* x, long long rows, const float* w, const float* b, float eps, void* out, cudaStream_t stream) {
  VLS_CUDA(launch_k(ln64_gelu_kernel, dim3((unsigned)((rows * 16 + 255) / 256)), dim3(256), 0, stream, x, rows, w, b, eps,
                    reinterpret_cast<bf16*>(out)));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_rows_to_nchw(const float* in, int B, int C, int T, const float* gate, const float* vec, float* out_f32,
                        void* out_bf16, cudaStream_t stream) {
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), blk(32, 8);
  VLS_CUDA(launch_k(rows_to_nchw_kernel, dim3(grid), dim3(blk), 0, stream, in, C, T, gate, vec, out_f32, reinterpret_cast<bf16*>(out_bf16)));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_small_linear(const SmallLinArgs& a, cudaStream_t stream) {
  VLS_REQUIRE(a.K % 8 == 0, "small_linear: K must be a multiple of 8");
  SmallLin p;
  p.x = a.x; p.x_sg = a.x_sg; p.x_sr = a.x_sr;
  p.xadd = a.xadd; p.xa_sg = a.xa_sg; p.xa_sr = a.xa_sr;
  p.W = reinterpret_cast<const bf16*>(a.W); p.w_sg = a.w_sg;
  p.bias = a.bias; p.b_sg = a.b_sg;
  p.res = a.res; p.r_sg = a.r_sg; p.r_sr = a.r_sr;
  p.out = a.out; p.o_sg = a.o_sg; p.o_sr = a.o_sr;
  p.G = a.G; p.R = a.R; p.N = a.N; p.K = a.K; p.act = a.act;
  const long long total = (long long)a.G * a.R * a.N;
  if (total == 0) return 0;
  const int wpb = 8;
  VLS_CUDA(launch_k(small_linear_kernel, dim3((unsigned)((total + wpb - 1) / wpb)), dim3(wpb * 32), 0, stream, p));
  VLS_POST_LAUNCH(1);
  return 0;
}

static int fill_small_lin(const SmallLinArgs& a, SmallLin* p) {
  VLS_REQUIRE(a.K % 8 == 0, "small_linear: K must be a multiple of 8");
  p->x = a.x; p->x_sg = a.x_sg; p->x_sr = a.x_sr;
  p->xadd = a.xadd; p->xa_sg = a.xa_sg; p->xa_sr = a.xa_sr;
  p->W = reinterpret_cast<const bf16*>(a.W); p->w_sg = a.w_sg;
  p->bias = a.bias; p->b_sg = a.b_sg;
  p->res = a.res; p->r_sg = a.r_sg; p->r_sr = a.r_sr;
  p->out = a.out; p->o_sg = a.o_sg; p->o_sr = a.o_sr;
  p->G = a.G; p->R = a.R; p->N = a.N; p->K = a.K; p->act = a.act;
  return 0;
}

int launch_small_linear_multi(const SmallLinArgs* a, int count, cudaStream_t stream) {
  VLS_REQUIRE(count >= 1 && count <= SMALL_LIN_MAX, "small_linear_multi: between 1 and %d problems", SMALL_LIN_MAX);
  if (count == 1) return launch_small_linear(a[0], stream);
  SmallLinMulti m;
  m.first[0] = 0;
  for (int i = 0; i < SMALL_LIN_MAX; ++i) {
    const SmallLinArgs& src = a[i < count ? i : 0];
    VLS_TRY(fill_small_lin(src, &m.p[i]));
    m.first[i + 1] = m.first[i] + (i < count ? (long long)src.G * src.R * src.N : 0);
  }
  if (m.first[count] == 0) return 0;
  const int wpb = 8;
  VLS_CUDA(launch_k(small_linear_multi_kernel, dim3((unsigned)((m.first[count] + wpb - 1) / wpb)), dim3(wpb * 32), 0, stream, m,
                    count));
  VLS_POST_LAUNCH(1);
  return 0;
}

int launch_ln256_small(const float* x, long long x_sr, int rows, const float* w, const float* b, float eps, float* out,
                       long long o_sr, cudaStream_t stream) {
  if (rows == 0) return 0;
  VLS_CUDA(launch_k(ln256_small_kernel, dim3((rows + 3) / 4), dim3(128), 0, stream, x, x_sr, rows, w, b, eps, out, o_sr));
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
