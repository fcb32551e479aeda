// Internal launcher declarations (one translation unit per kernel family). All launchers enqueue on
// `stream`, never synchronise, never allocate; they return 0 or a non-zero code with vls::set_error.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "host.h"

namespace vls {

// ---------------------------------------------------------------- tcgen05 GEMM (gemm_tc.cu)
// C[b][m][n] = epilogue( sum_k A[b][m][k] * W[b?][n][k] ),  A/W bf16 K-major, f32 accumulate in TMEM.
struct GemmArgs {
  const void* A = nullptr;   // bf16 [batch][M][lda]
  long long lda = 0, a_bstride = 0;
  const void* W = nullptr;   // bf16 [batch or 1][N][ldw]
  long long ldw = 0, w_bstride = 0;  // w_bstride == 0 -> shared across the batch
  int M = 0, N = 0, K = 0, batch = 1;
  const float* bias = nullptr;
  int bias_mode = 0;         // 0 none, 1 per output column (n), 2 per output row (m)
  int act = 0;               // 0 none, 1 ReLU, 2 GELU(erf)
  const float* rope_cos = nullptr;  // [rope_period][128]; rotates column pairs inside every 256-column block
  const float* rope_sin = nullptr;
  int rope_period = 0;       // table row = m % rope_period
  int rope_rows = 0;         // only rows m < rope_rows are rotated
  const float* residual = nullptr;  // f32, added after activation
  long long ld_res = 0, res_bstride = 0;
  void* C = nullptr;
  int c_bf16 = 1;            // 1: bf16 output, 0: f32 output
  long long ldc = 0, c_bstride = 0;
};
int launch_gemm(const GemmArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------- tcgen05 flash attention (attn_tc.cu)
// O[b][q][:] = softmax(Q[b][q][:] . K[b][k][:] * scale) @ V  with head dim 256, one head.
// Q: bf16 [B][Nq][ldq], K: bf16 [B][Nk][ldk], Vt: bf16 [B][256][ldvt] (V transposed: row = channel).
struct AttnArgs {
  const void* Q = nullptr; long long ldq = 0, q_bstride = 0;
  const void* K = nullptr; long long ldk = 0, k_bstride = 0;
  const void* Vt = nullptr; long long ldvt = 0, vt_bstride = 0;
  int B = 1, Nq = 0, Nk = 0;
  float scale = 0.0625f;
  int splits = 1;            // KV splits per query tile (>=1)
  void* O = nullptr;         // bf16 [B][Nq][ldo]
  long long ldo = 0, o_bstride = 0;
  float* part_o = nullptr;   // workspace when splits > 1: f32 [B][splits][Nq][256]
  float* part_ml = nullptr;  // f32 [B][splits][Nq][2] (running max in log2 domain, sum)
};
size_t attn_workspace_bytes(int B, int Nq, int splits);
int attn_pick_splits(int B, int Nq, int Nk);
int launch_attention(const AttnArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------- connected components (cc.cu)
size_t cc_workspace_bytes(int n, int h, int w, bool fill);
int launch_cc_label(const uint8_t* img, int n, int h, int w, int32_t* labels, int32_t* counts, void* ws,
                    size_t ws_bytes, cudaStream_t stream);
int launch_fill_holes(float* scores, int n, int h, int w, int max_area, float fill_value, void* ws, size_t ws_bytes,
                      cudaStream_t stream);

}  // namespace vls
