// Internal launcher declarations (one translation unit per kernel family). All launchers enqueue on
// `stream`, never synchronise, never allocate; they return 0 or a non-zero code with vls::set_error.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "host.h"
#include "vls_b200.h"

namespace vls {

// ---------------------------------------------------------------- tcgen05 GEMM (gemm_tc.cu)
// C[b][m][n] = epilogue( sum_k A[b][m][k] * W[b?][n][k] ),  A/W bf16 K-major, f32 accumulate in TMEM.
struct GemmArgs {
  const void* A = nullptr;   // bf16 [batch][M][lda]
  long long lda = 0, a_bstride = 0;
  const void* W = nullptr;   // bf16 [batch or 1][N][ldw]
  long long ldw = 0, w_bstride = 0;  // w_bstride == 0 -> shared across the batch
  int M = 0, N = 0, K = 0, batch = 1;
  // optional explicit batch maps: operand batch index = (blockIdx.z / div) % batches  (0 = derive from strides)
  int a_batches = 0, a_div = 1, w_batches = 0, w_div = 1;
  const float* bias = nullptr;
  int bias_mode = 0;         // 0 none, 1 per output column (n), 2 per output row (m)
  long long bias_bstride = 0; int bias_batches = 0, bias_div = 1;  // bias + ((z / div) % batches) * stride
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
  // optional "plane" output: column n is stored at (n / c_colblock) * c_colblock_stride + n % c_colblock (+ m * ldc): the
  // mask decoder's image-side projections are written head-major, [head][T][16], so that a head's keys are contiguous
  int c_colblock = 0; long long c_colblock_stride = 0;
};
int launch_gemm(const GemmArgs& a, cudaStream_t stream);
extern int g_gemm_ring2_above;  // K <= 256 GEMMs with more CTAs than this use a 2-stage operand ring (more CTAs resident per SM)
extern int g_gemm_bn64_below;   // GEMMs with fewer 128 x 128 tiles than this use 128 x 64 tiles

// ---------------------------------------------------------------- tcgen05 flash attention (attn_tc.cu)
// O[b][q][:] = softmax(Q[b][q][:] . K[b][k][:] * scale) @ V  with q/k head dim 256, one head, value dim dv (256 or 64).
// Q: bf16 [B][Nq][ldq], K: bf16 [B][Nk][ldk], Vt: bf16 [B][dv][ldvt] (V transposed: row = channel), or -- v_rows, dv == 64
// only -- V as rows bf16 [B][Nk][ldvt] exactly as the memory bank stores them.  O: bf16 [B][Nq][ldo] with dv columns.
struct AttnArgs {
  const void* Q = nullptr; long long ldq = 0, q_bstride = 0;
  const void* K = nullptr; long long ldk = 0, k_bstride = 0;
  const void* Vt = nullptr; long long ldvt = 0, vt_bstride = 0;
  int dv = 256, v_rows = 0;
  int B = 1, Nq = 0, Nk = 0;
  float scale = 0.0625f;
  int splits = 1;            // KV splits per query tile (>=1); 0 = balanced ("stream-K") mode
  void* O = nullptr;         // bf16 [B][Nq][ldo]
  long long ldo = 0, o_bstride = 0;
  float* part_o = nullptr;   // workspace when splits > 1: f32 [B][splits][Nq][dv]
  float* part_ml = nullptr;  // f32 [B][splits][Nq][2] (running max in log2 domain, sum)
};
extern int g_attn_cluster;
extern int g_attn_balanced;
extern int g_attn_bal_min_tiles;
extern int g_attn_v_rows;
extern long long* g_attn_trace;
size_t attn_workspace_bytes(int B, int Nq, int splits, int dv = 256);
size_t attn_part_ml_offset(int B, int Nq, int splits, int dv = 256);
int attn_pick_splits(int B, int Nq, int Nk);
// picker for a given value operand: the two-query-tile kernel (attn_x2.cu) serves dv == 64 with V as rows when g_attn_x2
int attn_pick_splits_for(int B, int Nq, int Nk, int dv, int v_rows);
int launch_attention(const AttnArgs& a, cudaStream_t stream);
// memory cross-attention with two query tiles per CTA (attn_x2.cu): dv == 64, v_rows, 1..16 fixed KV splits
extern int g_attn_x2;
extern int g_attn_x2_poly;
int attn_x2_pick_splits(int B, int Nq, int Nk);
int launch_attention_x2(const AttnArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------- fused FFN (ffn_fused.cu)
// x[b][m][:] += act(t[b][m][:] W1^T + b1) W2^T + b2 in one cluster kernel (hidden activations stay in TMEM);
// (ff, gelu) = (2048, 0): memory-attention FFN, (1024, 1): CXBlock point-wise pair.
int launch_ffn_fused(const void* t, long long ldt, long long t_bstride, const void* w1, const float* b1, const void* w2,
                     const float* b2, float* x, long long x_bstride, int B, int M, cudaStream_t stream, int ff = 2048,
                     int gelu = 0);
extern int g_ffn_fused;
extern int g_tail_fused;
extern int g_tail_quarter;
extern long long* g_ffn_trace;
// x_mid = x_in + ao W0^T + b0;  x_out = x_mid + FFN(LN(x_mid));  t_out = LN2(x_out)   -- one cluster kernel
struct LayerTailArgs {
  const void* ao = nullptr;      // bf16 [B][M][64]
  const void* w0 = nullptr;      // bf16 [256][64]
  const float* b0 = nullptr;
  const float* ln_w = nullptr; const float* ln_b = nullptr; float ln_eps = 1e-5f;
  const void* w1 = nullptr; const float* b1 = nullptr;   // [2048][256]
  const void* w2 = nullptr; const float* b2 = nullptr;   // [256][2048]
  const float* x_in = nullptr; float* x_out = nullptr;   // f32 [B][M][256], different buffers
  const float* ln2_w = nullptr; const float* ln2_b = nullptr; float ln2_eps = 1e-5f;
  void* t_out = nullptr; int t_out_bf16 = 1; long long t_out_st = 256, t_out_sb = 0;   // (b,row,c) at b*sb + row*st + c
  int B = 1, M = 0;
};
int launch_layer_tail(const LayerTailArgs& a, cudaStream_t stream);

// x += ao Wo^T + bo;  q = RoPE(LayerNorm(x) Wq^T + bq)  -- the middle of a memory-attention layer in one cluster kernel
// (mid_fused.cu): self-attention output projection + residual, LayerNorm2, cross-attention query projection + axial RoPE
struct MidArgs {
  const void* ao = nullptr;                                  // bf16 [B][M][256]
  const void* wo = nullptr; const float* bo = nullptr;       // [256][256], [256]
  const float* ln_w = nullptr; const float* ln_b = nullptr; float ln_eps = 1e-5f;
  const void* wq = nullptr; const float* bq = nullptr;
  float* x = nullptr;                                        // f32 [B][M][256], in place
  const float* rope_cos = nullptr; const float* rope_sin = nullptr; int rope_period = 0;
  void* q = nullptr; long long ldq = 256, q_bstride = 0;      // bf16 rows
  int B = 1, M = 0;
};
int launch_mid_fused(const MidArgs& a, cudaStream_t stream);
extern int g_mid_fused;
extern int g_mem_attn_head_short, g_mem_attn_keys0_inline, g_mem_attn_keys_ahead_all;   // modules.cu

// ---------------------------------------------------------------- connected components (cc.cu)
size_t cc_workspace_bytes(int n, int h, int w, bool fill);
int launch_cc_label(const uint8_t* img, int n, int h, int w, int32_t* labels, int32_t* counts, void* ws,
                    size_t ws_bytes, cudaStream_t stream);
int launch_fill_holes(float* scores, int n, int h, int w, int max_area, float fill_value, void* ws, size_t ws_bytes,
                      cudaStream_t stream);

// ---------------------------------------------------------------- row / layout kernels (rowops.cu)
// LayerNorm over 256 channels of f32 rows [B][T][256]; writes f32 and/or bf16 at b*sb + t*st.
int launch_ln256(const float* x, int B, int T, const float* w, const float* b, float eps, int gelu, float* out_f32,
                 long long f_sb, long long f_st, void* out_bf16, long long h_sb, long long h_st, cudaStream_t stream);
// out[b][t][c] = a(t,b,c) + alpha * p(t,b,c); inputs f32/bf16 at t*st + b*sb + c; outputs contiguous rows.
int launch_axpy_rows(const void* a, int a_bf16, long long a_st, long long a_sb, const void* p, int p_bf16,
                     long long p_st, long long p_sb, float alpha, int B, int T, int C, float* out_f32, void* out_bf16,
                     cudaStream_t stream);
// ... with the output rows a slice of a taller [B][rows][C] matrix (batch stride out_sb elements)
int launch_axpy_rows_strided(const void* a, int a_bf16, long long a_st, long long a_sb, const void* p, int p_bf16,
                             long long p_st, long long p_sb, float alpha, int B, int T, int C, float* out_f32, void* out_bf16,
                             long long out_sb, cudaStream_t stream);
// NCHW view (strides sb,sc,sh,sw; zeros allowed) (+ addend) -> rows [B][H*W][C] f32 and/or bf16.
int launch_nchw_to_rows(const void* in, int in_bf16, const long long si[4], const void* add, int add_bf16,
                        const long long sa[4], int B, int C, int H, int W, float* out_f32, void* out_bf16,
                        cudaStream_t stream);
// rows [B][T][C] f32 -> NCHW [B][C][T] (+ gate[b]*vec[c]) f32 and/or bf16.
int launch_rows_to_nchw(const float* in, int B, int C, int T, const float* gate, const float* vec, float* out_f32,
                        void* out_bf16, cudaStream_t stream);
struct SmallLinArgs {
  const float* x = nullptr; long long x_sg = 0, x_sr = 0;
  const float* xadd = nullptr; long long xa_sg = 0, xa_sr = 0;
  const void* W = nullptr; long long w_sg = 0;          // bf16 [G][N][K]
  const float* bias = nullptr; long long b_sg = 0;
  const float* res = nullptr; long long r_sg = 0, r_sr = 0;
  float* out = nullptr; long long o_sg = 0, o_sr = 0;
  int G = 1, R = 0, N = 0, K = 0, act = 0;              // act: 0 none, 1 relu, 3 sigmoid
};
int launch_small_linear(const SmallLinArgs& a, cudaStream_t stream);
int launch_small_linear_multi(const SmallLinArgs* a, int count, cudaStream_t stream);  // up to 4 independent problems, one launch
int launch_ln256_small(const float* x, long long x_sr, int rows, const float* w, const float* b, float eps, float* out,
                       long long o_sr, cudaStream_t stream);

int launch_ln64_gelu(const float* x, long long rows, const float* w, const float* b, float eps, void* out, cudaStream_t stream);
extern int g_mds3_tc;
extern int g_dwconv_small; // 1 = small batches take the one-CTA-per-8-pixels kernel instead of the TMA strip kernel
extern int g_dwconv_tma;   // depth-wise 7x7 strip kernel: 1 = input rows staged by TMA, 0 = the r1 kernel (global loads)
int launch_build_tokens(const float* out_tokens, int n_out, const float* sparse, int Ns, int B, float* tok_a, float* tok_b,
                        cudaStream_t stream);
int launch_rows_gate_cast(const float* in, int B, int T, int C, const float* gate, const float* vec, void* out,
                          cudaStream_t stream);
int launch_bank_shift(void* bank, int B, int HW, int n_mem, int n_ptr, int k, const void* new_rows, const float* new_ptr,
                      cudaStream_t stream);
int launch_transpose_rows64(const void* in, int B, int T, void* out, long long ld, cudaStream_t stream);
int launch_multi_copy(const void* const* src, void* const* dst, const size_t* bytes, int n, cudaStream_t stream);
int launch_gather_rows(const float* src, long long sg, long long sr, int G, int R, int n, float* dst, cudaStream_t stream);

// ---------------------------------------------------------------- mask decoder kernels (decoder.cu)
int launch_tok_self_attn(const float* q, const float* k, const float* v, int B, int Nt, float* out, cudaStream_t stream);
int launch_t2i_attn(const float* q, const void* kv, long long ld, long long kv_sb, int koff, int voff, int B, int Nt,
                    int T, float* out, cudaStream_t stream);
// planes = 1: q is stored head-major, [qoff + h][T][16] per batch element (ld unused)
int launch_i2t_attn(const void* qrows, long long ld, long long q_sb, int qoff, const float* ktok, const float* vtok, int B,
                    int Nt, int T, void* out, cudaStream_t stream, int planes = 0);
int launch_up1_post(const void* g, const void* feat, int feat_bf16, long long feat_sb, int B, int h, int w,
                    const float* lnw, const float* lnb, float eps, void* out, cudaStream_t stream);
int launch_up2_masks(const void* u, const float* w2t, const float* bias, const void* feat, int feat_bf16,
                     long long feat_sb, const float* hyper, int B, int M, int h2, int w2, float* masks,
                     cudaStream_t stream);
int launch_up2_masks_tc(const void* u, const void* wh, const float* bias, const float* feat, long long feat_sb,
                        const float* hyper, int B, int M, int h2, int w2, float* masks, cudaStream_t stream);
extern int g_up2_tc;
int launch_select_best(const float* masks, const float* iou, const float* tokens, const float* obj_logits, int B, int M,
                       int multimask, int HW, float* low_res, float* tok_sel, int* best_idx, float* is_obj,
                       float* occluded, cudaStream_t stream);
int launch_gate_ptr(float* ptr, const float* is_obj, const float* no_obj_ptr, int B, cudaStream_t stream);

// ---------------------------------------------------------------- mask decoder, token side of a layer (dec_tok.cu)
// One cluster kernel (8 CTAs = 8 heads per batch element) runs any combination of the three token-side stages of a
// TwoWayAttentionBlock on the <= 16 token rows: SELF (self-attention + norm1), CROSS (token->image attention over the
// head-major K / V planes + norm2, also the decoder's final attention), MLP (MLP + norm3 + the k / v projections of the
// image->token attention).
enum { DEC_TOK_SELF = 1, DEC_TOK_CROSS = 2, DEC_TOK_MLP = 4, DEC_TOK_FIRST = 8 /* self-attention without PE / residual */ };
struct DecTokArgs {
  int B = 1, Nt = 0, T = 0, flags = 0;
  float eps = 1e-5f;
  float* queries = nullptr;        // f32 [B][Nt][256] in / out
  const float* pe = nullptr;       // f32 [B][Nt][256]
  const vls_attn_w* self_attn = nullptr; const float* n1_w = nullptr; const float* n1_b = nullptr;
  const vls_attn_w* t2i = nullptr; const float* n2_w = nullptr; const float* n2_b = nullptr;
  const void* planes = nullptr; long long planes_bstride = 0; int kplane = 0, vplane = 8;   // bf16 [B][planes][T][16]
  const void* m1_w = nullptr; const float* m1_b = nullptr; const void* m2_w = nullptr; const float* m2_b = nullptr;
  const float* n3_w = nullptr; const float* n3_b = nullptr;
  const vls_attn_w* i2t = nullptr;
  float* kt = nullptr; float* vt = nullptr;   // f32 [B][Nt][128]
};
extern int g_dec_fused;
extern long long* g_dec_trace;
bool dec_tok_supported(int Nt, int T);
int launch_dec_tok(const DecTokArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------- mask decoder, image side of a layer (dec_img.cu)
// image->token attention + output projection + residual + LayerNorm4 (keys f32 in place, bf16 copy) and the NEXT image-side
// projections ([K_t2i | V_t2i | Q_i2t] of the following layer: 384 columns, or the final attention's [K | V]: 256) written
// as head planes -- one 4-CTA cluster kernel per 128 image tokens.
struct DecImgArgs {
  int B = 1, T = 0, Nt = 0;
  const void* planes_in = nullptr; long long planes_in_bstride = 0; int qplane = 16;   // q of head h = plane qplane + h
  const float* kt = nullptr; const float* vt = nullptr;                                  // f32 [B][Nt][128]
  const void* wo = nullptr; const float* bo = nullptr;                                   // i2t out-proj [256][128]
  const float* ln_w = nullptr; const float* ln_b = nullptr; float ln_eps = 1e-5f;
  float* keys = nullptr; void* keys_h = nullptr;                                         // f32 in place / bf16 [B][T][256]
  int n_next = 0;                                                                        // 384, 256 or 0 (no projection)
  const void* wn = nullptr; const float* bn = nullptr; const float* pe_add = nullptr;    // [n_next][256], [n_next], f32 [T][n_next]
  void* planes_out = nullptr; long long planes_out_bstride = 0;
};
extern int g_dec_img_fused;
bool dec_img_supported(int Nt, int T);
int launch_dec_img(const DecImgArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------- memory encoder kernels (memenc.cu)
int launch_mds1(const float* src, int mode, int B, int H, int W, int factor, float scale, float bias_v, const float* wgt,
                const float* cb, const float* lnw, const float* lnb, float eps, void* out, cudaStream_t stream);
int launch_mds2(const void* in, int B, int H, int W, const float* wgt, const float* cb, const float* lnw, const float* lnb,
                float eps, void* out, cudaStream_t stream);
int launch_mds3(const void* in, int B, int H, int W, const float* wgt, const float* cb, const float* lnw, const float* lnb,
                float eps, void* out, cudaStream_t stream);
int launch_im2col3x3s2(const void* in, int B, int H, int W, int C, void* out, cudaStream_t stream);
int launch_dwconv7_ln(const float* x, int B, int H, int W, const float* wgt, const float* cb, const float* lnw,
                      const float* lnb, float eps, void* out, cudaStream_t stream);

// ---------------------------------------------------------------- resize (resize.cu)
int launch_resize_bilinear(const float* in, int n, int h, int w, float* out, int H, int W, cudaStream_t stream);
int launch_resize_binarize(const float* in, int n, int h, int w, int H, int W, float thresh, uint8_t* out_u8,
                           uint8_t* out_bits, cudaStream_t stream);

}  // namespace vls
