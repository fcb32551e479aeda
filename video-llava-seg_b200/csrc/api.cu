// extern "C" entry points of libvls_b200.so (declared in include/vls_b200.h).
#include "vls_b200.h"

#include <string>

#include "kernels.h"

using namespace vls;

extern "C" {

const char* vls_last_error(void) { return last_error(); }
int vls_abi_version(void) { return 2; }
long long vls_launch_count(void) { return launch_count(); }
void vls_launch_count_add(long long n) { count_launches((int)n); }
void vls_attention_trace(long long* device_buffer) { g_attn_trace = device_buffer; }
void vls_ffn_trace(long long* device_buffer) { g_ffn_trace = device_buffer; }
void vls_dec_trace(long long* device_buffer) { g_dec_trace = device_buffer; }
int vls_set_tuning(const char* key, int value) {
  VLS_REQUIRE(key != nullptr, "set_tuning: null key");
  if (std::string(key) == "attn_cluster") {
    VLS_REQUIRE(value == 1 || value == 2, "attn_cluster must be 1 or 2");
    g_attn_cluster = value;
    return 0;
  }
  if (std::string(key) == "attn_balanced") {   // 0: always fixed KV splits, 1: balanced ("stream-K") mode when it helps
    g_attn_balanced = value != 0;
    return 0;
  }
  if (std::string(key) == "attn_x2") {   // memory cross-attention: 1 = two query tiles per CTA (attn_x2.cu), 0 = attn_tc.cu
    g_attn_x2 = value != 0;
    return 0;
  }
  if (std::string(key) == "attn_x2_poly") {   // exponentials of every 4 score pairs computed on the FMA pipe (rest: MUFU)
    VLS_REQUIRE(value >= 0 && value <= 3, "attn_x2_poly must be 0..3");
    g_attn_x2_poly = value;
    return 0;
  }
  if (std::string(key) == "attn_v_rows") {   // memory cross-attention value operand: 1 = bank rows (MN-major), 0 = transposed copy
    g_attn_v_rows = value != 0;
    return 0;
  }
  if (std::string(key) == "ffn_fused") {   // memory-attention FFN: 1 = one cluster kernel (hidden stays in TMEM), 0 = two GEMMs
    g_ffn_fused = value != 0;
    return 0;
  }
  if (std::string(key) == "dec_fused") {   // mask decoder token side: 1 = cluster kernels (dec_tok.cu), 0 = chain of small kernels
    g_dec_fused = value != 0;
    return 0;
  }
  if (std::string(key) == "up2_tc") {   // mask decoder ConvT#2 + hyper product: 1 = tcgen05 GEMM, 0 = FP32-pipe kernel
    g_up2_tc = value != 0;
    return 0;
  }
  if (std::string(key) == "gemm_ring2_above") {   // GEMMs with K <= 256 and more CTAs than this stream operands through a 2-stage ring
    g_gemm_ring2_above = value;
    return 0;
  }
  if (std::string(key) == "gemm_bn64_below") {   // GEMMs with fewer 128x128 tiles than this run with 128x64 tiles
    g_gemm_bn64_below = value;
    return 0;
  }
  if (std::string(key) == "mem_attn_head_short") {   // pipelined frames: head = projections only (1) or through the query projection (0)
    g_mem_attn_head_short = value != 0;
    return 0;
  }
  if (std::string(key) == "mem_attn_keys_ahead_all") {   // pipelined frames: the head projects every layer's known keys (1) or layer 0's (0)
    g_mem_attn_keys_ahead_all = value != 0;
    return 0;
  }
  if (std::string(key) == "mem_attn_keys0_inline") {   // pipelined frames: layer 0's keys on the main stream (1) or on the fork (0)
    g_mem_attn_keys0_inline = value != 0;
    return 0;
  }
  if (std::string(key) == "mid_fused") {   // memory attention: self-attn out-proj + LN2 + cross-attn q-proj as one launch
    g_mid_fused = value != 0;
    return 0;
  }
  if (std::string(key) == "dec_img_fused") {   // mask decoder image side of a layer: 1 = one cluster kernel, 0 = i2t + GEMM + LN + GEMM
    g_dec_img_fused = value != 0;
    return 0;
  }
  if (std::string(key) == "mds3_tc") {   // mask down-sampler stage 3: 1 = im2col + tcgen05 GEMM + LN/GELU, 0 = FP32-pipe kernel
    g_mds3_tc = value != 0;
    return 0;
  }
  if (std::string(key) == "dwconv_small") {   // CXBlock depth-wise 7x7 at small batches: 1 = per-row kernel, 0 = TMA strip kernel
    g_dwconv_small = value != 0;
    return 0;
  }
  if (std::string(key) == "dwconv_tma") {   // CXBlock depth-wise 7x7 strip kernel: 1 = input rows staged by TMA, 0 = global loads
    g_dwconv_tma = value;   // 2: the variant capped at 128 registers (4 CTAs per SM)
    return 0;
  }
  if (std::string(key) == "attn_bal_min_tiles") {   // balanced attention mode only for at least this many 128-key tiles
    g_attn_bal_min_tiles = value;
    return 0;
  }
  if (std::string(key) == "tail_quarter") {   // layer-tail prologue by column quarters (1) or full width in every CTA (0)
    g_tail_quarter = value != 0;
    return 0;
  }
  if (std::string(key) == "tail_fused") {  // memory-attention layer tail (out-proj + LN3 + FFN + next LN) as one launch
    g_tail_fused = value != 0;
    return 0;
  }
  if (std::string(key) == "pdl") {   // programmatic dependent launch on/off (host.h)
    pdl_set(value != 0);
    return 0;
  }
  set_error("set_tuning: unknown key '%s'", key);
  return 1;
}
void vls_prof_enable(int on) { prof_set(on != 0); }
int vls_prof_collect(int slot, int* count, double* total_ms) {
  VLS_REQUIRE(slot >= 0 && slot < PROF_SLOTS && count && total_ms, "prof_collect: bad arguments");
  return prof_collect(slot, count, total_ms);
}

size_t vls_cc_workspace_bytes(int n, int h, int w) { return cc_workspace_bytes(n, h, w, false); }
int vls_cc_label(const uint8_t* img, int n, int h, int w, int32_t* labels, int32_t* counts, void* workspace,
                 size_t workspace_bytes, vls_stream_t stream) {
  return launch_cc_label(img, n, h, w, labels, counts, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t vls_fill_holes_workspace_bytes(int n, int h, int w) { return cc_workspace_bytes(n, h, w, true); }
int vls_fill_holes(float* scores, int n, int h, int w, int max_area, float fill_value, void* workspace,
                   size_t workspace_bytes, vls_stream_t stream) {
  return launch_fill_holes(scores, n, h, w, max_area, fill_value, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vls_gemm_bf16(const vls_gemm_desc* d, vls_stream_t stream) {
  VLS_REQUIRE(d != nullptr, "gemm: null descriptor");
  GemmArgs a;
  a.A = d->A; a.lda = d->lda; a.a_bstride = d->a_bstride;
  a.W = d->W; a.ldw = d->ldw; a.w_bstride = d->w_bstride;
  a.M = d->M; a.N = d->N; a.K = d->K; a.batch = d->batch;
  a.bias = d->bias; a.bias_mode = d->bias_mode; a.act = d->act;
  a.rope_cos = d->rope_cos; a.rope_sin = d->rope_sin; a.rope_period = d->rope_period; a.rope_rows = d->rope_rows;
  a.residual = d->residual; a.ld_res = d->ld_res; a.res_bstride = d->res_bstride;
  a.C = d->C; a.c_bf16 = d->c_bf16; a.ldc = d->ldc; a.c_bstride = d->c_bstride;
  return launch_gemm(a, (cudaStream_t)stream);
}

// splits > 0: fixed KV splits; 0: automatic; -1: force the balanced ("stream-K") mode (tests / tuning)
static int resolve_splits(int B, int Nq, int Nk, int splits, int dv = 256, int v_rows = 0) {
  if (splits == -1) return 0;
  return splits > 0 ? splits : attn_pick_splits_for(B, Nq, Nk, dv, v_rows);
}
size_t vls_attention_workspace_bytes(int B, int Nq, int Nk, int splits) {
  return attn_workspace_bytes(B, Nq, resolve_splits(B, Nq, Nk, splits), 256);
}
size_t vls_attention_qk256_workspace_bytes(int B, int Nq, int Nk, int dv, int splits) {
  // sized for either picker (v_rows is not known here): the two-query-tile kernel may use more splits
  const size_t a = attn_workspace_bytes(B, Nq, resolve_splits(B, Nq, Nk, splits), dv);
  const size_t b = dv == 64 ? attn_workspace_bytes(B, Nq, resolve_splits(B, Nq, Nk, splits, dv, 1), dv) : 0;
  return a > b ? a : b;
}

int vls_attention_qk256(const void* Q, long long ldq, long long q_bstride, const void* K, long long ldk,
                        long long k_bstride, const void* V, long long ldv, long long v_bstride, int dv, int v_rows, int B,
                        int Nq, int Nk, float scale, int splits, void* O, long long ldo, long long o_bstride,
                        void* workspace, size_t workspace_bytes, vls_stream_t stream) {
  splits = resolve_splits(B, Nq, Nk, splits, dv, v_rows);
  AttnArgs a;
  a.Q = Q; a.ldq = ldq; a.q_bstride = q_bstride;
  a.K = K; a.ldk = ldk; a.k_bstride = k_bstride;
  a.Vt = V; a.ldvt = ldv; a.vt_bstride = v_bstride; a.dv = dv; a.v_rows = v_rows;
  a.B = B; a.Nq = Nq; a.Nk = Nk; a.scale = scale; a.splits = splits;
  a.O = O; a.ldo = ldo; a.o_bstride = o_bstride;
  if (splits != 1) {   // fixed KV splits, or 0 = balanced mode (picked automatically)
    VLS_REQUIRE(dv == 256 || dv == 64, "attention: value dimension must be 256 or 64 (got %d)", dv);
    const size_t need = attn_workspace_bytes(B, Nq, splits, dv);
    VLS_REQUIRE(workspace && workspace_bytes >= need, "attention: workspace too small (%zu < %zu)", workspace_bytes,
                need);
    a.part_o = reinterpret_cast<float*>(workspace);
    a.part_ml = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + attn_part_ml_offset(B, Nq, splits, dv));
  }
  return launch_attention(a, (cudaStream_t)stream);
}

int vls_attention_d256(const void* Q, long long ldq, long long q_bstride, const void* K, long long ldk,
                       long long k_bstride, const void* Vt, long long ldvt, long long vt_bstride, int B, int Nq, int Nk,
                       float scale, int splits, void* O, long long ldo, long long o_bstride, void* workspace,
                       size_t workspace_bytes, vls_stream_t stream) {
  return vls_attention_qk256(Q, ldq, q_bstride, K, ldk, k_bstride, Vt, ldvt, vt_bstride, 256, 0, B, Nq, Nk, scale, splits, O,
                             ldo, o_bstride, workspace, workspace_bytes, stream);
}

int vls_ffn_fused(const void* t_bf16, long long ldt, long long t_bstride, const void* w1_bf16, const float* b1,
                  const void* w2_bf16, const float* b2, float* x, long long x_bstride, int B, int M, vls_stream_t stream) {
  return launch_ffn_fused(t_bf16, ldt, t_bstride, w1_bf16, b1, w2_bf16, b2, x, x_bstride, B, M, (cudaStream_t)stream);
}

int vls_mem_attn_layer_tail(const void* ao_bf16, const void* w0_bf16, const float* b0, const float* ln_w, const float* ln_b,
                            float ln_eps, const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2,
                            const float* x_in, float* x_out, const float* ln2_w, const float* ln2_b, float ln2_eps,
                            void* t_out, int t_out_dtype, long long t_out_st, long long t_out_sb, int B, int M,
                            vls_stream_t stream) {
  LayerTailArgs a;
  a.ao = ao_bf16; a.w0 = w0_bf16; a.b0 = b0; a.ln_w = ln_w; a.ln_b = ln_b; a.ln_eps = ln_eps;
  a.w1 = w1_bf16; a.b1 = b1; a.w2 = w2_bf16; a.b2 = b2; a.x_in = x_in; a.x_out = x_out;
  a.ln2_w = ln2_w; a.ln2_b = ln2_b; a.ln2_eps = ln2_eps;
  a.t_out = t_out; a.t_out_bf16 = t_out_dtype == VLS_BF16; a.t_out_st = t_out_st; a.t_out_sb = t_out_sb;
  a.B = B; a.M = M;
  return launch_layer_tail(a, (cudaStream_t)stream);
}

int vls_bank_shift(void* bank, int B, int HW, int n_mem, int n_ptr, int tokens_per_ptr, const void* new_rows,
                   const float* new_ptr, vls_stream_t stream) {
  return launch_bank_shift(bank, B, HW, n_mem, n_ptr, tokens_per_ptr, new_rows, new_ptr, (cudaStream_t)stream);
}
int vls_multi_copy(const void* const* src, void* const* dst, const size_t* bytes, int n, vls_stream_t stream) {
  VLS_REQUIRE(n == 0 || (src && dst && bytes), "multi_copy: null argument");
  return launch_multi_copy(src, dst, bytes, n, (cudaStream_t)stream);
}

int vls_resize_binarize(const float* in, int n, int h, int w, int H, int W, float thresh, uint8_t* out_u8, uint8_t* out_bits,
                        vls_stream_t stream) {
  VLS_REQUIRE(n == 0 || in, "resize_binarize: null pointer");
  return launch_resize_binarize(in, n, h, w, H, W, thresh, out_u8, out_bits, (cudaStream_t)stream);
}

int vls_resize_bilinear(const float* in, int n, int h, int w, float* out, int H, int W, vls_stream_t stream) {
  VLS_REQUIRE(n == 0 || (in && out), "resize: null pointer");
  return launch_resize_bilinear(in, n, h, w, out, H, W, (cudaStream_t)stream);
}

int vls_linear_f32(const float* x, long long ldx, const void* w_bf16, const float* bias, int rows, int n, int k, int act,
                   float* out, long long ldo, vls_stream_t stream) {
  VLS_REQUIRE(x && w_bf16 && out, "linear: null pointer");
  SmallLinArgs s;
  s.x = x; s.x_sr = ldx; s.W = w_bf16; s.bias = bias; s.out = out; s.o_sr = ldo;
  s.G = 1; s.R = rows; s.N = n; s.K = k; s.act = act;
  return launch_small_linear(s, (cudaStream_t)stream);
}

int vls_axpy_rows(const void* a, int a_dtype, long long a_st, long long a_sb, const void* p, int p_dtype, long long p_st,
                  long long p_sb, float alpha, int B, int T, int C, void* out, int out_dtype, vls_stream_t stream) {
  VLS_REQUIRE(a && out, "axpy_rows: null pointer");
  return launch_axpy_rows(a, a_dtype, a_st, a_sb, p, p_dtype, p_st, p_sb, alpha, B, T, C,
                          out_dtype == VLS_F32 ? (float*)out : nullptr, out_dtype == VLS_BF16 ? out : nullptr,
                          (cudaStream_t)stream);
}

int vls_dwconv7_ln(const float* x, int B, int H, int W, const float* dw_w, const float* dw_b, const float* ln_w,
                   const float* ln_b, float eps, void* out_bf16, vls_stream_t stream) {
  VLS_REQUIRE(x && dw_w && dw_b && ln_w && ln_b && out_bf16, "dwconv7_ln: null pointer");
  return launch_dwconv7_ln(x, B, H, W, dw_w, dw_b, ln_w, ln_b, eps, out_bf16, (cudaStream_t)stream);
}

int vls_layernorm256(const float* x, long long rows, const float* w, const float* b, float eps, int gelu, void* out_bf16,
                     vls_stream_t stream) {
  VLS_REQUIRE(x && w && b && out_bf16 && rows >= 0 && rows < (1ll << 31), "layernorm256: bad arguments");
  return launch_ln256(x, 1, (int)rows, w, b, eps, gelu, nullptr, 0, 0, out_bf16, 0, 256, (cudaStream_t)stream);
}

}  // extern "C"
