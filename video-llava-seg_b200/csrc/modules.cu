// Module-level entry points: each one sequences the kernels of one reference module on the caller's
// stream using a caller-provided workspace (no allocation, no synchronisation).
#include <cstdlib>

#include "vls_b200.h"

#include "kernels.h"

using namespace vls;

namespace {

constexpr int C = 256;      // d_model / transformer_dim
constexpr int CM = 64;      // mem_dim
constexpr int FFN = 2048;
constexpr float LN_EPS = 1e-5f;   // nn.LayerNorm default
constexpr float LN2D_EPS = 1e-6f; // LayerNorm2d (sam2_utils.py:141-153)

inline long long rup(long long x, long long m) { return (x + m - 1) / m * m; }

GemmArgs lin(const void* A, long long lda, long long a_bs, const void* W, int M, int N, int K, int batch, const float* bias,
             void* Cout, int c_bf16, long long ldc, long long c_bs) {
  GemmArgs g;
  g.A = A; g.lda = lda; g.a_bstride = a_bs;
  g.W = W; g.ldw = K; g.w_bstride = 0;
  g.M = M; g.N = N; g.K = K; g.batch = batch;
  g.bias = bias; g.bias_mode = bias ? 1 : 0;
  g.C = Cout; g.c_bf16 = c_bf16; g.ldc = ldc; g.c_bstride = c_bs;
  return g;
}

}  // namespace

extern "C" {

// ================================================================== memory attention
static size_t mem_attn_ws(int B, int Nq, int Nk, int L) {
  const long long ldv = rup(Nk, 64), ldvs = rup(Nq, 64);
  size_t n = 0;
  n += 2 * align256((size_t)B * Nq * C * 4);        // x (two buffers: the fused layer tail reads one and writes the other)
  n += align256((size_t)B * Nq * C * 2);            // t
  n += align256((size_t)B * Nq * 2 * C * 2);        // qk
  n += 2 * align256((size_t)B * Nk * CM * 2);       // mem, mempos
  n += align256((size_t)L * B * Nk * C * 2);        // rotated cross-attention keys of every layer
  n += align256((size_t)B * CM * ldv * 2);          // transposed memory (only used with attn_v_rows = 0)
  n += align256((size_t)B * C * ldvs * 2);          // transposed self-attention values
  n += align256((size_t)B * Nq * C * 2);            // ao
  n += align256((size_t)B * Nq * FFN * 2);          // h
  const int s1 = attn_pick_splits(B, Nq, Nq), s2 = attn_pick_splits(B, Nq, Nk), s3 = attn_x2_pick_splits(B, Nq, Nk);
  const size_t a1 = attn_workspace_bytes(B, Nq, s1, C), a2 = attn_workspace_bytes(B, Nq, s2, CM);
  const size_t a3 = attn_workspace_bytes(B, Nq, s3, CM);   // either cross-attention kernel may be selected at run time
  n += align256(a1 > a2 ? (a1 > a3 ? a1 : a3) : (a2 > a3 ? a2 : a3));
  return n + 4096;
}

size_t vls_mem_attn_workspace_bytes(int B, int Nq, int Nk) { return mem_attn_ws(B, Nq, Nk, 8); }

}  // extern "C"
// vls_set_tuning("mem_attn_head_short"): 1 = the head (phase 1) ends after layer 0's q/k/v projections, and the self-attention
// + out-projection / LayerNorm2 / query projection open the rest next to the key projection; 0 = the head runs through the
// cross-attention query projection (default).  Measured (A/B on one box, frame ms): no pipelining 0.911, full head 0.894, short
// head 0.920 -- with the short head the self-attention (128 CTAs, high-priority chain) opens the frame and the key projection
// on its fork only gets SMs ~17 us later, so the first cross-attention starts 54 us into the frame instead of 33.
namespace vls { int g_mem_attn_head_short = 0; int g_mem_attn_keys0_inline = 1; int g_mem_attn_keys_ahead_all = 0; }
extern "C" {
// phase 0: the whole stack.  phase 1 (HEAD): only what depends on `curr` alone -- x = curr + 0.1 pos, layer 0's LayerNorm1,
// q/k/v projections, self-attention, out-projection + LayerNorm2 + cross-attention query projection -- leaving x and the
// rotated queries in the workspace.  phase 2 (REST): everything else, starting at layer 0's memory K projection and
// cross-attention, on a workspace whose head has been run.  The graph path runs frame t+1's head next to frame t's mask
// decoder and memory encoder (graphed.py), so a frame starts at its first cross-attention.
static int mem_attn_forward_impl(const vls_mem_attn_weights* w, const void* curr, int curr_dtype, long long curr_st,
                         long long curr_sb, const void* curr_pos, int pos_dtype, long long pos_st, long long pos_sb,
                         const void* memory, int mem_dtype, long long mem_st, long long mem_sb, const void* memory_pos,
                         int mpos_dtype, long long mpos_st, long long mpos_sb, int B, int Nq, int Nk,
                         int num_obj_ptr_tokens, void* out, int out_dtype, long long out_st, long long out_sb,
                         void* workspace, size_t workspace_bytes, vls_stream_t stream_, int phase, int ahead_rows,
                         int ahead_shift_from, int ahead_shift) {
  cudaStream_t st = (cudaStream_t)stream_;
  VLS_REQUIRE(phase >= 0 && phase <= 4, "mem_attn: phase must be 0 (whole), 1 (head), 2 (rest), 3 (head front) or 4 (head back)");
  // phase 3 + phase 4 = phase 1: front = x and layer 0's q/k/v projections, back = self-attention ... query projection + the
  // keys projected ahead.  The graph path runs the back half behind the mask decoder (graphed._frame_body).
  const bool is_head = phase == 1 || phase == 3 || phase == 4;
  if (phase == 0) ahead_rows = 0;
  VLS_REQUIRE(ahead_rows >= 0 && ahead_rows <= Nk - num_obj_ptr_tokens && ahead_rows % Nq == 0,
              "mem_attn: keys projected ahead (%d rows) must be whole rotated blocks of Nq = %d keys", ahead_rows, Nq);
  VLS_REQUIRE(ahead_rows == 0 || (ahead_shift_from >= 0 && ahead_shift_from <= ahead_rows && ahead_shift >= 0 &&
                                  ahead_rows + ahead_shift <= Nk),
              "mem_attn: bad shift of the keys projected ahead (from %d by %d of %d)", ahead_shift_from, ahead_shift, ahead_rows);
  VLS_REQUIRE(w && (curr || phase == 2 || phase == 4) && (memory || is_head) && (out || is_head), "mem_attn: null argument");
  VLS_REQUIRE(w->num_layers >= 1 && w->num_layers <= 8, "mem_attn: num_layers out of range");
  VLS_REQUIRE(B >= 1 && Nq >= 1 && Nk >= 1, "mem_attn: bad shape");
  VLS_REQUIRE(num_obj_ptr_tokens >= 0 && num_obj_ptr_tokens <= Nk, "mem_attn: bad num_obj_ptr_tokens");
  VLS_REQUIRE(w->rope_cos && w->rope_sin && w->rope_len == Nq, "mem_attn: RoPE table length %d != Nq %d", w->rope_len, Nq);
  VLS_REQUIRE((Nk - num_obj_ptr_tokens) % Nq == 0, "mem_attn: rotated keys (%d) must be a multiple of Nq (%d)",
              Nk - num_obj_ptr_tokens, Nq);  // rope_k_repeat (sam/transformer.py:329-338)
  const int L = w->num_layers;
  VLS_REQUIRE(workspace && workspace_bytes >= mem_attn_ws(B, Nq, Nk, L), "mem_attn: workspace too small");
  VLS_REQUIRE(w->ca_k_w_all && w->ca_k_b_all, "mem_attn: stacked K weights missing");
  Workspace ws(workspace, workspace_bytes);
  const long long ldv = rup(Nk, 64), ldvs = rup(Nq, 64);
  float* x = (float*)ws.take((size_t)B * Nq * C * 4);
  float* x_alt = (float*)ws.take((size_t)B * Nq * C * 4);
  void* t = ws.take((size_t)B * Nq * C * 2);
  char* qk = (char*)ws.take((size_t)B * Nq * 2 * C * 2);
  void* mem = ws.take((size_t)B * Nk * CM * 2);
  void* mempos = ws.take((size_t)B * Nk * CM * 2);
  char* kc_all = (char*)ws.take((size_t)L * B * Nk * C * 2);
  void* memT = ws.take((size_t)B * CM * ldv * 2);
  void* vts = ws.take((size_t)B * C * ldvs * 2);
  void* ao = ws.take((size_t)B * Nq * C * 2);
  void* h = ws.take((size_t)B * Nq * FFN * 2);
  const int s_self = attn_pick_splits(B, Nq, Nq), s_cross = attn_pick_splits_for(B, Nq, Nk, CM, g_attn_v_rows);
  const size_t a1 = attn_workspace_bytes(B, Nq, s_self, C), a2 = attn_workspace_bytes(B, Nq, s_cross, CM);
  char* aws = (char*)ws.take(a1 > a2 ? a1 : a2);
  VLS_REQUIRE(x && x_alt && t && qk && mem && mempos && kc_all && memT && vts && ao && h && (aws || (a1 == 0 && a2 == 0)),
              "mem_attn: workspace carve failed");

  // x = curr + 0.1 * curr_pos (memory_attention.py:141); memory -> bf16; memory + pos -> bf16 (:76)
  if (phase != 2 && phase != 4)
    VLS_TRY(launch_axpy_rows(curr, curr_dtype, curr_st, curr_sb, curr_pos, pos_dtype, pos_st, pos_sb, 0.1f, B, Nq, C, x,
                             nullptr, st));
  // The two memory-side conversions feed the K projection and the cross-attention only: they run on the K projection's fork
  // (they were 8.7 us at the head of the main chain).  A bank that already is contiguous bf16 rows [B][Nk][64] -- the device
  // bank of the graph path -- is attended over in place.
  const bool mem_alias = mem_dtype == VLS_BF16 && mem_st == CM && (B == 1 || mem_sb == (long long)Nk * CM) &&
                         (reinterpret_cast<uintptr_t>(memory) & 15) == 0;
  const void* memr = mem_alias ? memory : mem;

  // memory K projections of ALL layers in one launch (they do not depend on x): batch index z = l*B + b.
  //   K_l = RoPE((mem + pos) Wk_l^T + bk_l)  (pointer tokens un-rotated)
  // There is NO value projection: softmax rows sum to one, so softmax(QK^T)(mem Wv^T + bv) = (softmax(QK^T) mem) Wv^T
  // + bv -- the cross-attention kernel attends over the raw 64-d memory rows and Wo.Wv / Wo.bv + bo are folded into the
  // output projection at packing time (memory_attention.py:66-81, sam/transformer.py:311-360).
  // The K launch runs on a forked side stream (event fork/join, capturable into CUDA graphs) so that it overlaps layer
  // 0's LayerNorm -> q/k/v projection -> self-attention -> out-projection chain, whose kernels fill < 1 wave of SMs.
  // One launch per layer with a milestone after each: layer l's cross-attention waits for ITS keys only.  Layer 0's keys are
  // projected on the fork at once, layer l+1's when layer l's cross-attention has finished -- in the background of the layer
  // tail and the next self-attention chain, whose kernels leave SMs idle.  (As one batched launch of L x B problems the first
  // cross-attention waited for all four layers' keys: the 54 us of projection work and layer 0's self-attention chain, which
  // needs the SMs as well, added up to ~104 us before it started; all four launched at the start slowed layer 0's kernels.)
  const bool k_per_layer = L <= 8 || ahead_rows > 0;
  // memory + positional rows [r0, r1) of the bank -> bf16 operand rows; src_shift: the memory rows are read that many rows
  // further on (a device bank that has not been shifted yet: see ahead_rows below)
  auto convert_rows = [&](int r0, int r1, int src_shift, cudaStream_t s_) -> int {
    if (r1 <= r0) return 0;
    const size_t mel = mem_dtype == VLS_BF16 ? 2 : 4, pel = mpos_dtype == VLS_BF16 ? 2 : 4;
    const char* msrc = static_cast<const char*>(memory) + (size_t)(r0 + src_shift) * mem_st * mel;
    const char* psrc = memory_pos ? static_cast<const char*>(memory_pos) + (size_t)r0 * mpos_st * pel : nullptr;
    return launch_axpy_rows_strided(msrc, mem_dtype, mem_st, mem_sb, psrc, mpos_dtype, mpos_st, mpos_sb, memory_pos ? 1.0f : 0.f,
                                    B, r1 - r0, CM, nullptr, static_cast<char*>(mempos) + (size_t)r0 * CM * 2,
                                    (long long)Nk * CM, s_);
  };
  // K_l rows [r0, r1) (r0 a multiple of Nq, so the RoPE phase of a row is its index inside the launch); l < 0: all layers
  auto gemm_keys = [&](int l, int r0, int r1, cudaStream_t s_) -> int {
    if (r1 <= r0) return 0;
    GemmArgs k;
    k.A = static_cast<char*>(mempos) + (size_t)r0 * CM * 2; k.lda = CM; k.a_bstride = (long long)Nk * CM; k.a_batches = B; k.a_div = 1;
    k.ldw = CM; k.w_bstride = (long long)C * CM; k.w_div = B;
    k.M = r1 - r0; k.N = C; k.K = CM;
    k.bias_mode = 1; k.bias_bstride = C; k.bias_div = B;
    const int rot = Nk - num_obj_ptr_tokens;
    k.rope_cos = w->rope_cos; k.rope_sin = w->rope_sin; k.rope_period = Nq; k.rope_rows = (r1 < rot ? r1 : rot) - r0;
    if (k.rope_rows < 0) k.rope_rows = 0;
    k.c_bf16 = 1; k.ldc = C; k.c_bstride = (long long)Nk * C;
    if (l >= 0) {
      k.W = static_cast<const char*>(w->ca_k_w_all) + (size_t)l * C * CM * 2; k.w_batches = 1;
      k.bias = w->ca_k_b_all + (size_t)l * C; k.bias_batches = 1;
      k.C = kc_all + ((size_t)l * B * Nk + r0) * C * 2; k.batch = B;
    } else {
      k.W = w->ca_k_w_all; k.w_batches = L;
      k.bias = w->ca_k_b_all; k.bias_batches = L;
      k.C = kc_all + (size_t)r0 * C * 2; k.batch = L * B;
    }
    return launch_gemm(k, s_);
  };
  // ahead_rows > 0 (phases 1 and 2 only): layer 0's keys of the first ahead_rows bank rows are projected by the HEAD -- one frame
  // ahead, next to the previous frame's mask decoder -- and the rest projects only the rows behind them at the start of the
  // frame (the newest memory and the object pointers: 1/7 of the bank).  With ahead_shift the head reads the memory rows of
  // [ahead_shift_from, ahead_rows) that many rows further on: the caller's bank has not been shifted yet.
  // phase 2 behind a full head: nothing runs on the main stream before layer 0's cross-attention, so its keys are projected
  // there instead of on the fork (the cross-stream edge at the root of a captured graph cost ~8 us before the GEMM started)
  const bool keys0_inline = phase == 2 && !g_mem_attn_head_short && k_per_layer && g_mem_attn_keys0_inline;
  auto project_keys = [&](int l) -> int {     // l < 0: all layers in one launch
    cudaStream_t side = st;
    const bool inl = keys0_inline && l == 0;
    if (!inl) VLS_TRY(fork_begin(0, st, &side));
    // keys projected ahead: layer 0's, or (keys_ahead_all) every layer's -- then no full-bank projection runs in the
    // background of the frame's own attention kernels at all; measured 0.8748 vs 0.8733 ms per frame (A/B on one box): the
    // extra 44 MB the head then writes next to the mask decoder cost more than the quieter attention phase gained, so off
    const int r0 = (phase == 2 && (l == 0 || g_mem_attn_keys_ahead_all)) ? ahead_rows : 0;
    if (l <= 0) {
      if (!mem_alias) VLS_TRY(launch_axpy_rows(memory, mem_dtype, mem_st, mem_sb, nullptr, 0, 0, 0, 0.f, B, Nk, CM, nullptr, mem, side));
      VLS_TRY(convert_rows(r0, Nk, 0, side));
    }
    VLS_TRY(gemm_keys(l, r0, Nk, side));
    if (l <= 0 && !g_attn_v_rows) VLS_TRY(launch_transpose_rows64(memr, B, Nk, memT, ldv, side));
    if (l >= 0 && !inl) VLS_TRY(fork_mark(0, l));
    return 0;
  };
  if (!is_head) VLS_TRY(project_keys(k_per_layer ? 0 : -1));
  bool joined = false;

  // dv = 256: V^T [256][ldvt]; dv = 64: the memory itself, as rows [Nk][64] (v_rows) or transposed [64][ldvt]
  auto attention = [&](const void* K, long long ldk, long long k_bs, const void* V, long long ldvt, long long v_bs, int dv,
                       int v_rows, int nk, int splits) -> int {
    AttnArgs a;
    a.Q = qk; a.ldq = 2 * C; a.q_bstride = (long long)Nq * 2 * C;
    a.K = K; a.ldk = ldk; a.k_bstride = k_bs;
    a.Vt = V; a.ldvt = ldvt; a.vt_bstride = v_bs; a.dv = dv; a.v_rows = v_rows;
    a.B = B; a.Nq = Nq; a.Nk = nk; a.scale = 0.0625f; a.splits = splits;
    a.O = ao; a.ldo = dv; a.o_bstride = (long long)Nq * dv;
    if (splits != 1) {   // fixed KV splits, or 0 = balanced mode
      a.part_o = (float*)aws;
      a.part_ml = (float*)(aws + attn_part_ml_offset(B, Nq, splits, dv));
    }
    return launch_attention(a, st);
  };

  auto keys_ahead = [&]() -> int {   // behind the head, on its stream: nothing in this call waits for them
    if (ahead_rows <= 0) return 0;
    VLS_TRY(convert_rows(0, ahead_shift_from, 0, st));
    VLS_TRY(convert_rows(ahead_shift_from, ahead_rows, ahead_shift, st));
    return gemm_keys(g_mem_attn_keys_ahead_all ? -1 : 0, 0, ahead_rows, st);
  };
  const bool tail_fused = g_ffn_fused && g_tail_fused;
  bool have_t = false, out_done = false;   // tail_fused: the previous layer's tail kernel already produced t = LN1(x) / the output
  for (int l = 0; l < L; ++l) {
    const vls_mem_attn_layer& Lw = w->layers[l];
    // ---- self attention (memory_attention.py:58-64): q = k = v = LN1(x); RoPE on q and k
    const bool first_rest = phase == 2 && l == 0;   // phase 2: layer 0's chain up to the head's end was run ahead
    if (!first_rest && phase != 4) {
    if (!have_t) VLS_TRY(launch_ln256(x, B, Nq, Lw.n1_w, Lw.n1_b, LN_EPS, 0, nullptr, 0, 0, t, (long long)Nq * C, C, st));
    {
      GemmArgs g = lin(t, C, (long long)Nq * C, Lw.sa_qk_w, Nq, 2 * C, C, B, Lw.sa_qk_b, qk, 1, 2 * C, (long long)Nq * 2 * C);
      g.rope_cos = w->rope_cos; g.rope_sin = w->rope_sin; g.rope_period = Nq; g.rope_rows = Nq;
      cudaStream_t vside;          // the value projection runs next to the q/k projection (both read LN1(x))
      VLS_TRY(fork_begin(2, st, &vside));
      GemmArgs v;  // V^T[c][t] = sum_k Wv[c][k] * t[t][k] + bv[c]
      v.A = Lw.sa_v_w; v.lda = C; v.a_bstride = 0;
      v.W = t; v.ldw = C; v.w_bstride = (long long)Nq * C;
      v.M = C; v.N = Nq; v.K = C; v.batch = B;
      v.bias = Lw.sa_v_b; v.bias_mode = 2;
      v.C = vts; v.c_bf16 = 1; v.ldc = ldvs; v.c_bstride = (long long)C * ldvs;
      VLS_TRY(launch_gemm(v, vside));
      VLS_TRY(launch_gemm(g, st));
      VLS_TRY(fork_join(2, st));
    }
    }
    if (phase == 3) return 0;
    if (is_head && g_mem_attn_head_short) return keys_ahead();   // short head: the projections only (kernels of < 10 us)
    if (!(first_rest && !g_mem_attn_head_short)) {
    VLS_TRY(attention(qk + (size_t)C * 2, 2 * C, (long long)Nq * 2 * C, vts, ldvs, (long long)C * ldvs, C, 0, Nq, s_self));
    if (g_mid_fused) {
      // ---- self-attention output projection + residual, LayerNorm2 and the cross-attention query projection (+ RoPE) in
      //      ONE cluster kernel (mid_fused.cu)
      MidArgs a;
      a.ao = ao; a.wo = Lw.sa_o_w; a.bo = Lw.sa_o_b; a.ln_w = Lw.n2_w; a.ln_b = Lw.n2_b; a.ln_eps = LN_EPS;
      a.wq = Lw.ca_q_w; a.bq = Lw.ca_q_b; a.x = x; a.rope_cos = w->rope_cos; a.rope_sin = w->rope_sin; a.rope_period = Nq;
      a.q = qk; a.ldq = 2 * C; a.q_bstride = (long long)Nq * 2 * C; a.B = B; a.M = Nq;
      VLS_TRY(launch_mid_fused(a, st));
    } else {
    {
      GemmArgs g = lin(ao, C, (long long)Nq * C, Lw.sa_o_w, Nq, C, C, B, Lw.sa_o_b, x, 0, C, (long long)Nq * C);
      g.residual = x; g.ld_res = C; g.res_bstride = (long long)Nq * C;
      VLS_TRY(launch_gemm(g, st));
    }
    // ---- cross attention to the memory bank (memory_attention.py:66-81)
    VLS_TRY(launch_ln256(x, B, Nq, Lw.n2_w, Lw.n2_b, LN_EPS, 0, nullptr, 0, 0, t, (long long)Nq * C, C, st));
    {
      GemmArgs g = lin(t, C, (long long)Nq * C, Lw.ca_q_w, Nq, C, C, B, Lw.ca_q_b, qk, 1, 2 * C, (long long)Nq * 2 * C);
      g.rope_cos = w->rope_cos; g.rope_sin = w->rope_sin; g.rope_period = Nq; g.rope_rows = Nq;
      VLS_TRY(launch_gemm(g, st));
    }
    }
    }
    if (is_head) return keys_ahead();
    if (k_per_layer) {
      if (!(keys0_inline && l == 0)) VLS_TRY(fork_wait(0, l, st));
    } else if (!joined) {
      VLS_TRY(fork_join(0, st));
      joined = true;
    }
    if (g_attn_v_rows)
      VLS_TRY(attention(kc_all + (size_t)l * B * Nk * C * 2, C, (long long)Nk * C, memr, CM, (long long)Nk * CM, CM, 1, Nk, s_cross));
    else
      VLS_TRY(attention(kc_all + (size_t)l * B * Nk * C * 2, C, (long long)Nk * C, memT, ldv, (long long)CM * ldv, CM, 0, Nk, s_cross));
    if (k_per_layer && l + 1 < L) VLS_TRY(project_keys(l + 1));
    if (tail_fused) {
      // ---- the rest of the layer in ONE cluster kernel (ffn_fused.cu): folded out-projection + residual, LayerNorm3, FFN
      //      + residual, and the LayerNorm that follows (next layer's norm1, or the final norm straight into `out`)
      LayerTailArgs a;
      a.ao = ao; a.w0 = Lw.ca_ov_w; a.b0 = Lw.ca_ov_b;
      a.ln_w = Lw.n3_w; a.ln_b = Lw.n3_b; a.ln_eps = LN_EPS;
      a.w1 = Lw.l1_w; a.b1 = Lw.l1_b; a.w2 = Lw.l2_w; a.b2 = Lw.l2_b;
      a.x_in = x; a.x_out = x_alt; a.B = B; a.M = Nq; a.ln2_eps = LN_EPS;
      if (l + 1 < L) {
        a.ln2_w = w->layers[l + 1].n1_w; a.ln2_b = w->layers[l + 1].n1_b;
        a.t_out = t; a.t_out_bf16 = 1; a.t_out_st = C; a.t_out_sb = (long long)Nq * C;
      } else {
        a.ln2_w = w->norm_w; a.ln2_b = w->norm_b;
        a.t_out = out; a.t_out_bf16 = out_dtype == VLS_BF16; a.t_out_st = out_st; a.t_out_sb = out_sb;
        out_done = true;
      }
      VLS_TRY(launch_layer_tail(a, st));
      float* tmp = x; x = x_alt; x_alt = tmp;
      have_t = true;
      continue;
    }
    {   // x += (P mem) (Wo Wv)^T + (Wo bv + bo): the folded value/output projection, K = 64
      GemmArgs g = lin(ao, CM, (long long)Nq * CM, Lw.ca_ov_w, Nq, C, CM, B, Lw.ca_ov_b, x, 0, C, (long long)Nq * C);
      g.residual = x; g.ld_res = C; g.res_bstride = (long long)Nq * C;
      VLS_TRY(launch_gemm(g, st));
    }
    // ---- FFN (memory_attention.py:95-98)
    VLS_TRY(launch_ln256(x, B, Nq, Lw.n3_w, Lw.n3_b, LN_EPS, 0, nullptr, 0, 0, t, (long long)Nq * C, C, st));
    if (g_ffn_fused) {   // one cluster kernel: the [Nq][2048] hidden tensor stays in tensor memory (ffn_fused.cu)
      VLS_TRY(launch_ffn_fused(t, C, (long long)Nq * C, Lw.l1_w, Lw.l1_b, Lw.l2_w, Lw.l2_b, x, (long long)Nq * C, B, Nq, st));
    } else {
      GemmArgs g = lin(t, C, (long long)Nq * C, Lw.l1_w, Nq, FFN, C, B, Lw.l1_b, h, 1, FFN, (long long)Nq * FFN);
      g.act = 1;
      VLS_TRY(launch_gemm(g, st));
      GemmArgs g2 = lin(h, FFN, (long long)Nq * FFN, Lw.l2_w, Nq, C, FFN, B, Lw.l2_b, x, 0, C, (long long)Nq * C);
      g2.residual = x; g2.ld_res = C; g2.res_bstride = (long long)Nq * C;
      VLS_TRY(launch_gemm(g2, st));
    }
  }
  if (out_done) return 0;
  if (out_dtype == VLS_BF16)
    return launch_ln256(x, B, Nq, w->norm_w, w->norm_b, LN_EPS, 0, nullptr, 0, 0, out, out_sb, out_st, st);
  return launch_ln256(x, B, Nq, w->norm_w, w->norm_b, LN_EPS, 0, (float*)out, out_sb, out_st, nullptr, 0, 0, st);
}

int vls_mem_attn_forward(const vls_mem_attn_weights* w, const void* curr, int curr_dtype, long long curr_st,
                         long long curr_sb, const void* curr_pos, int pos_dtype, long long pos_st, long long pos_sb,
                         const void* memory, int mem_dtype, long long mem_st, long long mem_sb, const void* memory_pos,
                         int mpos_dtype, long long mpos_st, long long mpos_sb, int B, int Nq, int Nk,
                         int num_obj_ptr_tokens, void* out, int out_dtype, long long out_st, long long out_sb,
                         void* workspace, size_t workspace_bytes, vls_stream_t stream_) {
  return mem_attn_forward_impl(w, curr, curr_dtype, curr_st, curr_sb, curr_pos, pos_dtype, pos_st, pos_sb, memory, mem_dtype,
                               mem_st, mem_sb, memory_pos, mpos_dtype, mpos_st, mpos_sb, B, Nq, Nk, num_obj_ptr_tokens, out,
                               out_dtype, out_st, out_sb, workspace, workspace_bytes, stream_, 0, 0, 0, 0);
}

int vls_mem_attn_forward_phase(const vls_mem_attn_weights* w, const void* curr, int curr_dtype, long long curr_st,
                               long long curr_sb, const void* curr_pos, int pos_dtype, long long pos_st, long long pos_sb,
                               const void* memory, int mem_dtype, long long mem_st, long long mem_sb,
                               const void* memory_pos, int mpos_dtype, long long mpos_st, long long mpos_sb, int B, int Nq,
                               int Nk, int num_obj_ptr_tokens, void* out, int out_dtype, long long out_st, long long out_sb,
                               void* workspace, size_t workspace_bytes, vls_stream_t stream_, int phase, int ahead_rows,
                               int ahead_shift_from, int ahead_shift) {
  return mem_attn_forward_impl(w, curr, curr_dtype, curr_st, curr_sb, curr_pos, pos_dtype, pos_st, pos_sb, memory, mem_dtype,
                               mem_st, mem_sb, memory_pos, mpos_dtype, mpos_st, mpos_sb, B, Nq, Nk, num_obj_ptr_tokens, out,
                               out_dtype, out_st, out_sb, workspace, workspace_bytes, stream_, phase, ahead_rows,
                               ahead_shift_from, ahead_shift);
}

// ================================================================== mask decoder
static size_t dec_token_floats(int B, int Nt) {
  // tokens0, queries, q, k, v, a (256 each) + qt/kt/vt/at (128 each) + mlp hidden (2048) + heads scratch
  return (size_t)B * Nt * (6 * 256 + 4 * 128 + 2048) + (size_t)B * (4 * 256 * 2 + 4 * 32 + 2 * 256 * 2) + 1024;
}

size_t vls_mask_decoder_workspace_bytes(int B, int Ns, int H, int W) {
  const size_t T = (size_t)H * W;
  const int Nt = 6 + Ns;
  size_t n = 0;
  n += align256(B * T * C * 4);       // keys f32
  n += align256(B * T * C * 2);       // keys bf16
  n += align256(B * T * 384 * 2);     // fused image-side projections
  n += align256(B * T * 128 * 2);     // i2t attention output
  n += align256(B * T * C * 4);       // f32 scratch (pre-LN keys / upscaling GEMM)
  n += align256(B * 4 * T * 64 * 2);  // upscaled stage-1 rows
  n += align256(dec_token_floats(B, Nt) * 4);
  return n + 4096;
}

int vls_mask_decoder_forward(const vls_mask_decoder_weights* w, const void* image_embeddings, int emb_dtype,
                             const long long emb_strides[4], const void* dense, int dense_dtype,
                             const long long dense_strides[4], const float* sparse, const void* feat_s0, int s0_dtype,
                             long long s0_bstride, const void* feat_s1, int s1_dtype, long long s1_bstride, int B, int Ns,
                             int H, int W, float* masks, float* iou, float* tokens_out, float* obj_logits,
                             void* workspace, size_t workspace_bytes, vls_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  VLS_REQUIRE(w && image_embeddings && feat_s0 && feat_s1 && masks && iou && tokens_out && obj_logits,
              "mask_decoder: null argument");
  VLS_REQUIRE(B >= 1 && Ns >= 0 && (Ns == 0 || sparse), "mask_decoder: bad prompt arguments");
  VLS_REQUIRE(workspace && workspace_bytes >= vls_mask_decoder_workspace_bytes(B, Ns, H, W),
              "mask_decoder: workspace too small");
  const int T = H * W, Nt = 6 + Ns, R = B * Nt;
  VLS_REQUIRE(Nt <= 32, "mask_decoder: at most 26 sparse prompt tokens are supported (got %d)", Ns);
  Workspace ws(workspace, workspace_bytes);
  float* keys = (float*)ws.take((size_t)B * T * C * 4);
  void* keys_h = ws.take((size_t)B * T * C * 2);
  void* kvq = ws.take((size_t)B * T * 384 * 2);
  void* ai = ws.take((size_t)B * T * 128 * 2);
  float* scratch = (float*)ws.take((size_t)B * T * C * 4);
  void* up1 = ws.take((size_t)B * 4 * T * 64 * 2);
  float* tk = (float*)ws.take(dec_token_floats(B, Nt) * 4);
  VLS_REQUIRE(keys && keys_h && kvq && ai && scratch && up1 && tk, "mask_decoder: workspace carve failed");
  float* tokens0 = tk;               tk += (size_t)R * 256;
  float* queries = tk;               tk += (size_t)R * 256;
  float* q = tk;                     tk += (size_t)R * 256;
  float* k = tk;                     tk += (size_t)R * 256;
  float* v = tk;                     tk += (size_t)R * 256;
  float* a = tk;                     tk += (size_t)R * 256;
  float* qt = tk;                    tk += (size_t)R * 128;
  float* kt = tk;                    tk += (size_t)R * 128;
  float* vt = tk;                    tk += (size_t)R * 128;
  float* at = tk;                    tk += (size_t)R * 128;
  float* hid = tk;                   tk += (size_t)R * 2048;
  float* hy1 = tk;                   tk += (size_t)B * 4 * 256;
  float* hy2 = tk;                   tk += (size_t)B * 4 * 256;
  float* hyper = tk;                 tk += (size_t)B * 4 * 32;
  float* hd1 = tk;                   tk += (size_t)B * 256 * 2;
  float* hd2 = tk;                   tk += (size_t)B * 256 * 2;

  // tokens = [obj_score, iou, mask x4, sparse...] (mask_decoder.py:179-197); queries = tokens
  VLS_TRY(launch_build_tokens(w->out_tokens, 6, sparse, Ns, B, tokens0, queries, st));
  // keys = image_embeddings (+ repeat) + dense, NCHW -> token rows (mask_decoder.py:200-205).  With the cluster kernels the
  // conversion runs on the fork that carries the image-side projection, next to the token chain (tokens -> self-attention).
  const bool keys_on_fork = dec_tok_supported(Nt, T) && dec_img_supported(Nt, T);
  cudaStream_t keys_stream = st;
  if (keys_on_fork) VLS_TRY(fork_begin(1, st, &keys_stream));
  VLS_TRY(launch_nchw_to_rows(image_embeddings, emb_dtype, emb_strides, dense, dense_dtype, dense_strides, B, C, H, W, keys,
                              keys_h, keys_stream));

  auto tok_args = [&](const float* x, const float* xadd, int K, const void* Wt, const float* bias, int N, int act,
                      const float* res, float* out) {
    SmallLinArgs s;
    s.x = x; s.x_sr = K; s.xadd = xadd; s.xa_sr = K;
    s.W = Wt; s.bias = bias; s.res = res; s.r_sr = N; s.out = out; s.o_sr = N;
    s.G = 1; s.R = R; s.N = N; s.K = K; s.act = act;
    return s;
  };
  auto tok_lin = [&](const float* x, const float* xadd, int K, const void* Wt, const float* bias, int N, int act,
                     const float* res, float* out) -> int {
    return launch_small_linear(tok_args(x, xadd, K, Wt, bias, N, act, res, out), st);
  };
  auto t2i = [&](const vls_attn_w& A, const void* rows, long long ld, const float* nw, const float* nb) -> int {
    VLS_TRY(tok_lin(queries, tokens0, 256, A.q_w, A.q_b, 128, 0, nullptr, qt));
    VLS_TRY(launch_t2i_attn(qt, rows, ld, (long long)T * ld, 0, 128, B, Nt, T, at, st));
    VLS_TRY(tok_lin(at, nullptr, 128, A.o_w, A.o_b, 256, 0, queries, a));
    return launch_ln256_small(a, 256, R, nw, nb, LN_EPS, queries, 256, st);
  };

  // Token side of a layer as cluster kernels (dec_tok.cu) when the shape allows it: the image-side projections are then
  // written head-major ("planes" [24 or 16][T][16]) so that the token->image attention reads contiguous keys per head.
  const bool fused = dec_tok_supported(Nt, T);
  auto planes_out = [&](GemmArgs& g) {
    if (!fused) return;
    g.ldc = 16; g.c_colblock = 16; g.c_colblock_stride = (long long)T * 16;
  };
  auto dec_tok = [&](int flags, const vls_dec_layer* L, const vls_attn_w* t2i_w, const float* n2w, const float* n2b,
                     long long planes_bstride, cudaStream_t s_) -> int {
    DecTokArgs d;
    d.B = B; d.Nt = Nt; d.T = T; d.flags = flags; d.eps = LN_EPS;
    d.queries = queries; d.pe = tokens0;
    if (L) {
      d.self_attn = &L->self_attn; d.n1_w = L->n1_w; d.n1_b = L->n1_b;
      d.m1_w = L->mlp1_w; d.m1_b = L->mlp1_b; d.m2_w = L->mlp2_w; d.m2_b = L->mlp2_b; d.n3_w = L->n3_w; d.n3_b = L->n3_b;
      d.i2t = &L->i2t;
    }
    d.t2i = t2i_w; d.n2_w = n2w; d.n2_b = n2b;
    d.planes = kvq; d.planes_bstride = planes_bstride; d.kplane = 0; d.vplane = 8;
    d.kt = kt; d.vt = vt;
    return launch_dec_tok(d, s_);
  };

  const bool img_fused = fused && dec_img_supported(Nt, T);
  if (img_fused) {
    // ---- both layers as cluster kernels: token side (dec_tok.cu) and image side (dec_img.cu).  Only layer 0's image-side
    //      projection is a stand-alone GEMM (next to the token self-attention); every later one is produced by dec_img.
    {
      cudaStream_t side = keys_stream;   // forked above, before the key conversion
      GemmArgs g = lin(keys_h, C, (long long)T * C, w->layers[0].img_w, T, 384, C, B, w->layers[0].img_b, kvq, 1, 384, (long long)T * 384);
      g.residual = w->layers[0].img_pe_add; g.ld_res = 384; g.res_bstride = 0;
      planes_out(g);
      VLS_TRY(launch_gemm(g, side));
      VLS_TRY(dec_tok(DEC_TOK_SELF | DEC_TOK_FIRST, &w->layers[0], nullptr, nullptr, nullptr, 0, st));
      VLS_TRY(fork_join(1, st));
    }
    for (int l = 0; l < 2; ++l) {
      const vls_dec_layer& L = w->layers[l];
      VLS_TRY(dec_tok(DEC_TOK_CROSS | DEC_TOK_MLP, &L, &L.t2i, L.n2_w, L.n2_b, (long long)T * 384, st));
      cudaStream_t side = st;
      if (l == 0) {   // the next layer's token self-attention only needs the token rows: next to the image side
        VLS_TRY(fork_begin(1, st, &side));
        VLS_TRY(dec_tok(DEC_TOK_SELF, &w->layers[1], nullptr, nullptr, nullptr, 0, side));
      }
      DecImgArgs d;
      d.B = B; d.T = T; d.Nt = Nt;
      d.planes_in = kvq; d.planes_in_bstride = (long long)T * 384; d.qplane = 16;
      d.kt = kt; d.vt = vt; d.wo = L.i2t.o_w; d.bo = L.i2t.o_b; d.ln_w = L.n4_w; d.ln_b = L.n4_b; d.ln_eps = LN_EPS;
      d.keys = keys; d.keys_h = keys_h;
      d.planes_out = kvq;   // in place: a cluster reads its rows' q planes before its first barrier and writes after it
      if (l == 0) {
        d.n_next = 384; d.wn = w->layers[1].img_w; d.bn = w->layers[1].img_b; d.pe_add = w->layers[1].img_pe_add;
        d.planes_out_bstride = (long long)T * 384;
      } else {
        d.n_next = 256; d.wn = w->final_img_w; d.bn = w->final_img_b; d.pe_add = w->final_pe_add;
        d.planes_out_bstride = (long long)T * 256;
      }
      VLS_TRY(launch_dec_img(d, st));
      if (l == 0) VLS_TRY(fork_join(1, st));
    }
  }
  for (int l = 0; l < 2 && !img_fused; ++l) {
    const vls_dec_layer& L = w->layers[l];
    // -- token self attention (sam/transformer.py:183-191); layer 0 drops the PE and the residual
    const float* pe = l == 0 ? nullptr : tokens0;
    // -- image-side projections for this layer in one GEMM: [K_t2i | V_t2i | Q_i2t], PE folded in as a residual.
    //    They only depend on the image keys, so they run on the side stream next to the token self-attention chain.
    {
      cudaStream_t side;
      VLS_TRY(fork_begin(1, st, &side));
      GemmArgs g = lin(keys_h, C, (long long)T * C, L.img_w, T, 384, C, B, L.img_b, kvq, 1, 384, (long long)T * 384);
      g.residual = L.img_pe_add; g.ld_res = 384; g.res_bstride = 0;
      planes_out(g);
      VLS_TRY(launch_gemm(g, side));
    }
    if (fused) {
      // self-attention + norm1 next to the image-side GEMM, then token->image attention + MLP + i2t k/v in one launch
      VLS_TRY(dec_tok(DEC_TOK_SELF | (l == 0 ? DEC_TOK_FIRST : 0), &L, nullptr, nullptr, nullptr, 0, st));
      VLS_TRY(fork_join(1, st));
      VLS_TRY(dec_tok(DEC_TOK_CROSS | DEC_TOK_MLP, &L, &L.t2i, L.n2_w, L.n2_b, (long long)T * 384, st));
      VLS_TRY(launch_i2t_attn(kvq, 384, (long long)T * 384, 16, kt, vt, B, Nt, T, ai, st, 1));
    } else {
    {   // q, k, v projections: three independent problems, one launch
      const SmallLinArgs qkv[3] = {tok_args(queries, pe, 256, L.self_attn.q_w, L.self_attn.q_b, 256, 0, nullptr, q),
                                   tok_args(queries, pe, 256, L.self_attn.k_w, L.self_attn.k_b, 256, 0, nullptr, k),
                                   tok_args(queries, nullptr, 256, L.self_attn.v_w, L.self_attn.v_b, 256, 0, nullptr, v)};
      VLS_TRY(launch_small_linear_multi(qkv, 3, st));
    }
    VLS_TRY(launch_tok_self_attn(q, k, v, B, Nt, a, st));
    VLS_TRY(tok_lin(a, nullptr, 256, L.self_attn.o_w, L.self_attn.o_b, 256, 0, l == 0 ? nullptr : queries, q));
    VLS_TRY(launch_ln256_small(q, 256, R, L.n1_w, L.n1_b, LN_EPS, queries, 256, st));
    VLS_TRY(fork_join(1, st));
    // -- tokens -> image cross attention (:193-198)
    VLS_TRY(t2i(L.t2i, kvq, 384, L.n2_w, L.n2_b));
    // -- token MLP (:200-203)
    VLS_TRY(tok_lin(queries, nullptr, 256, L.mlp1_w, L.mlp1_b, 2048, 1, nullptr, hid));
    VLS_TRY(tok_lin(hid, nullptr, 2048, L.mlp2_w, L.mlp2_b, 256, 0, queries, a));
    VLS_TRY(launch_ln256_small(a, 256, R, L.n3_w, L.n3_b, LN_EPS, queries, 256, st));
    // -- image -> tokens cross attention (:205-210)
    {
      const SmallLinArgs kv[2] = {tok_args(queries, tokens0, 256, L.i2t.k_w, L.i2t.k_b, 128, 0, nullptr, kt),
                                  tok_args(queries, nullptr, 256, L.i2t.v_w, L.i2t.v_b, 128, 0, nullptr, vt)};
      VLS_TRY(launch_small_linear_multi(kv, 2, st));
    }
    VLS_TRY(launch_i2t_attn(kvq, 384, (long long)T * 384, 256, kt, vt, B, Nt, T, ai, st));
    }
    {
      GemmArgs g = lin(ai, 128, (long long)T * 128, L.i2t.o_w, T, C, 128, B, L.i2t.o_b, scratch, 0, C, (long long)T * C);
      g.residual = keys; g.ld_res = C; g.res_bstride = (long long)T * C;
      VLS_TRY(launch_gemm(g, st));
    }
    VLS_TRY(launch_ln256(scratch, B, T, L.n4_w, L.n4_b, LN_EPS, 0, keys, (long long)T * C, C, keys_h, (long long)T * C, C, st));
  }
  // -- upscaling (mask_decoder.py:218-225): ConvT(256->64) as a GEMM, + feat_s1, LN2d, GELU.  It only needs the final image
  //    keys, so it runs on the side stream next to the final token->image attention and the token-side heads
  cudaStream_t up_side;
  VLS_TRY(fork_begin(1, st, &up_side));
  {
    GemmArgs g = lin(keys_h, C, (long long)T * C, w->up1_w, T, 256, C, B, w->up1_b, scratch, 0, 256, (long long)T * 256);
    VLS_TRY(launch_gemm(g, up_side));
  }
  VLS_TRY(launch_up1_post(scratch, feat_s1, s1_dtype, s1_bstride, B, H, W, w->up_ln_w, w->up_ln_b, LN2D_EPS, up1, up_side));
  // -- final tokens -> image attention (sam/transformer.py:127-132)
  if (!img_fused) {
    GemmArgs g = lin(keys_h, C, (long long)T * C, w->final_img_w, T, 256, C, B, w->final_img_b, kvq, 1, 256, (long long)T * 256);
    g.residual = w->final_pe_add; g.ld_res = 256; g.res_bstride = 0;
    planes_out(g);
    VLS_TRY(launch_gemm(g, st));
  }
  if (fused) VLS_TRY(dec_tok(DEC_TOK_CROSS, nullptr, &w->final_t2i, w->nf_w, w->nf_b, (long long)T * 256, st));
  else VLS_TRY(t2i(w->final_t2i, kvq, 256, w->nf_w, w->nf_b));
  // queries == hs: [0]=obj score token, [1]=iou token, [2..5]=mask tokens (mask_decoder.py:213-215)

  // -- the three token-side heads advance layer by layer in ONE launch per layer: hyper-network MLPs on the 4 mask
  //    tokens (mask_decoder.py:227-232), IoU head on hs[:,1] and object-score head on hs[:,0] (:237-240)
  {
    SmallLinArgs h[3];
    SmallLinArgs& hy = h[0];
    hy.G = 4; hy.R = B; hy.K = 256; hy.N = 256; hy.act = 1;
    hy.x = queries + 2 * 256; hy.x_sg = 256; hy.x_sr = (long long)Nt * 256;
    hy.W = w->hyper_w[0]; hy.w_sg = 256 * 256; hy.bias = w->hyper_b[0]; hy.b_sg = 256;
    hy.out = hy1; hy.o_sg = (long long)B * 256; hy.o_sr = 256;
    for (int head = 0; head < 2; ++head) {   // 0: IoU, 1: object score
      SmallLinArgs& s = h[1 + head];
      s.G = 1; s.R = B; s.K = 256; s.N = 256; s.act = 1;
      s.x = queries + (head == 0 ? 256 : 0); s.x_sr = (long long)Nt * 256;
      s.W = (head == 0 ? w->iou_w : w->obj_w)[0]; s.bias = (head == 0 ? w->iou_b : w->obj_b)[0];
      s.out = hd1 + (size_t)head * B * 256; s.o_sr = 256;
    }
    VLS_TRY(launch_small_linear_multi(h, 3, st));
    hy.x = hy1; hy.x_sg = (long long)B * 256; hy.x_sr = 256;
    hy.W = w->hyper_w[1]; hy.bias = w->hyper_b[1]; hy.out = hy2;
    for (int head = 0; head < 2; ++head) {
      SmallLinArgs& s = h[1 + head];
      s.x = hd1 + (size_t)head * B * 256; s.x_sr = 256;
      s.W = (head == 0 ? w->iou_w : w->obj_w)[1]; s.bias = (head == 0 ? w->iou_b : w->obj_b)[1];
      s.out = hd2 + (size_t)head * B * 256;
    }
    VLS_TRY(launch_small_linear_multi(h, 3, st));
    hy.x = hy2; hy.N = 32; hy.act = 0;
    hy.W = w->hyper_w[2]; hy.w_sg = 32 * 256; hy.bias = w->hyper_b[2]; hy.b_sg = 32;
    hy.out = hyper; hy.o_sg = 32; hy.o_sr = 4 * 32;
    for (int head = 0; head < 2; ++head) {
      SmallLinArgs& s = h[1 + head];
      s.x = hd2 + (size_t)head * B * 256;
      s.W = (head == 0 ? w->iou_w : w->obj_w)[2]; s.bias = (head == 0 ? w->iou_b : w->obj_b)[2];
      if (head == 0) { s.N = 4; s.act = w->iou_sigmoid ? 3 : 0; s.out = iou; s.o_sr = 4; }
      else { s.N = 1; s.act = 0; s.out = obj_logits; s.o_sr = 1; }
    }
    VLS_TRY(launch_small_linear_multi(h, 3, st));
  }
  // -- ConvT(64->32) + feat_s0 + GELU + (hyper @ upscaled) fused (mask_decoder.py:225,234): needs both branches
  VLS_TRY(fork_join(1, st));
  if (w->up2_wh && s0_dtype == VLS_F32 && (2 * W) % 32 == 0 && g_up2_tc)   // tcgen05 path (decoder.cu)
    VLS_TRY(launch_up2_masks_tc(up1, w->up2_wh, w->up2_b, reinterpret_cast<const float*>(feat_s0), s0_bstride, hyper, B, 4, 2 * H,
                                2 * W, masks, st));
  else
    VLS_TRY(launch_up2_masks(up1, w->up2_w, w->up2_b, feat_s0, s0_dtype, s0_bstride, hyper, B, 4, 2 * H, 2 * W, masks, st));
  // -- mask tokens out
  return launch_gather_rows(queries + 2 * 256, (long long)Nt * 256, 256, B, 4, 256, tokens_out, st);
}

// ================================================================== post-decoder glue
// defer: the object-pointer MLP (3 token linears + gate; only the bank shift / the session state need its result) runs on
// a forked stream and is NOT joined here: the caller overlaps it with the memory encoder and calls vls_sam_heads_join.
static int sam_heads_post_impl(const vls_obj_ptr_weights* w, const float* masks, const float* iou, const float* tokens,
                               const float* obj_logits, int B, int multimask, int HW, float* low_res_masks, float* obj_ptr,
                               int* best_idx, float* is_obj, float* occluded, void* workspace, size_t workspace_bytes,
                               cudaStream_t st, bool defer) {
  VLS_REQUIRE(w && masks && iou && tokens && obj_logits && low_res_masks && obj_ptr && best_idx && is_obj && occluded,
              "sam_heads_post: null argument");
  VLS_REQUIRE(workspace && workspace_bytes >= (size_t)B * 256 * 3 * 4, "sam_heads_post: workspace too small");
  float* tok = (float*)workspace;
  float* h1 = tok + (size_t)B * 256;
  float* h2 = h1 + (size_t)B * 256;
  VLS_TRY(launch_select_best(masks, iou, tokens, obj_logits, B, 4, multimask, HW, low_res_masks, tok, best_idx, is_obj,
                             occluded, st));
  cudaStream_t ps = st;
  if (defer) VLS_TRY(fork_begin(4, st, &ps));
  SmallLinArgs s;
  s.G = 1; s.R = B; s.K = 256; s.N = 256; s.act = 1;
  s.x = tok; s.x_sr = 256; s.W = w->w[0]; s.bias = w->b[0]; s.out = h1; s.o_sr = 256;
  VLS_TRY(launch_small_linear(s, ps));
  s.x = h1; s.W = w->w[1]; s.bias = w->b[1]; s.out = h2;
  VLS_TRY(launch_small_linear(s, ps));
  s.x = h2; s.W = w->w[2]; s.bias = w->b[2]; s.out = obj_ptr; s.act = 0;
  VLS_TRY(launch_small_linear(s, ps));
  return launch_gate_ptr(obj_ptr, is_obj, w->no_obj_ptr, B, ps);
}

int vls_sam_heads_post(const vls_obj_ptr_weights* w, const float* masks, const float* iou, const float* tokens,
                       const float* obj_logits, int B, int multimask, int HW, float* low_res_masks, float* obj_ptr,
                       int* best_idx, float* is_obj, float* occluded, void* workspace, size_t workspace_bytes,
                       vls_stream_t stream_) {
  return sam_heads_post_impl(w, masks, iou, tokens, obj_logits, B, multimask, HW, low_res_masks, obj_ptr, best_idx, is_obj,
                             occluded, workspace, workspace_bytes, (cudaStream_t)stream_, false);
}

int vls_sam_heads_post_deferred(const vls_obj_ptr_weights* w, const float* masks, const float* iou, const float* tokens,
                                const float* obj_logits, int B, int multimask, int HW, float* low_res_masks, float* obj_ptr,
                                int* best_idx, float* is_obj, float* occluded, void* workspace, size_t workspace_bytes,
                                vls_stream_t stream_) {
  return sam_heads_post_impl(w, masks, iou, tokens, obj_logits, B, multimask, HW, low_res_masks, obj_ptr, best_idx, is_obj,
                             occluded, workspace, workspace_bytes, (cudaStream_t)stream_, true);
}

int vls_sam_heads_join(vls_stream_t stream_) { return fork_join(4, (cudaStream_t)stream_); }

// ================================================================== memory encoder
size_t vls_mem_encoder_workspace_bytes(int B, int H, int W) {
  const size_t T = (size_t)H * W;
  size_t n = 0;
  n += align256(B * 64 * T * 4 * 2);    // stage 1 rows (8H x 8W x 4)
  n += align256(B * 16 * T * 16 * 2);   // stage 2 rows (4H x 4W x 16)
  n += align256(B * 4 * T * 64 * 2);    // stage 3 rows (2H x 2W x 64)
  n += align256(B * T * 576 * 2);       // im2col
  n += 2 * align256(B * T * C * 4);     // f32 scratch + x
  n += 2 * align256(B * T * C * 2);     // bf16 t, pix rows
  n += align256(B * T * 1024 * 2);      // CXBlock hidden
  n += align256(B * T * 64 * 4);        // out rows f32
  return n + 4096;
}

int vls_mem_encoder_forward(const vls_mem_encoder_weights* w, const void* pix_feat, int pix_dtype, int pix_layout,
                            const long long pix_strides[4], const float* mask, int mask_mode, float sig_scale,
                            float sig_bias, const float* occluded_gate, int B, int H, int W, void* out_nchw,
                            int out_dtype, void* out_rows_bf16, void* workspace, size_t workspace_bytes,
                            vls_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  VLS_REQUIRE(w && pix_feat && mask && (out_nchw || out_rows_bf16), "mem_encoder: null argument");
  VLS_REQUIRE(mask_mode >= 0 && mask_mode <= 4, "mem_encoder: bad mask_mode");
  VLS_REQUIRE(workspace && workspace_bytes >= vls_mem_encoder_workspace_bytes(B, H, W), "mem_encoder: workspace too small");
  const int T = H * W;
  Workspace ws(workspace, workspace_bytes);
  void* m1 = ws.take((size_t)B * 64 * T * 4 * 2);
  void* m2 = ws.take((size_t)B * 16 * T * 16 * 2);
  void* m3 = ws.take((size_t)B * 4 * T * 64 * 2);
  void* col = ws.take((size_t)B * T * 576 * 2);
  float* scratch = (float*)ws.take((size_t)B * T * C * 4);
  float* x = (float*)ws.take((size_t)B * T * C * 4);
  void* t = ws.take((size_t)B * T * C * 2);
  void* pix = ws.take((size_t)B * T * C * 2);
  void* hid = ws.take((size_t)B * T * 1024 * 2);
  float* orow = (float*)ws.take((size_t)B * T * 64 * 4);
  VLS_REQUIRE(m1 && m2 && m3 && col && scratch && x && t && pix && hid && orow, "mem_encoder: workspace carve failed");

  // x = pix_feat_proj(pix_feat) (memory_encoder.py:174) does not depend on the mask: layout / precision conversion +
  // projection run on the side stream next to the mask down-sampler, whose last 1x1 conv then adds x as its residual
  {
    cudaStream_t pside;
    VLS_TRY(fork_begin(3, st, &pside));
    if (pix_layout == 0) {
      VLS_TRY(launch_nchw_to_rows(pix_feat, pix_dtype, pix_strides, nullptr, 0, nullptr, B, C, H, W, nullptr, pix, pside));
    } else {
      VLS_TRY(launch_axpy_rows(pix_feat, pix_dtype, pix_strides[0], pix_strides[1], nullptr, 0, 0, 0, 0.f, B, T, C, nullptr,
                               pix, pside));
    }
    VLS_TRY(launch_gemm(lin(pix, C, (long long)T * C, w->pix_w, T, C, C, B, w->pix_b, x, 0, C, (long long)T * C), pside));
  }
  // mask down-sampler (memory_encoder.py:17-58): 3x (conv3x3 s2 + LN2d + GELU) on CUDA cores, the 4th as im2col + GEMM
  VLS_TRY(launch_mds1(mask, mask_mode, B, 16 * H, 16 * W, (mask_mode == 2 || mask_mode == 3) ? 4 : 1, sig_scale, sig_bias, w->c1_w, w->c1_b,
                      w->ln1_w, w->ln1_b, LN2D_EPS, m1, st));
  VLS_TRY(launch_mds2(m1, B, 8 * H, 8 * W, w->c2_w, w->c2_b, w->ln2_w, w->ln2_b, LN2D_EPS, m2, st));
  if (w->c3_wh && g_mds3_tc) {
    // stage 3 (16 -> 64 channels) on tensor cores: im2col [4T][144] + tcgen05 GEMM (K = 144, zero-filled to 192 by TMA) + LN2d(64)
    // + GELU; `col` and `scratch` are free until stage 4 (same sizes: 4T x 144 = T x 576, 4T x 64 = T x 256)
    VLS_TRY(launch_im2col3x3s2(m2, B, 4 * H, 4 * W, 16, col, st));
    VLS_TRY(launch_gemm(lin(col, 144, (long long)4 * T * 144, w->c3_wh, 4 * T, 64, 144, B, w->c3_b, scratch, 0, 64, (long long)4 * T * 64), st));
    VLS_TRY(launch_ln64_gelu(scratch, (long long)B * 4 * T, w->ln3_w, w->ln3_b, LN2D_EPS, m3, st));
  } else {
    VLS_TRY(launch_mds3(m2, B, 4 * H, 4 * W, w->c3_w, w->c3_b, w->ln3_w, w->ln3_b, LN2D_EPS, m3, st));
  }
  VLS_TRY(launch_im2col3x3s2(m3, B, 2 * H, 2 * W, 64, col, st));
  VLS_TRY(launch_gemm(lin(col, 576, (long long)T * 576, w->c4_w, T, C, 576, B, w->c4_b, scratch, 0, C, (long long)T * C), st));
  VLS_TRY(launch_ln256(scratch, B, T, w->ln4_w, w->ln4_b, LN2D_EPS, 1, nullptr, 0, 0, t, (long long)T * C, C, st));
  // x += mask features (memory_encoder.py:175)
  VLS_TRY(fork_join(3, st));
  {
    GemmArgs g = lin(t, C, (long long)T * C, w->c5_w, T, C, C, B, w->c5_b, x, 0, C, (long long)T * C);
    g.residual = x; g.ld_res = C; g.res_bstride = (long long)T * C;
    VLS_TRY(launch_gemm(g, st));
  }
  // fuser: 2 x CXBlock (memory_encoder.py:103-117); gamma is folded into pw2
  for (int i = 0; i < 2; ++i) {
    const vls_cx_block& cx = w->cx[i];
    VLS_TRY(launch_dwconv7_ln(x, B, H, W, cx.dw_w, cx.dw_b, cx.ln_w, cx.ln_b, LN2D_EPS, t, st));
    if (g_ffn_fused) {   // pwconv1 + GELU + pwconv2 + residual in one cluster kernel: the [T][1024] hidden tensor stays in TMEM
      VLS_TRY(launch_ffn_fused(t, C, (long long)T * C, cx.pw1_w, cx.pw1_b, cx.pw2_w, cx.pw2_b, x, (long long)T * C, B, T, st, 1024, 1));
      continue;
    }
    GemmArgs g1 = lin(t, C, (long long)T * C, cx.pw1_w, T, 1024, C, B, cx.pw1_b, hid, 1, 1024, (long long)T * 1024);
    g1.act = 2;
    VLS_TRY(launch_gemm(g1, st));
    GemmArgs g2 = lin(hid, 1024, (long long)T * 1024, cx.pw2_w, T, C, 1024, B, cx.pw2_b, x, 0, C, (long long)T * C);
    g2.residual = x; g2.ld_res = C; g2.res_bstride = (long long)T * C;
    VLS_TRY(launch_gemm(g2, st));
  }
  // out_proj 256 -> 64 (memory_encoder.py:177) (+ occlusion embedding, sam2_base.py:716-722)
  VLS_TRY(launch_axpy_rows(x, 0, C, (long long)T * C, nullptr, 0, 0, 0, 0.f, B, T, C, nullptr, t, st));
  VLS_TRY(launch_gemm(lin(t, C, (long long)T * C, w->out_w, T, 64, C, B, w->out_b, orow, 0, 64, (long long)T * 64), st));
  const float* gate = (occluded_gate && w->no_obj_embed) ? occluded_gate : nullptr;
  if (out_nchw)
    VLS_TRY(launch_rows_to_nchw(orow, B, 64, T, gate, w->no_obj_embed, out_dtype == VLS_F32 ? (float*)out_nchw : nullptr,
                                out_dtype == VLS_BF16 ? out_nchw : nullptr, st));
  if (out_rows_bf16) VLS_TRY(launch_rows_gate_cast(orow, B, T, 64, gate, w->no_obj_embed, out_rows_bf16, st));
  return 0;
}

}  // extern "C"
