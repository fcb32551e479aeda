/* vls_b200.h -- C ABI of libvls_b200.so: the B200 (sm_100a) implementation of the SAM 2.1
 * per-frame mask-propagation hot path used by Ali2500/Video-LLaVA-Seg.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every buffer is CALLER-OWNED DEVICE memory unless stated
 *   - work is enqueued on `stream` (a cudaStream_t); no allocation, no host synchronisation inside
 *     (vls_mem_attn_forward forks one internal side stream with event fork/join; it is graph-capturable)
 *   - return 0 on success; non-zero on error, with a message in vls_last_error() (thread-local)
 *   - no CPU fallback: without a CUDA device the calls fail, they never compute on the host
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   vls_cc_label            sam2/csrc/connected_components.cu:213-282 (pybind `sam2._C.get_connected_componnets`)
 *   vls_fill_holes          sam2/utils/misc.py:312-338 (fill_holes_in_mask_scores)
 *   vls_mem_attn_forward    sam2/modeling/memory_attention.py:119-169 (MemoryAttention.forward)
 *   vls_mask_decoder_forward sam2/modeling/sam/mask_decoder.py:110-245 (MaskDecoder.forward)
 *   vls_mem_encoder_forward sam2/modeling/memory_encoder.py:158-181 (MemoryEncoder.forward)
 *   vls_sam_heads_post      sam2/modeling/sam2_base.py:359-403 (object gate, best-IoU select, obj_ptr)
 *   vls_gemm_bf16 / vls_attention_d256 / vls_layernorm ...  building blocks, exported for parity tests
 */
#ifndef VLS_B200_H_
#define VLS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vls_stream_t; /* cudaStream_t */
/* dtype codes for activations handed over by the host */
#define VLS_F32 0
#define VLS_BF16 1

const char* vls_last_error(void);
int vls_abi_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's `gpu_launches`). */
long long vls_launch_count(void);
/* A CUDA graph that captured n of this library's launches reports each replay here, so the counter keeps
 * meaning "kernels of this library executed" (bench.py's `gpu_launches`). */
void vls_launch_count_add(long long n);
/* Performance knobs (results are identical for every setting).  "attn_cluster": 1 = each CTA of the attention
 * kernel loads its own K / V^T tiles, 2 = CTAs run as cluster pairs that TMA-multicast half a tile each.
 * "attn_balanced": 1 (default) = long key sequences whose fixed KV split would leave SMs idle are run in balanced mode:
 * the (query tile, key tile) units are dealt out evenly to one persistent CTA per SM; 0 = always fixed splits.
 * "attn_v_rows": 1 (default) = the memory cross-attention reads its value operand straight from the bank rows; 0 = from
 * a transposed copy made once per frame.
 * "ffn_fused": 1 (default) = the memory-attention FFN runs as one cluster kernel; 0 = as two GEMM launches.
 * "tail_fused": 1 (default, needs ffn_fused) = folded out-projection, LayerNorm3, FFN and the following LayerNorm of a
 * memory-attention layer run as ONE launch; 0 = as separate launches.
 * "dec_fused": 1 (default) = the token side of the mask decoder's two-way layers (<= 16 token rows, image tokens a
 * multiple of 64) runs as thread-block-cluster kernels (dec_tok.cu) over head-major image projections; 0 = as the chain
 * of small kernels.
 * "up2_tc": 1 (default) = the mask decoder's second ConvTranspose + hyper-network product runs as a tcgen05 GEMM with a
 * fused epilogue (f32 skip features, width a multiple of 32); 0 = on the FP32 pipe.
 * "mid_fused": 1 (default) = self-attention output projection + residual, LayerNorm2 and the cross-attention query
 * projection (+ RoPE) of a memory-attention layer run as one cluster kernel; 0 = GEMM, LayerNorm, GEMM.
 * "dec_img_fused": 1 (default, needs dec_fused) = image->token attention, its output projection, LayerNorm4 and the next
 * image-side projections of the mask decoder run as one cluster kernel per layer (dec_img.cu); 0 = as four launches.
 * "mds3_tc": 1 (default) = stage 3 of the mask down-sampler (16 -> 64 channels) runs as im2col + tcgen05 GEMM + LayerNorm2d/GELU;
 * 0 = on the FP32 pipe.
 * "pdl": 1 = kernels are launched with programmatic stream serialisation (they all begin with griddepcontrol.wait), so
 * launch latency overlaps the previous kernel's tail; default 0 (also settable with the environment variable VLS_PDL=1):
 * inside the CUDA-graph replay of the steady-state frame it measured no gain.
 * vls_mem_attn_forward_phase only: "mem_attn_head_short": 0 (default) = the head runs through the cross-attention query
 * projection, 1 = it stops after layer 0's q/k/v projections (set it identically for head and rest);
 * "mem_attn_keys0_inline": 1 (default) = behind a full head, the rest projects layer 0's remaining keys on the caller's
 * stream, 0 = on the internal fork; "mem_attn_keys_ahead_all": 0 (default) = ahead_rows applies to layer 0's keys, 1 = to
 * every layer's (identically for head and rest). */
int vls_set_tuning(const char* key, int value);
/* Developer aid: when non-NULL, CTA (0,0,0) of every attention launch writes clock64() stamps of its producer /
 * MMA / softmax roles for the first 64 key tiles into this device buffer of 3*64*8 int64 (tools/trace_attention.py). */
void vls_attention_trace(long long* device_buffer);
/* Same for the fused FFN / layer-tail kernel: 16 int64 stamps of the first epilogue thread of CTA (0,0,0) (tools/trace_ffn.py). */
void vls_ffn_trace(long long* device_buffer);
/* Same for the mask decoder's token-side cluster kernel: every launch while the buffer is set writes 24 int64 stamps of
 * thread 0 of CTA (0,0) and advances the buffer by 24 entries (tools/trace_dec.py); NULL switches it off. */
void vls_dec_trace(long long* device_buffer);
/* Optional live kernel timing: when enabled, the attention launcher brackets its kernel with CUDA events
 * on the launching stream; vls_prof_collect(slot) synchronises them and returns count / total ms and clears
 * the slot.  slot 0 = memory cross-attention launches (Nk > Nq), slot 1 = self-attention launches. */
void vls_prof_enable(int on);
int vls_prof_collect(int slot, int* count, double* total_ms);

/* ---- connected components ------------------------------------------------------------------
 * img: uint8 [n,1,h,w] (non-zero = foreground), h and w even (else error, as the reference
 * asserts at connected_components.cu:226-227).  labels/counts: int32 [n,1,h,w], fully written.
 * label = 1 + min over the component of ((r&~1)*w + (c&~1)); count = component area (8-conn.).
 * workspace: vls_cc_workspace_bytes() bytes (0 when (h/2)*(w/2) <= 16384, e.g. 256x256). */
size_t vls_cc_workspace_bytes(int n, int h, int w);
int vls_cc_label(const uint8_t* img, int n, int h, int w, int32_t* labels, int32_t* counts, void* workspace,
                 size_t workspace_bytes, vls_stream_t stream);
/* In-place: scores f32 [n,1,h,w]; every connected component of (score <= 0) with area <= max_area
 * is overwritten with fill_value (0.1 in the reference). */
size_t vls_fill_holes_workspace_bytes(int n, int h, int w);
int vls_fill_holes(float* scores, int n, int h, int w, int max_area, float fill_value, void* workspace,
                   size_t workspace_bytes, vls_stream_t stream);

/* ---- building blocks (exported for parity tests and for the Python host modules) ------------
 * C[b][m][n] = act(sum_k A[b][m][k] * W[n][k] + bias) (+ residual); A, W bf16; f32 accumulate. */
typedef struct vls_gemm_desc {
  const void* A; long long lda, a_bstride;
  const void* W; long long ldw, w_bstride;   /* w_bstride 0: W shared by all batches */
  int M, N, K, batch;
  const float* bias; int bias_mode;          /* 0 none, 1 per column n, 2 per row m */
  int act;                                   /* 0 none, 1 ReLU, 2 GELU(erf) */
  const float* rope_cos; const float* rope_sin; int rope_period, rope_rows;
  const float* residual; long long ld_res, res_bstride;
  void* C; int c_bf16; long long ldc, c_bstride;
} vls_gemm_desc;
int vls_gemm_bf16(const vls_gemm_desc* d, vls_stream_t stream);

/* softmax(Q K^T * scale) V, one head of dim 256. Q bf16 [B][Nq][ldq], K bf16 [B][Nk][ldk],
 * Vt bf16 [B][256][ldvt] (V transposed), O bf16 [B][Nq][ldo].  splits > 0: that many KV splits per query tile;
 * 0: chosen automatically (fixed splits, or the balanced mode that deals (query tile, key tile) units out evenly to one
 * persistent CTA per SM when fixed splits would idle SMs); -1: force the balanced mode (fails if the shape does not fit). */
size_t vls_attention_workspace_bytes(int B, int Nq, int Nk, int splits);
int vls_attention_d256(const void* Q, long long ldq, long long q_bstride, const void* K, long long ldk,
                       long long k_bstride, const void* Vt, long long ldvt, long long vt_bstride, int B, int Nq, int Nk,
                       float scale, int splits, void* O, long long ldo, long long o_bstride, void* workspace,
                       size_t workspace_bytes, vls_stream_t stream);

/* Generalisation used by the memory cross-attention (r2): q/k head dimension 256, VALUE dimension dv in {256, 64}.
 * v_rows = 0: V is given transposed, bf16 [B][dv][ldv]; v_rows = 1 (dv == 64 only): V is given as rows bf16 [B][Nk][ldv],
 * exactly as the memory bank stores it (consumed as an MN-major tensor-core operand; nothing is transposed or copied).
 * O: bf16 [B][Nq][ldo] with dv columns.  With dv = 64 the memory attention computes softmax(QK^T) mem and folds the
 * value projection into the output projection (sam/transformer.py:311-360, memory_attention.py:66-81). */
size_t vls_attention_qk256_workspace_bytes(int B, int Nq, int Nk, int dv, int splits);
int vls_attention_qk256(const void* Q, long long ldq, long long q_bstride, const void* K, long long ldk,
                        long long k_bstride, const void* V, long long ldv, long long v_bstride, int dv, int v_rows, int B,
                        int Nq, int Nk, float scale, int splits, void* O, long long ldo, long long o_bstride,
                        void* workspace, size_t workspace_bytes, vls_stream_t stream);

/* Back-to-back FFN of a memory-attention layer (memory_attention.py:95-98) in one cluster kernel:
 * x[b][m][:] += relu(t[b][m][:] W1^T + b1) W2^T + b2, t bf16 [B][M][256] (row stride ldt), W1 bf16 [2048][256],
 * W2 bf16 [256][2048], x f32 [B][M][256] updated in place.  The [M][2048] hidden tensor never leaves the SM. */
int vls_ffn_fused(const void* t_bf16, long long ldt, long long t_bstride, const void* w1_bf16, const float* b1,
                  const void* w2_bf16, const float* b2, float* x, long long x_bstride, int B, int M, vls_stream_t stream);

/* The tail of a memory-attention layer in one cluster kernel (memory_attention.py:76-98 + the LayerNorm that follows):
 *   x_mid = x_in + ao W0^T + b0          ao bf16 [B][M][64] = softmax(QK^T) mem, W0 bf16 [256][64] = out_proj.W v_proj.W
 *   x_out = x_mid + relu(LN(x_mid) W1^T + b1) W2^T + b2
 *   t_out = LN2(x_out)                   bf16 or f32, element (b, row, c) at b*t_out_sb + row*t_out_st + c
 * x_in / x_out: f32 [B][M][256], DIFFERENT buffers. */
int vls_mem_attn_layer_tail(const void* ao_bf16, const void* w0_bf16, const float* b0, const float* ln_w, const float* ln_b,
                            float ln_eps, const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2,
                            const float* x_in, float* x_out, const float* ln2_w, const float* ln2_b, float ln2_eps,
                            void* t_out, int t_out_dtype, long long t_out_st, long long t_out_sb, int B, int M,
                            vls_stream_t stream);

/* Bilinear resize of n f32 images [h,w] -> [H,W], align_corners=False, no antialiasing
 * (F.interpolate as used at sam2_base.py:373-378 and sam2_video_predictor.py:416-421). */
int vls_resize_bilinear(const float* in, int n, int h, int w, float* out, int H, int W, vls_stream_t stream);

/* Device-resident memory bank of the steady-state tracker (replaces the per-frame flatten/permute/cat of
 * sam2_base.py:533-646).  bank: bf16 [B][n_mem*HW + n_ptr*tokens_per_ptr][64] in the reference's
 * key order [cond | t-6 .. t-1 | ptr(cond), ptr(t-1) .. ptr(t-(n_ptr-1))].  Advances it one frame in place, in one
 * launch: memory slots 1..n_mem-2 <- 2..n_mem-1, slot n_mem-1 <- new_rows (bf16 [B][HW][64]); pointer slots
 * 2..n_ptr-1 <- 1..n_ptr-2, slot 1 <- new_ptr (f32 [B][tokens_per_ptr*64]).  Slot 0 of both (conditioning frame) stays. */
int vls_bank_shift(void* bank, int B, int HW, int n_mem, int n_ptr, int tokens_per_ptr, const void* new_rows,
                   const float* new_ptr, vls_stream_t stream);
/* n <= 8 device-to-device copies in one launch (src[i] -> dst[i]; in 16-byte vectors when size and addresses are multiples
 * of 16, else -- at most 4096 bytes -- byte by byte): the per-frame snapshots of the outputs a replayed CUDA graph leaves
 * in its static buffers. */
int vls_multi_copy(const void* const* src, void* const* dst, const size_t* bytes, int n, vls_stream_t stream);

/* Fused output stage: the same bilinear resize followed by `> thresh`, without materialising the f32 [H,W] logits
 * (sam2_video_predictor.py:404-424 + the caller's threshold, e.g. llava/inference/utils.py:71-85).
 * out_u8  : [n][H][W] uint8 0/1, or NULL.   out_bits: [n][H][ceil(W/8)] bytes, first pixel = most significant bit
 * (numpy.packbits order), or NULL.  At least one output must be given.  Bit-identical to (vls_resize_bilinear > thresh). */
int vls_resize_binarize(const float* in, int n, int h, int w, int H, int W, float thresh, uint8_t* out_u8, uint8_t* out_bits,
                        vls_stream_t stream);

/* out[r][:n] = act(x[r][:k] . W[n][k]^T + bias), x/out f32, W bf16 row-major, k % 8 == 0;
 * act: 0 none, 1 ReLU, 3 sigmoid.  Warp-per-output kernel for the handful-of-rows linears
 * (object-pointer projections, sam2_base.py:393,633). */
int vls_linear_f32(const float* x, long long ldx, const void* w_bf16, const float* bias, int rows, int n, int k, int act,
                   float* out, long long ldo, vls_stream_t stream);

/* out[b][t][c] = a(t,b,c) + alpha * p(t,b,c), inputs f32/bf16 addressed as t*st + b*sb + c (strides may
 * be 0 to broadcast), output contiguous rows [B][T][C] in f32 or bf16.  C % 4 == 0.  Used for the
 * no-memory embedding add (sam2_base.py:653) and layout/precision conversion. */
int vls_axpy_rows(const void* a, int a_dtype, long long a_st, long long a_sb, const void* p, int p_dtype, long long p_st,
                  long long p_sb, float alpha, int B, int T, int C, void* out, int out_dtype, vls_stream_t stream);

/* CXBlock front half (memory_encoder.py:103-105): depth-wise 7x7 conv (pad 3) + LayerNorm2d over 256 channels.
 * x f32 NHWC [B][H*W][256], dw_w f32 [49][256], out bf16 [B][H*W][256].  Exported for parity / roofline tests. */
int vls_dwconv7_ln(const float* x, int B, int H, int W, const float* dw_w, const float* dw_b, const float* ln_w,
                   const float* ln_b, float eps, void* out_bf16, vls_stream_t stream);
/* LayerNorm over 256 channels of f32 rows -> bf16 rows (optional GELU). */
int vls_layernorm256(const float* x, long long rows, const float* w, const float* b, float eps, int gelu, void* out_bf16,
                     vls_stream_t stream);

/* ---- module-level entry points ---------------------------------------------------------------
 * dtype codes for activations handed over by the host: */

/* Memory attention (memory_attention.py:119-169). All weights bf16 [out][in] row-major (nn.Linear
 * layout) and biases / LayerNorm affines f32, packed once by the host.  d_model 256, kv_in 64, FFN 2048,
 * one head of 256 (sam2.1 YAMLs :26-58). */
typedef struct vls_mem_attn_layer {
  const void* sa_qk_w; const float* sa_qk_b;  /* [512,256] = [q_proj; k_proj] of self_attn */
  const void* sa_v_w;  const float* sa_v_b;   /* [256,256] */
  const void* sa_o_w;  const float* sa_o_b;   /* [256,256] */
  const void* ca_q_w;  const float* ca_q_b;   /* [256,256] cross_attn_image.q_proj */
  const void* ca_ov_w; const float* ca_ov_b;  /* [256,64] = out_proj.W @ v_proj.W, out_proj.W @ v_proj.b + out_proj.b:
                                               * the cross-attention value and output projections folded into one
                                               * (softmax rows sum to 1, so P (mem Wv^T + bv) Wo^T = (P mem)(Wo Wv)^T + Wo bv) */
  const void* l1_w;    const float* l1_b;     /* [2048,256] */
  const void* l2_w;    const float* l2_b;     /* [256,2048] */
  const float *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b;
} vls_mem_attn_layer;
typedef struct vls_mem_attn_weights {
  int num_layers;                 /* <= 8 */
  vls_mem_attn_layer layers[8];
  const float *norm_w, *norm_b;
  /* cross_attn_image.k_proj of all layers stacked, so the memory bank is projected for every layer in one launch:
   * bf16 [num_layers][256][64], f32 [num_layers][256].  (There is no value projection: see ca_ov_w.) */
  const void* ca_k_w_all; const float* ca_k_b_all;
  const float *rope_cos, *rope_sin; /* f32 [Nq][128]: axial table for a sqrt(Nq) x sqrt(Nq) grid */
  int rope_len;                   /* must equal Nq */
} vls_mem_attn_weights;
/* curr/curr_pos: element (t,b,c) at t*st + b*sb + c, c < 256 (seq-first [Nq,B,256] -> st=B*256, sb=256);
 * memory/memory_pos likewise with 64 channels; the last num_obj_ptr_tokens keys are not rotated.
 * out uses the same addressing.  curr_pos may be NULL (pos_enc_at_input off). */
size_t vls_mem_attn_workspace_bytes(int B, int Nq, int Nk);
int vls_mem_attn_forward(const vls_mem_attn_weights* w, const void* curr, int curr_dtype, long long curr_st,
                         long long curr_sb, const void* curr_pos, int pos_dtype, long long pos_st, long long pos_sb,
                         const void* memory, int mem_dtype, long long mem_st, long long mem_sb, const void* memory_pos,
                         int mpos_dtype, long long mpos_st, long long mpos_sb, int B, int Nq, int Nk,
                         int num_obj_ptr_tokens, void* out, int out_dtype, long long out_st, long long out_sb,
                         void* workspace, size_t workspace_bytes, vls_stream_t stream);
/* The same call in two halves, for callers that software-pipeline consecutive frames (the reference runs
 * MemoryAttention.forward, memory_attention.py:119-169, as one call per frame; nothing in it before layer 0's
 * cross-attention, :66-81, reads the memory):
 *   phase 1 (head): x = curr + 0.1 curr_pos, layer 0's norm1 / self-attention / norm2 / cross-attention query projection;
 *                   x and the rotated queries stay in `workspace`; memory, memory_pos and out are not touched (may be NULL)
 *   phase 2 (rest): everything from layer 0's key projection and cross-attention on, for the head that was run LAST on the
 *                   same workspace with the same B, Nq, Nk; curr / curr_pos are not read (may be NULL)
 *   phase 0       : both, = vls_mem_attn_forward.
 *   phase 3, 4    : the head in two halves (3: x, norm1 and layer 0's q/k/v projections; 4: self-attention ... query projection
 *                   and the keys projected ahead), for callers that place the halves differently; 3 then 4 = 1.
 * ahead_rows > 0 (a multiple of Nq, same value in both phases): the head also projects layer 0's keys of memory rows
 * [0, ahead_rows) -- which requires memory / memory_pos in phase 1 -- and the rest only those of rows [ahead_rows, Nk).
 * In phase 1 the memory rows of [ahead_shift_from, ahead_rows) are read ahead_shift rows further on: a caller whose bank is
 * a sliding window calls the head BEFORE it shifts the window by ahead_shift rows (sam2_base.py:533-568: the six most recent
 * memories move down one slot per frame, the conditioning memory stays).
 * Head + rest compute exactly what the whole call computes, element by element: results are bit-identical. */
int vls_mem_attn_forward_phase(const vls_mem_attn_weights* w, const void* curr, int curr_dtype, long long curr_st,
                               long long curr_sb, const void* curr_pos, int pos_dtype, long long pos_st, long long pos_sb,
                               const void* memory, int mem_dtype, long long mem_st, long long mem_sb,
                               const void* memory_pos, int mpos_dtype, long long mpos_st, long long mpos_sb, int B, int Nq,
                               int Nk, int num_obj_ptr_tokens, void* out, int out_dtype, long long out_st, long long out_sb,
                               void* workspace, size_t workspace_bytes, vls_stream_t stream, int phase, int ahead_rows,
                               int ahead_shift_from, int ahead_shift);

/* Mask decoder (sam/mask_decoder.py:110-245 + sam/transformer.py:90-286). */
typedef struct vls_attn_w {
  const void* q_w; const float* q_b; const void* k_w; const float* k_b;
  const void* v_w; const float* v_b; const void* o_w; const float* o_b;
} vls_attn_w;
typedef struct vls_dec_layer {
  vls_attn_w self_attn;        /* 256 -> 256, 8 heads */
  vls_attn_w t2i;              /* q_w [128,256] (tokens), o_w [256,128]; k/v live in img_w */
  vls_attn_w i2t;              /* k_w, v_w [128,256] (tokens), o_w [256,128]; q lives in img_w */
  const void* img_w; const float* img_b;  /* [384,256] = [t2i.k; t2i.v; i2t.q], bias [384] */
  const float* img_pe_add;     /* f32 [T][384] = [image_pe . t2i.k^T | 0 | image_pe . i2t.q^T] */
  const void* mlp1_w; const float* mlp1_b; const void* mlp2_w; const float* mlp2_b;
  const float *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b, *n4_w, *n4_b;
} vls_dec_layer;
typedef struct vls_mask_decoder_weights {
  vls_dec_layer layers[2];
  vls_attn_w final_t2i;        /* q_w [128,256], o_w [256,128] */
  const void* final_img_w; const float* final_img_b;  /* [256,256] = [k; v] */
  const float* final_pe_add;   /* f32 [T][256] = [image_pe . k^T | 0] */
  const float *nf_w, *nf_b;
  const float* out_tokens;     /* f32 [6][256]: obj_score, iou, mask x4 */
  const void* up1_w; const float* up1_b;   /* bf16 [(dy*2+dx)*64+co][ci], f32 [256] (bias tiled x4) */
  const float *up_ln_w, *up_ln_b;          /* [64] */
  const float* up2_w; const float* up2_b;  /* f32 [4 pos][64 ci][32 co], [32] */
  const void* up2_wh;                      /* bf16 [2 (hi, lo)][(dy*2+dx)*32+co][64 ci]: up2_w = hi + lo, for the tensor-core path */
  const void* hyper_w[3]; const float* hyper_b[3]; /* 4 MLPs batched: [4][256][256] x2, [4][32][256] */
  const void* iou_w[3];   const float* iou_b[3];   /* [256,256] x2, [4,256] */
  const void* obj_w[3];   const float* obj_b[3];   /* [256,256] x2, [1,256] */
  int iou_sigmoid;
} vls_mask_decoder_weights;
/* image_embeddings / dense: NCHW views given by 4 element strides (b,c,y,x); a 0 batch stride
 * implements repeat_image / expand().  sparse: f32 [B][Ns][256].  feat_s0 [B or 1][32][4H][4W],
 * feat_s1 [B or 1][64][2H][2W] (batch stride 0 when shared).  Outputs (all f32): masks [B][4][4H][4W],
 * iou [B][4], tokens_out [B][4][256], obj_logits [B]. */
size_t vls_mask_decoder_workspace_bytes(int B, int Ns, int H, int W);
int vls_mask_decoder_forward(const vls_mask_decoder_weights* w, const void* image_embeddings, int emb_dtype,
                             const long long emb_strides[4], const void* dense, int dense_dtype,
                             const long long dense_strides[4], const float* sparse, const void* feat_s0, int s0_dtype,
                             long long s0_bstride, const void* feat_s1, int s1_dtype, long long s1_bstride, int B, int Ns,
                             int H, int W, float* masks, float* iou, float* tokens_out, float* obj_logits,
                             void* workspace, size_t workspace_bytes, vls_stream_t stream);

/* Post-decoder glue of SAM2Base._forward_sam_heads (sam2_base.py:359-403).  masks [B][4][HW], iou [B][4],
 * tokens [B][4][256], obj_logits [B] (all f32).  Outputs: low_res_masks [B][HW] (best-IoU mask, -1024 where the
 * object is absent), obj_ptr [B][256], best_idx [B], is_obj [B] and occluded [B] = 1 - is_obj (f32). */
typedef struct vls_obj_ptr_weights {
  const void* w[3]; const float* b[3];  /* obj_ptr_proj MLP 256-256-256-256 */
  const float* no_obj_ptr;              /* f32 [256] */
} vls_obj_ptr_weights;
int vls_sam_heads_post(const vls_obj_ptr_weights* w, const float* masks, const float* iou, const float* tokens,
                       const float* obj_logits, int B, int multimask, int HW, float* low_res_masks, float* obj_ptr,
                       int* best_idx, float* is_obj, float* occluded, void* workspace, size_t workspace_bytes,
                       vls_stream_t stream);
/* Same, but the object-pointer MLP (obj_ptr is only needed by the memory-bank update and the session state) is enqueued
 * on an internal forked stream and NOT joined: low_res_masks / best_idx / is_obj / occluded are ordered on `stream` as
 * usual, obj_ptr only after vls_sam_heads_join(stream).  Lets the caller overlap it with the memory encoder
 * (sam2_base.py:396-403 vs :709-722 are independent). */
int vls_sam_heads_post_deferred(const vls_obj_ptr_weights* w, const float* masks, const float* iou, const float* tokens,
                                const float* obj_logits, int B, int multimask, int HW, float* low_res_masks, float* obj_ptr,
                                int* best_idx, float* is_obj, float* occluded, void* workspace, size_t workspace_bytes,
                                vls_stream_t stream);
int vls_sam_heads_join(vls_stream_t stream);

/* Memory encoder (memory_encoder.py:158-181). */
typedef struct vls_cx_block {
  const float *dw_w, *dw_b;    /* f32 [49][256], [256] */
  const float *ln_w, *ln_b;
  const void* pw1_w; const float* pw1_b;  /* [1024,256] */
  const void* pw2_w; const float* pw2_b;  /* [256,1024], gamma folded into both */
} vls_cx_block;
typedef struct vls_mem_encoder_weights {
  const float *c1_w, *c1_b, *ln1_w, *ln1_b;   /* f32 [4][9] */
  const float *c2_w, *c2_b, *ln2_w, *ln2_b;   /* f32 [9][4][16] */
  const float *c3_w, *c3_b, *ln3_w, *ln3_b;   /* f32 [9][16][64] */
  const void* c3_wh;                          /* bf16 [64][(ky*3+kx)*16+ci]: the same weights for the tensor-core path (may be NULL) */
  const void* c4_w; const float *c4_b, *ln4_w, *ln4_b;  /* bf16 [256][9*64] (tap-major), f32 */
  const void* c5_w; const float* c5_b;        /* bf16 [256][256] */
  const void* pix_w; const float* pix_b;      /* bf16 [256][256] */
  vls_cx_block cx[2];
  const void* out_w; const float* out_b;      /* bf16 [64][256] */
  const float* no_obj_embed;                  /* f32 [64] or NULL */
} vls_mem_encoder_weights;
/* pix_feat: pix_layout 0 = NCHW view with strides[4] = (b,c,y,x); 1 = token rows, element (t,b,c) at
 * t*strides[0] + b*strides[1] + c.   mask_mode: 0 high-res mask [B][16H][16W] used as is, 1 = sigmoid(it)*scale+bias,
 * 4 = (it > 0)*scale+bias, 2 = LOW-res logits [B][4H][4W] -> sigmoid(bilinear x4)*scale+bias,
 * 3 = (bilinear x4 > 0)*scale+bias.
 * occluded_gate: f32 [B] = (1 - is_obj) or NULL.  Outputs (either may be NULL): out_nchw [B][64][H*W]
 * (out_dtype), out_rows bf16 [B][H*W][64]. */
size_t vls_mem_encoder_workspace_bytes(int B, int H, int W);
int vls_mem_encoder_forward(const vls_mem_encoder_weights* w, const void* pix_feat, int pix_dtype, int pix_layout,
                            const long long pix_strides[4], const float* mask, int mask_mode, float sig_scale,
                            float sig_bias, const float* occluded_gate, int B, int H, int W, void* out_nchw,
                            int out_dtype, void* out_rows_bf16, void* workspace, size_t workspace_bytes,
                            vls_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VLS_B200_H_ */
