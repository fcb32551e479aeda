/* vls_b200.h -- C ABI of libvls_b200.so: the B200 (sm_100a) implementation of the SAM 2.1
 * per-frame mask-propagation hot path used by Ali2500/Video-LLaVA-Seg.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every buffer is CALLER-OWNED DEVICE memory unless stated
 *   - work is enqueued on `stream` (a cudaStream_t); no allocation, no synchronisation inside
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

const char* vls_last_error(void);
int vls_abi_version(void);

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
 * Vt bf16 [B][256][ldvt] (V transposed), O bf16 [B][Nq][ldo].  splits <= 0: chosen automatically. */
size_t vls_attention_workspace_bytes(int B, int Nq, int Nk, int splits);
int vls_attention_d256(const void* Q, long long ldq, long long q_bstride, const void* K, long long ldk,
                       long long k_bstride, const void* Vt, long long ldvt, long long vt_bstride, int B, int Nq, int Nk,
                       float scale, int splits, void* O, long long ldo, long long o_bstride, void* workspace,
                       size_t workspace_bytes, vls_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VLS_B200_H_ */
