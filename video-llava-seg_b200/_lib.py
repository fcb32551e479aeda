"""ctypes binding of libvls_b200.so (C ABI: include/vls_b200.h).

The library is the product: if it is missing or a call fails this module RAISES -- there is no
eager/PyTorch/CPU fallback anywhere on the path (the reference instead swallows a missing
`sam2._C` and silently skips hole filling, sam2/utils/misc.py:321-336).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvls_b200.so")
_lib = None

c_void_p, c_int, c_ll, c_float, c_size_t = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float,
                                            ctypes.c_size_t)
VLS_F32, VLS_BF16 = 0, 1
VLS_DTYPE = {torch.float32: VLS_F32, torch.bfloat16: VLS_BF16}
LL4 = c_ll * 4
LLP = ctypes.POINTER(c_ll)


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_ll), ("a_bstride", c_ll),
        ("W", c_void_p), ("ldw", c_ll), ("w_bstride", c_ll),
        ("M", c_int), ("N", c_int), ("K", c_int), ("batch", c_int),
        ("bias", c_void_p), ("bias_mode", c_int),
        ("act", c_int),
        ("rope_cos", c_void_p), ("rope_sin", c_void_p), ("rope_period", c_int), ("rope_rows", c_int),
        ("residual", c_void_p), ("ld_res", c_ll), ("res_bstride", c_ll),
        ("C", c_void_p), ("c_bf16", c_int), ("ldc", c_ll), ("c_bstride", c_ll),
    ]


def _declare(lib):
    lib.vls_last_error.restype = ctypes.c_char_p
    lib.vls_abi_version.restype = c_int
    lib.vls_launch_count.restype = c_ll
    lib.vls_launch_count_add.restype = None
    lib.vls_launch_count_add.argtypes = [c_ll]
    lib.vls_attention_trace.restype = None
    lib.vls_attention_trace.argtypes = [c_void_p]
    lib.vls_ffn_trace.restype = None
    lib.vls_ffn_trace.argtypes = [c_void_p]
    lib.vls_dec_trace.restype = None
    lib.vls_dec_trace.argtypes = [c_void_p]
    lib.vls_set_tuning.restype = c_int
    lib.vls_set_tuning.argtypes = [ctypes.c_char_p, c_int]
    lib.vls_prof_enable.restype = None
    lib.vls_prof_enable.argtypes = [c_int]
    lib.vls_prof_collect.restype = c_int
    lib.vls_prof_collect.argtypes = [c_int, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_double)]
    lib.vls_cc_workspace_bytes.restype = c_size_t
    lib.vls_cc_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.vls_cc_label.restype = c_int
    lib.vls_cc_label.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.vls_fill_holes_workspace_bytes.restype = c_size_t
    lib.vls_fill_holes_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.vls_fill_holes.restype = c_int
    lib.vls_fill_holes.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]
    lib.vls_gemm_bf16.restype = c_int
    lib.vls_gemm_bf16.argtypes = [ctypes.POINTER(GemmDesc), c_void_p]
    lib.vls_attention_workspace_bytes.restype = c_size_t
    lib.vls_attention_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.vls_attention_qk256_workspace_bytes.restype = c_size_t
    lib.vls_attention_qk256_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.vls_attention_qk256.restype = c_int
    lib.vls_attention_qk256.argtypes = [c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_int, c_int,
                                        c_int, c_int, c_int, c_float, c_int, c_void_p, c_ll, c_ll, c_void_p, c_size_t,
                                        c_void_p]
    lib.vls_ffn_fused.restype = c_int
    lib.vls_ffn_fused.argtypes = [c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int,
                                  c_void_p]
    lib.vls_mem_attn_layer_tail.restype = c_int
    lib.vls_mem_attn_layer_tail.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                            c_int, c_ll, c_ll, c_int, c_int, c_void_p]
    lib.vls_attention_d256.restype = c_int
    lib.vls_attention_d256.argtypes = [c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_int, c_int,
                                       c_int, c_float, c_int, c_void_p, c_ll, c_ll, c_void_p, c_size_t, c_void_p]


def _declare_modules(lib):
    lib.vls_bank_shift.restype = c_int
    lib.vls_bank_shift.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.vls_multi_copy.restype = c_int
    lib.vls_multi_copy.argtypes = [ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_size_t), c_int, c_void_p]
    lib.vls_resize_binarize.restype = c_int
    lib.vls_resize_binarize.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]
    lib.vls_resize_bilinear.restype = c_int
    lib.vls_resize_bilinear.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]
    lib.vls_dwconv7_ln.restype = c_int
    lib.vls_dwconv7_ln.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                   c_void_p]
    lib.vls_layernorm256.restype = c_int
    lib.vls_layernorm256.argtypes = [c_void_p, c_ll, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p]
    lib.vls_axpy_rows.restype = c_int
    lib.vls_axpy_rows.argtypes = [c_void_p, c_int, c_ll, c_ll, c_void_p, c_int, c_ll, c_ll, c_float, c_int, c_int, c_int,
                                  c_void_p, c_int, c_void_p]
    lib.vls_linear_f32.restype = c_int
    lib.vls_linear_f32.argtypes = [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll,
                                   c_void_p]
    lib.vls_mem_attn_workspace_bytes.restype = c_size_t
    lib.vls_mem_attn_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.vls_mem_attn_forward.restype = c_int
    lib.vls_mem_attn_forward.argtypes = [c_void_p, c_void_p, c_int, c_ll, c_ll, c_void_p, c_int, c_ll, c_ll, c_void_p,
                                         c_int, c_ll, c_ll, c_void_p, c_int, c_ll, c_ll, c_int, c_int, c_int, c_int,
                                         c_void_p, c_int, c_ll, c_ll, c_void_p, c_size_t, c_void_p]
    lib.vls_mem_attn_forward_phase.restype = c_int
    lib.vls_mem_attn_forward_phase.argtypes = lib.vls_mem_attn_forward.argtypes + [c_int, c_int, c_int, c_int]
    lib.vls_mask_decoder_workspace_bytes.restype = c_size_t
    lib.vls_mask_decoder_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.vls_mask_decoder_forward.restype = c_int
    lib.vls_mask_decoder_forward.argtypes = [c_void_p, c_void_p, c_int, LLP, c_void_p, c_int, LLP, c_void_p, c_void_p,
                                             c_int, c_ll, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_int, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.vls_sam_heads_post.restype = c_int
    lib.vls_sam_heads_post.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.vls_sam_heads_post_deferred.restype = c_int
    lib.vls_sam_heads_post_deferred.argtypes = lib.vls_sam_heads_post.argtypes
    lib.vls_sam_heads_join.restype = c_int
    lib.vls_sam_heads_join.argtypes = [c_void_p]
    lib.vls_mem_encoder_workspace_bytes.restype = c_size_t
    lib.vls_mem_encoder_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.vls_mem_encoder_forward.restype = c_int
    lib.vls_mem_encoder_forward.argtypes = [c_void_p, c_void_p, c_int, c_int, LLP, c_void_p, c_int, c_float, c_float,
                                            c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                            c_size_t, c_void_p]


def lib():
    """Load (once) and return the ctypes handle; raise loudly if the CUDA extension is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the sm_100a CUDA extension is the only implementation of this path "
                "(no CPU fallback). Build it with `python -m video_llava_seg_b200.build`.")
        handle = ctypes.CDLL(LIB_PATH)
        _declare(handle)
        _declare_modules(handle)
        _lib = handle
        for kv in filter(None, os.environ.get("VLS_TUNING", "").split(",")):   # dev aid: VLS_TUNING="dec_fused=0,up2_tc=0"
            k, v = kv.split("=")
            check(handle.vls_set_tuning(k.strip().encode(), int(v)), f"vls_set_tuning({kv})")
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().vls_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libvls_b200 {what} failed (code {rc}): {msg}")


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    """The current stream of the CURRENT device -- which require_cuda() has made the operands' device."""
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    """All operands must live on ONE CUDA device; that device becomes the current one (the library launches on the
    current device and uses per-device side streams, so operands on another device would mean foreign pointers)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("inputs must be a CUDA tensor")  # connected_components.cu:215
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"operands are on different devices ({dev} and {t.device})")
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        torch.cuda.set_device(dev)
