"""Thin Python wrappers over the building-block entry points of libvls_b200.so (used by the host
modules and by the parity tests).  Everything here launches hand-written sm_100a kernels."""
import torch

from . import _lib
from ._lib import GemmDesc, check, lib, ptr, stream

ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2}


def gemm(a, w, bias=None, bias_mode=1, act=None, rope=None, rope_rows=0, residual=None, out=None,
         out_dtype=torch.bfloat16):
    """C[b,m,n] = act(A[b,m,:] . W[n,:] + bias) (+ residual).

    a: bf16 [M,K] or [B,M,K] (last dim contiguous); w: bf16 [N,K] or [B,N,K];
    rope: (cos, sin) f32 [P,128] tables -> rotates column pairs of rows m < rope_rows at position m % P.
    """
    _lib.require_cuda(a, w)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    a3 = a if a.dim() == 3 else a.unsqueeze(0)
    w3 = w if w.dim() == 3 else w.unsqueeze(0)
    B, M, K = a3.shape
    N = w3.shape[1]
    assert w3.shape[2] == K and a3.stride(2) == 1 and w3.stride(2) == 1
    if out is None:
        out = torch.empty((B, M, N), device=a.device, dtype=out_dtype)
    o3 = out if out.dim() == 3 else out.unsqueeze(0)
    assert o3.stride(2) == 1 and o3.shape == (B, M, N)
    d = GemmDesc()
    d.A, d.lda, d.a_bstride = ptr(a3), a3.stride(1), a3.stride(0)
    d.W, d.ldw, d.w_bstride = ptr(w3), w3.stride(1), (w3.stride(0) if w3.shape[0] > 1 else 0)
    d.M, d.N, d.K, d.batch = M, N, K, B
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        d.bias, d.bias_mode = ptr(bias), bias_mode
    d.act = ACT[act]
    if rope is not None:
        cos, sin = rope
        assert cos.dtype == torch.float32 and cos.shape[1] == 128 and cos.is_contiguous() and sin.is_contiguous()
        d.rope_cos, d.rope_sin, d.rope_period, d.rope_rows = ptr(cos), ptr(sin), cos.shape[0], rope_rows
    if residual is not None:
        r3 = residual if residual.dim() == 3 else residual.unsqueeze(0)
        assert r3.dtype == torch.float32 and r3.stride(2) == 1
        d.residual, d.ld_res, d.res_bstride = ptr(r3), r3.stride(1), (r3.stride(0) if r3.shape[0] > 1 else 0)
    d.C, d.c_bf16, d.ldc, d.c_bstride = ptr(o3), int(o3.dtype == torch.bfloat16), o3.stride(1), o3.stride(0)
    check(lib().vls_gemm_bf16(d, stream()), "vls_gemm_bf16")
    return out if a.dim() == 3 else out.reshape(M, N) if out.dim() == 3 else out


def attention_d256(q, k, vt, scale=None, splits=0, out=None):
    """softmax(q k^T * scale) v for one head of dim 256.  q [B,Nq,256], k [B,Nk,256], vt [B,256,>=Nk]
    (V transposed, row stride a multiple of 8), all bf16 with unit inner stride."""
    _lib.require_cuda(q, k, vt)
    B, Nq, D = q.shape
    Nk = k.shape[1]
    assert D == 256 and k.shape[2] == 256 and vt.shape[1] == 256 and vt.shape[2] >= Nk
    assert q.stride(2) == 1 and k.stride(2) == 1 and vt.stride(2) == 1
    if out is None:
        out = torch.empty((B, Nq, 256), device=q.device, dtype=torch.bfloat16)
    scale = float(scale) if scale is not None else 1.0 / 16.0
    nbytes = lib().vls_attention_workspace_bytes(B, Nq, Nk, splits)
    ws = torch.empty(max(nbytes, 1), device=q.device, dtype=torch.uint8)
    check(lib().vls_attention_d256(ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0), ptr(vt),
                                   vt.stride(1), vt.stride(0), B, Nq, Nk, scale, splits, ptr(out), out.stride(1),
                                   out.stride(0), ptr(ws), nbytes, stream()), "vls_attention_d256")
    return out


def attention_qk256(q, k, v, v_rows, scale=None, splits=0, out=None):
    """softmax(q k^T * scale) v with q/k head dim 256 and value dim dv in {256, 64}.  v_rows=False: v is V transposed
    [B,dv,>=Nk]; v_rows=True (dv == 64): v is [B,Nk,64] rows as the memory bank stores them.  Returns [B,Nq,dv] bf16."""
    _lib.require_cuda(q, k, v)
    B, Nq, D = q.shape
    Nk = k.shape[1]
    dv = v.shape[2] if v_rows else v.shape[1]
    assert D == 256 and k.shape[2] == 256 and (v.shape[1] == Nk if v_rows else v.shape[2] >= Nk)
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    if out is None:
        out = torch.empty((B, Nq, dv), device=q.device, dtype=torch.bfloat16)
    scale = float(scale) if scale is not None else 1.0 / 16.0
    nbytes = lib().vls_attention_qk256_workspace_bytes(B, Nq, Nk, dv, splits)
    ws = torch.empty(max(nbytes, 1), device=q.device, dtype=torch.uint8)
    check(lib().vls_attention_qk256(ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0), ptr(v),
                                    v.stride(1), v.stride(0), dv, int(bool(v_rows)), B, Nq, Nk, scale, splits, ptr(out),
                                    out.stride(1), out.stride(0), ptr(ws), nbytes, stream()), "vls_attention_qk256")
    return out


def ffn_fused(t, w1, b1, w2, b2, x):
    """x += relu(t @ w1^T + b1) @ w2^T + b2 in place.  t [B,M,256] bf16, w1 [2048,256] bf16, w2 [256,2048] bf16,
    b1 / b2 f32, x [B,M,256] f32 contiguous."""
    _lib.require_cuda(t, w1, b1, w2, b2, x)
    B, M, C = t.shape
    assert C == 256 and tuple(w1.shape) == (2048, 256) and tuple(w2.shape) == (256, 2048) and x.shape == t.shape
    assert t.dtype == torch.bfloat16 and w1.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16
    assert x.dtype == torch.float32 and x.is_contiguous() and t.stride(2) == 1 and w1.is_contiguous() and w2.is_contiguous()
    check(lib().vls_ffn_fused(ptr(t), t.stride(1), t.stride(0), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(x), x.stride(0), B, M,
                              stream()), "vls_ffn_fused")
    return x


def mem_attn_layer_tail(ao, w0, b0, ln_w, ln_b, w1, b1, w2, b2, x_in, ln2_w, ln2_b, out_dtype=torch.bfloat16, eps=1e-5):
    """x_mid = x_in + ao @ w0^T + b0; x_out = x_mid + relu(LN(x_mid) @ w1^T + b1) @ w2^T + b2; t = LN2(x_out).
    ao [B,M,64] bf16, w0 [256,64] bf16, x_in [B,M,256] f32.  Returns (x_out f32, t)."""
    _lib.require_cuda(ao, w0, b0, ln_w, ln_b, w1, b1, w2, b2, x_in, ln2_w, ln2_b)
    B, M, _ = ao.shape
    assert ao.is_contiguous() and x_in.is_contiguous() and x_in.dtype == torch.float32
    x_out = torch.empty_like(x_in)
    t = torch.empty((B, M, 256), device=ao.device, dtype=out_dtype)
    check(lib().vls_mem_attn_layer_tail(ptr(ao), ptr(w0), ptr(b0), ptr(ln_w), ptr(ln_b), eps, ptr(w1), ptr(b1), ptr(w2), ptr(b2),
                                        ptr(x_in), ptr(x_out), ptr(ln2_w), ptr(ln2_b), eps, ptr(t), _lib.VLS_DTYPE[out_dtype],
                                        t.stride(1), t.stride(0), B, M, stream()), "vls_mem_attn_layer_tail")
    return x_out, t


def resize_bilinear(x, size):
    """F.interpolate(x, size, mode="bilinear", align_corners=False) for f32 [N,C,h,w]."""
    _lib.require_cuda(x)
    x = x.float().contiguous()
    n, c, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty((n, c, H, W), device=x.device, dtype=torch.float32)
    check(lib().vls_resize_bilinear(ptr(x), n * c, h, w, ptr(out), H, W, stream()), "vls_resize_bilinear")
    return out


def bank_shift(bank, hw, n_mem, n_ptr, tokens_per_ptr, new_rows, new_ptr):
    """Advance the device memory bank [B, n_mem*hw + n_ptr*tokens_per_ptr, 64] (bf16, reference key order) by one frame
    in place: see vls_bank_shift."""
    _lib.require_cuda(bank, new_rows, new_ptr)
    assert bank.dtype == torch.bfloat16 and bank.is_contiguous() and bank.shape[1] == n_mem * hw + n_ptr * tokens_per_ptr
    assert new_rows.dtype == torch.bfloat16 and new_rows.is_contiguous() and new_ptr.dtype == torch.float32 and new_ptr.is_contiguous()
    check(lib().vls_bank_shift(ptr(bank), bank.shape[0], hw, n_mem, n_ptr, tokens_per_ptr, ptr(new_rows), ptr(new_ptr),
                               stream()), "vls_bank_shift")


def _same_dense_layout(a, b):
    """Both tensors cover their memory without holes in the same (possibly permuted) order: a flat byte copy is dst.copy_(src)."""
    if a.shape != b.shape or a.stride() != b.stride():
        return False
    expect = 1
    for st, sz in sorted((st, sz) for st, sz in zip(a.stride(), a.shape) if sz > 1):
        if st != expect:
            return False
        expect *= sz
    return True


def copy_many(srcs, dsts):
    """dst.copy_(src) for every pair with ONE kernel launch per 8 pairs (vls_multi_copy) when both sides are CUDA tensors of
    one dtype with the same dense memory layout (contiguous, or e.g. both channel-last views) -- in 16-byte vectors, small
    unaligned tensors (<= 4 KB) in bytes; other pairs fall back to Tensor.copy_ (a memcpy / kernel of their own)."""
    import ctypes

    fast = []
    for i, (a, b) in enumerate(zip(srcs, dsts)):
        nb = a.numel() * a.element_size()
        vec = nb % 16 == 0 and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
        if (a.is_cuda and b.is_cuda and a.dtype == b.dtype and nb > 0 and (vec or nb <= 4096)
                and ((a.is_contiguous() and b.is_contiguous() and a.numel() == b.numel()) or _same_dense_layout(a, b))):
            fast.append(i)
        else:
            b.copy_(a)
    for g in range(0, len(fast), 8):
        grp = fast[g:g + 8]
        n = len(grp)
        src = (ctypes.c_void_p * n)(*[srcs[i].data_ptr() for i in grp])
        dst = (ctypes.c_void_p * n)(*[dsts[i].data_ptr() for i in grp])
        nb = (ctypes.c_size_t * n)(*[srcs[i].numel() * srcs[i].element_size() for i in grp])
        check(lib().vls_multi_copy(src, dst, nb, n, stream()), "vls_multi_copy")
    return dsts


def clone_many(tensors):
    """[t.clone() for t in tensors] with one launch (see copy_many)."""
    return copy_many(tensors, [torch.empty_like(t) for t in tensors])


def resize_binarize(x, size, thresh=0.0, packed=False):
    """(F.interpolate(x, size, mode="bilinear", align_corners=False) > thresh) in one pass for f32 [N,C,h,w], without
    the f32 full-resolution intermediate.  Returns uint8 [N,C,H,W] (0/1), or with packed=True the numpy.packbits
    layout uint8 [N,C,H,ceil(W/8)] (first pixel = most significant bit)."""
    _lib.require_cuda(x)
    x = x.float().contiguous()
    n, c, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty((n, c, H, (W + 7) // 8 if packed else W), device=x.device, dtype=torch.uint8)
    check(lib().vls_resize_binarize(ptr(x), n * c, h, w, H, W, float(thresh), None if packed else ptr(out),
                                    ptr(out) if packed else None, stream()), "vls_resize_binarize")
    return out


def linear_f32(x, w_bf16, bias=None, act=None):
    """Small-row linear: x f32 [R,K], w bf16 [N,K] -> f32 [R,N]."""
    _lib.require_cuda(x, w_bf16)
    x = x.float().contiguous()
    R, K = x.shape
    N = w_bf16.shape[0]
    assert w_bf16.dtype == torch.bfloat16 and w_bf16.is_contiguous() and w_bf16.shape[1] == K
    out = torch.empty((R, N), device=x.device, dtype=torch.float32)
    code = {None: 0, "none": 0, "relu": 1, "sigmoid": 3}[act]
    check(lib().vls_linear_f32(ptr(x), K, ptr(w_bf16), ptr(bias), R, N, K, code, ptr(out), N, stream()),
          "vls_linear_f32")
    return out


def add_rowvec(x, vec):
    """x [T,B,C] (f32/bf16, unit channel stride) + vec [..., C] broadcast -> f32 [T,B,C] (seq-first view)."""
    _lib.require_cuda(x, vec)
    if x.stride(2) != 1:
        x = x.contiguous()
    T, B, C = x.shape
    v = vec.reshape(-1).float().contiguous()
    out = torch.empty((B, T, C), device=x.device, dtype=torch.float32)
    check(lib().vls_axpy_rows(ptr(x), _lib.VLS_DTYPE[x.dtype], x.stride(0), x.stride(1), ptr(v), 0, 0, 0, 1.0, B, T, C,
                              ptr(out), 0, stream()), "vls_axpy_rows")
    return out.transpose(0, 1)
