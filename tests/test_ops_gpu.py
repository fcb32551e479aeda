"""GPU parity tests of the building-block kernels, called through the C ABI (libvls_b200.so):
tcgen05 GEMM epilogues and d=256 flash attention against torch fp32 math on the same bf16-rounded
inputs (tolerances stated per test), connected components bit-exact against the C oracle."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(vls_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


@pytest.mark.parametrize("M,N,K,B", [(128, 64, 64, 1), (4096, 256, 256, 1), (300, 200, 192, 2), (4096, 2048, 256, 1),
                                     (4096, 256, 2048, 1), (256, 28736, 64, 1), (1000, 384, 256, 3)])
def test_gemm_plain(dev, M, N, K, B):
    from video_llava_seg_b200 import ops

    a = _rand((B, M, K), dev, 1).bfloat16()
    w = _rand((N, K), dev, 2, 1.0 / math.sqrt(K)).bfloat16()
    ldc = (N + 7) // 8 * 8
    outbuf = torch.zeros((B, M, ldc), device=dev, dtype=torch.float32)
    out = ops.gemm(a, w, out=outbuf[:, :, :N])
    ref = a.float() @ w.float().t()
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


def test_gemm_epilogues(dev):
    from video_llava_seg_b200 import ops

    M, N, K = 1024, 256, 256
    a = _rand((M, K), dev, 3).bfloat16()
    w = _rand((N, K), dev, 4, 1.0 / 16).bfloat16()
    bias = _rand((N,), dev, 5)
    res = _rand((M, N), dev, 6)
    # bias + relu + residual, f32 out
    out = ops.gemm(a, w, bias=bias, act="relu", residual=res, out_dtype=torch.float32)
    ref = torch.relu(a.float() @ w.float().t() + bias) + res
    assert (out - ref).abs().max().item() < 5e-3
    # gelu, bf16 out
    out = ops.gemm(a, w, bias=bias, act="gelu", out_dtype=torch.bfloat16)
    ref = torch.nn.functional.gelu(a.float() @ w.float().t() + bias)
    assert (out.float() - ref).abs().max().item() < 3e-2
    # per-row bias (used for V^T = Wv X^T)
    brow = _rand((M,), dev, 7)
    out = ops.gemm(a, w, bias=brow, bias_mode=2, out_dtype=torch.float32)
    ref = a.float() @ w.float().t() + brow[:, None]
    assert (out - ref).abs().max().item() < 5e-3


def test_gemm_rope(dev):
    """RoPE epilogue == apply_rotary_enc (position_encoding.py:195-222) on the projected tensor."""
    from oracle import sam2_path as O
    from video_llava_seg_b200 import ops

    side, K = 16, 64
    nq = side * side
    M = 2 * nq + 8  # two repeated memory frames + 8 un-rotated pointer tokens
    a = _rand((M, K), dev, 8).bfloat16()
    w = _rand((256, K), dev, 9, 1.0 / 8).bfloat16()
    bias = _rand((256,), dev, 10, 0.1)
    cos, sin = O.axial_rope_table(side, side)
    out = ops.gemm(a, w, bias=bias, rope=(cos.to(dev).contiguous(), sin.to(dev).contiguous()), rope_rows=2 * nq,
                   out_dtype=torch.float32)
    y = (a.float() @ w.float().t() + bias).cpu()
    ref = torch.cat([O.apply_rope(y[None, None, : 2 * nq], cos.repeat(2, 1), sin.repeat(2, 1))[0, 0], y[2 * nq:]], 0)
    assert (out.cpu() - ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("B,Nq,Nk,splits", [(1, 128, 64, 1), (1, 256, 520, 1), (2, 256, 1000, 2), (1, 4096, 4096, 0),
                                            (1, 4096, 28736, 0), (1, 4096, 28700, 4),
                                            # splits=0 with long keys and few query tiles -> balanced ("stream-K") mode:
                                            # ragged last key tile, 2 batch elements, 3 query tiles (most CTAs cross a tile)
                                            (1, 4096, 28700, 0), (2, 2048, 16500, 0), (1, 300, 40000, 0),
                                            (4, 1024, 16500, 0),
                                            # forced balanced mode (-1): 256 query tiles over 148 CTAs = up to 3 segments
                                            # per CTA (tail of a tile, a whole tile, head of the next)
                                            (8, 4096, 9000, -1), (4, 2048, 16500, -1)])
def test_attention_d256(dev, B, Nq, Nk, splits):
    """vs softmax(QK^T/16)V in fp32 on the same bf16 inputs; P is rounded to bf16 inside the kernel
    (as flash-attention does), so the tolerance is 2e-2 on outputs of magnitude ~1."""
    from video_llava_seg_b200 import ops

    q = _rand((B, Nq, 256), dev, 11).bfloat16()
    k = _rand((B, Nk, 256), dev, 12).bfloat16()
    v = _rand((B, Nk, 256), dev, 13).bfloat16()
    ld = (Nk + 63) // 64 * 64
    vt = torch.zeros((B, 256, ld), device=dev, dtype=torch.bfloat16)
    vt[:, :, :Nk] = v.transpose(1, 2)
    out = ops.attention_d256(q, k, vt[:, :, :Nk] if ld == Nk else vt, splits=splits)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err


@pytest.mark.parametrize("B,M", [(1, 128), (1, 4096), (2, 1000), (8, 4096)])
def test_ffn_fused(dev, B, M):
    """x += relu(t W1^T + b1) W2^T + b2 as ONE cluster kernel (hidden activations in TMEM, partial outputs reduced over
    distributed shared memory) vs the same math in fp32 with the hidden rounded to bf16 as the kernel rounds it; run
    twice to check that the DSMEM reduction is bitwise deterministic."""
    from video_llava_seg_b200 import ops

    t = _rand((B, M, 256), dev, 21).bfloat16()
    w1 = _rand((2048, 256), dev, 22, 1.0 / 16).bfloat16()
    b1 = _rand((2048,), dev, 23, 0.1)
    w2 = _rand((256, 2048), dev, 24, 1.0 / 45).bfloat16()
    b2 = _rand((256,), dev, 25, 0.1)
    x0 = _rand((B, M, 256), dev, 26)
    h = torch.relu(t.float() @ w1.float().t() + b1).bfloat16().float()
    ref = x0 + h @ w2.float().t() + b2
    x = x0.clone()
    ops.ffn_fused(t, w1, b1, w2, b2, x)
    x2 = x0.clone()
    ops.ffn_fused(t, w1, b1, w2, b2, x2)
    torch.cuda.synchronize()
    err = (x - ref).abs().max().item()
    assert err < 5e-3, err
    assert torch.equal(x, x2)


@pytest.mark.parametrize("B,M,out_dtype", [(1, 128, torch.bfloat16), (1, 4096, torch.bfloat16), (2, 1000, torch.float32),
                                           (8, 4096, torch.bfloat16)])
def test_mem_attn_layer_tail(dev, B, M, out_dtype):
    """The whole tail of a memory-attention layer in ONE launch -- folded out-projection (K = 64) + residual, LayerNorm,
    FFN + residual, following LayerNorm -- vs the same chain in fp32 torch (operands rounded to bf16 where the kernel
    rounds them).  Twice: the distributed-shared-memory reductions must be bitwise deterministic."""
    import torch.nn.functional as F

    from video_llava_seg_b200 import ops

    ao = _rand((B, M, 64), dev, 31).bfloat16()
    w0 = _rand((256, 64), dev, 32, 1.0 / 8).bfloat16()
    b0 = _rand((256,), dev, 33, 0.1)
    ln_w, ln_b = 1.0 + _rand((256,), dev, 34, 0.1), _rand((256,), dev, 35, 0.05)
    w1 = _rand((2048, 256), dev, 22, 1.0 / 16).bfloat16()
    b1 = _rand((2048,), dev, 23, 0.1)
    w2 = _rand((256, 2048), dev, 24, 1.0 / 45).bfloat16()
    b2 = _rand((256,), dev, 25, 0.1)
    ln2_w, ln2_b = 1.0 + _rand((256,), dev, 36, 0.1), _rand((256,), dev, 37, 0.05)
    x_in = _rand((B, M, 256), dev, 26)
    x_mid = x_in + ao.float() @ w0.float().t() + b0
    t3 = F.layer_norm(x_mid, (256,), ln_w, ln_b, 1e-5).bfloat16().float()
    h = torch.relu(t3 @ w1.float().t() + b1).bfloat16().float()
    x_ref = x_mid + h @ w2.float().t() + b2
    t_ref = F.layer_norm(x_ref, (256,), ln2_w, ln2_b, 1e-5)
    x_out, t = ops.mem_attn_layer_tail(ao, w0, b0, ln_w, ln_b, w1, b1, w2, b2, x_in, ln2_w, ln2_b, out_dtype)
    x_out2, t2 = ops.mem_attn_layer_tail(ao, w0, b0, ln_w, ln_b, w1, b1, w2, b2, x_in, ln2_w, ln2_b, out_dtype)
    torch.cuda.synchronize()
    ex, et = (x_out - x_ref).abs().max().item(), (t.float() - t_ref).abs().max().item()
    print(f"layer tail B={B} M={M}: x err {ex:.3e} t err {et:.3e}")
    assert ex < 1e-2 and et < (3e-2 if out_dtype == torch.bfloat16 else 1e-2), (ex, et)
    assert torch.equal(x_out, x_out2) and torch.equal(t, t2)


@pytest.mark.parametrize("v_rows", [True, False])
@pytest.mark.parametrize("B,Nq,Nk,splits", [(1, 128, 64, 1), (1, 256, 520, 1), (2, 256, 1000, 2), (1, 4096, 28736, 0),
                                            (1, 4096, 28700, 4), (1, 4096, 28700, 0), (2, 2048, 16500, 0),
                                            (1, 300, 40000, 0), (8, 4096, 9000, -1), (8, 4096, 28736, 0),
                                            (3, 1024, 4100, 1)])
def test_attention_value_dim_64(dev, B, Nq, Nk, splits, v_rows):
    """The memory cross-attention's shape since r2: q/k dim 256, VALUE dim 64 (the raw memory rows; the value projection
    is folded into the output projection).  v_rows: V read as [Nk,64] rows (MN-major tensor-core operand) or from a
    transposed [64,ld] copy.  vs fp32 SDPA on the same bf16 inputs; P is rounded to bf16 inside the kernel."""
    from video_llava_seg_b200 import ops

    q = _rand((B, Nq, 256), dev, 11).bfloat16()
    k = _rand((B, Nk, 256), dev, 12).bfloat16()
    v = _rand((B, Nk, 64), dev, 13).bfloat16()
    if v_rows:
        out = ops.attention_qk256(q, k, v, True, splits=splits)
    else:
        ld = (Nk + 63) // 64 * 64
        vt = torch.zeros((B, 64, ld), device=dev, dtype=torch.bfloat16)
        vt[:, :, :Nk] = v.transpose(1, 2)
        out = ops.attention_qk256(q, k, vt, False, splits=splits)
    assert out.shape == (B, Nq, 64)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err


def test_attention_value_dim_64_is_bitwise_deterministic(dev, vls_lib):
    """Race detector for the two-stage K ring + single V stage of the dv = 64 kernel (balanced mode, full bank)."""
    from video_llava_seg_b200 import ops

    q = _rand((1, 4096, 256), dev, 41).bfloat16()
    k = _rand((1, 28736, 256), dev, 42).bfloat16()
    v = _rand((1, 28736, 64), dev, 43).bfloat16()
    flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
    base = ops.attention_qk256(q, k, v, True).clone()
    for it in range(120):
        if it % 3 == 0:
            flush.zero_()
        if it % 7 == 0:
            (q.float() @ k.float().transpose(1, 2)).sum()
        out = ops.attention_qk256(q, k, v, True)
        assert torch.equal(out, base), f"run {it} differs from run 0"


@pytest.mark.parametrize("shape,density", [((1, 1, 256, 256), 0.5), ((8, 1, 256, 256), 0.62), ((3, 1, 64, 96), 0.4),
                                           ((2, 1, 2, 2), 0.5), ((1, 1, 250, 130), 0.55), ((2, 1, 512, 384), 0.6),
                                           ((1, 1, 1024, 1024), 0.58), ((4, 1, 256, 256), 0.0), ((4, 1, 256, 256), 1.0)])
def test_cc_bit_exact(dev, shape, density):
    from oracle import cc as cc_oracle
    from video_llava_seg_b200.utils.misc import get_connected_components

    g = torch.Generator().manual_seed(hash(shape) % 1000 + int(density * 100))
    m = torch.rand(shape, generator=g) < density
    labels, counts = get_connected_components(m.to(dev))
    rl, rc = cc_oracle.cc_label(m)
    assert torch.equal(labels.cpu(), rl)
    assert torch.equal(counts.cpu(), rc)


def _blobby(n, h, w, seed, thr=0.0):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n, 1, max(h // 8, 1), max(w // 8, 1), generator=g)
    z = torch.nn.functional.interpolate(z, size=(h, w), mode="bilinear", align_corners=False)
    return (z + 0.15 * torch.randn(n, 1, h, w, generator=g)) > thr


def test_cc_lockfree_unions_stress(dev):
    """The union phase is lock-free; a wrong protocol shows up as ONE split component in one launch out of ~50 (found in
    r2: linking roots with atomicMin instead of compare-and-swap lets a path leave its set for a moment, and a concurrent
    path compression then cuts a link of the neighbouring set).  Many launches on mask-like inputs, with the L2 disturbed
    in between, compared on the device against the C oracle."""
    from oracle import cc as cc_oracle
    from video_llava_seg_b200.utils.misc import get_connected_components

    junk = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    for shape, seed, iters in (((32, 1, 256, 256), 256, 150), ((2, 1, 512, 768), 11, 40)):
        m = _blobby(shape[0], shape[2], shape[3], seed)
        rl, rc = cc_oracle.cc_label(m)
        md, rld, rcd = m.to(dev), rl.to(dev), rc.to(dev)
        bad = torch.zeros((), dtype=torch.int64, device=dev)
        for it in range(iters):
            if it % 3 == 0:
                junk.zero_()
            labels, counts = get_connected_components(md)
            bad += (labels != rld).sum() + (counts != rcd).sum()
        assert int(bad.item()) == 0, f"{shape}: {int(bad.item())} wrong pixels over {iters} launches"


@pytest.mark.parametrize("shape", [(32, 1, 256, 256), (6, 1, 32, 1024), (5, 1, 1024, 32), (3, 1, 128, 512), (2, 1, 48, 80)])
def test_cc_blobby_large_components(dev, shape):
    """Mask-like inputs (few large components with ragged borders + speckle): long union-find chains, many
    concurrent unions on the same roots.  Repeated, because a lock-free race would be timing dependent."""
    from oracle import cc as cc_oracle
    from video_llava_seg_b200.utils.misc import get_connected_components

    for rep in range(3):
        m = _blobby(shape[0], shape[2], shape[3], 7 * rep + shape[2])
        rl, rc = cc_oracle.cc_label(m)
        md = m.to(dev)
        for _ in range(3):
            labels, counts = get_connected_components(md)
            assert torch.equal(labels.cpu(), rl) and torch.equal(counts.cpu(), rc)
    # a view whose base address is not 16-byte aligned takes the scalar load path: same result
    buf = torch.zeros(shape[2] * shape[3] + 3, dtype=torch.uint8, device=dev)
    m1 = _blobby(1, shape[2], shape[3], 99)
    view = buf[3:].view(1, 1, shape[2], shape[3])
    view.copy_(m1.to(dev))
    from video_llava_seg_b200 import _lib
    from video_llava_seg_b200._lib import check, ptr, stream

    labels = torch.empty((1, 1, shape[2], shape[3]), dtype=torch.int32, device=dev)
    counts = torch.empty_like(labels)
    check(_lib.lib().vls_cc_label(ptr(view), 1, shape[2], shape[3], ptr(labels), ptr(counts), None, 0, stream()))
    rl, rc = cc_oracle.cc_label(m1)
    assert torch.equal(labels.cpu(), rl) and torch.equal(counts.cpu(), rc)


def test_cc_structured_and_errors(dev):
    from oracle import cc as cc_oracle
    from video_llava_seg_b200.utils.misc import fill_holes_in_mask_scores, get_connected_components

    m = torch.zeros(2, 1, 256, 256, dtype=torch.bool)
    m[0, 0, 10:200, 20:220] = True
    m[0, 0, 50:60, 50:60] = False          # a hole (background component) inside
    m[0, 0, ::7, :] ^= True                 # stripes
    m[1, 0] = torch.arange(256).view(1, -1).expand(256, -1) % 3 == 0  # thin vertical lines
    for t in (m, ~m):
        labels, counts = get_connected_components(t.to(dev))
        rl, rc = cc_oracle.cc_label(t)
        assert torch.equal(labels.cpu(), rl) and torch.equal(counts.cpu(), rc)
    with pytest.raises(RuntimeError):
        get_connected_components(torch.zeros(1, 1, 5, 4, dtype=torch.bool, device=dev))
    with pytest.raises(RuntimeError):
        get_connected_components(torch.zeros(1, 1, 4, 4, dtype=torch.bool))  # CPU tensor
    assert get_connected_components(torch.zeros(0, 1, 4, 4, dtype=torch.bool, device=dev))[0].shape == (0, 1, 4, 4)
    # fused hole filling == reference recipe on top of the oracle labels (utils/misc.py:322-325)
    from oracle import sam2_path as O
    g = torch.Generator().manual_seed(5)
    for shape in ((3, 1, 256, 256), (1, 1, 512, 512)):
        s = torch.randn(shape, generator=g)
        s = torch.nn.functional.avg_pool2d(s, 5, 1, 2) * 3
        got = fill_holes_in_mask_scores(s.to(dev), 8)
        ref = O.fill_holes_in_mask_scores(s, 8, cc=cc_oracle.cc_label)
        assert torch.equal(got.cpu(), ref)
        assert (got.cpu() != s).any()


@pytest.mark.parametrize("shape,size", [((3, 1, 256, 256), (1024, 1024)), ((2, 2, 64, 48), (100, 77)), ((1, 1, 256, 256), (720, 1284))])
def test_resize_binarize_matches_interpolate(dev, shape, size):
    """Fused output stage: (bilinear(x) > t) as bytes and as packed bits == torch's interpolate + threshold, except on
    pixels whose interpolated logit is within float rounding of the threshold (none expected at these seeds)."""
    from video_llava_seg_b200 import ops

    x = _rand(shape, dev, 77) * 3
    ref = torch.nn.functional.interpolate(x, size=size, mode="bilinear", align_corners=False)
    ours = ops.resize_bilinear(x, size)
    for t in (0.0, 0.5):
        u8 = ops.resize_binarize(x, size, t)
        bits = ops.resize_binarize(x, size, t, packed=True)
        assert u8.dtype == torch.uint8 and u8.shape == ref.shape
        assert torch.equal(u8, (ours > t).to(torch.uint8)), "must agree bit for bit with the unfused kernels"
        unsure = (ref - t).abs() < 1e-5
        assert ((u8 == (ref > t).to(torch.uint8)) | unsure).all()
        import numpy as np
        assert torch.equal(bits.cpu(), torch.from_numpy(np.packbits(u8.cpu().numpy(), axis=-1)))


@pytest.mark.parametrize("B,H,W,variant", [(1, 64, 64, 1), (1, 64, 64, 3), (2, 20, 12, 1), (2, 20, 12, 3), (3, 7, 5, 1), (8, 64, 64, 1),
                                             (24, 64, 64, 1), (24, 64, 64, 2), (24, 64, 64, 0),
                                             (40, 30, 22, 1), (40, 30, 22, 2), (40, 30, 22, 0), (64, 10, 66, 1)])
def test_dwconv7_layernorm2d(dev, vls_lib, B, H, W, variant):
    """CXBlock head (memory_encoder.py:86-93): depth-wise 7x7 conv (pad 3) + LayerNorm2d(eps 1e-6) on NHWC rows, f32 in,
    bf16 out.  Small batches take the per-row kernel, large ones the FFMA2 column-strip kernels (variant 1 / 2: input rows
    staged by TMA with the halo zero-filled by the tensor map, 3 / 4 CTAs per SM; 0: the global-load kernel); odd sizes
    exercise the halo / ragged strips / partial last row pieces of all of them."""
    from video_llava_seg_b200._lib import check, ptr, stream

    check(vls_lib.vls_set_tuning(b"dwconv_tma", 1 if variant == 3 else variant))
    check(vls_lib.vls_set_tuning(b"dwconv_small", int(variant == 3)))     # 3: the one-CTA-per-8-pixels kernel of small batches (the default)
    try:
        _dwconv7_case(dev, vls_lib, B, H, W)
    finally:
        check(vls_lib.vls_set_tuning(b"dwconv_tma", 1))
        check(vls_lib.vls_set_tuning(b"dwconv_small", 1))


def _dwconv7_case(dev, vls_lib, B, H, W):
    from video_llava_seg_b200._lib import check, ptr, stream

    x = _rand((B, H * W, 256), dev, 3)
    w = _rand((256, 1, 7, 7), dev, 4) * 0.2
    cb, lw, lb = _rand((256,), dev, 5) * 0.1, 1 + 0.1 * _rand((256,), dev, 6), 0.1 * _rand((256,), dev, 7)
    out = torch.empty((B, H * W, 256), dtype=torch.bfloat16, device=dev)
    wt = w.reshape(256, 49).t().contiguous()
    check(vls_lib.vls_dwconv7_ln(ptr(x), B, H, W, ptr(wt), ptr(cb), ptr(lw), ptr(lb), 1e-6, ptr(out), stream()))
    nchw = x.view(B, H, W, 256).permute(0, 3, 1, 2)
    y = torch.nn.functional.conv2d(nchw, w, cb, padding=3, groups=256).permute(0, 2, 3, 1)
    ref = torch.nn.functional.layer_norm(y, (256,), lw, lb, 1e-6).reshape(B, H * W, 256)
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-2, err                       # bf16 output rounding of O(1) values
    assert (out.float() - ref).abs().mean().item() < 3e-3


@pytest.mark.parametrize("B,HW,n_mem,n_ptr,k", [(1, 4096, 7, 16, 4), (3, 64, 7, 16, 4), (2, 8, 3, 2, 4), (2, 16, 4, 21, 4), (1, 8, 2, 17, 2)])
def test_bank_shift_and_clone_many(dev, B, HW, n_mem, n_ptr, k):
    """Device memory bank: the one-launch in-place age shift equals the slice copies it replaces (exact, bf16 moves +
    one f32->bf16 rounding of the new pointer), and clone_many equals Tensor.clone()."""
    from video_llava_seg_b200 import ops

    Nk = n_mem * HW + n_ptr * k
    bank = _rand((B, Nk, 64), dev, 21).bfloat16()
    rows = _rand((B, HW, 64), dev, 22).bfloat16()
    new_ptr = _rand((B, k * 64), dev, 23)
    want = bank.clone()
    po = n_mem * HW
    want[:, HW:(n_mem - 1) * HW] = bank[:, 2 * HW:n_mem * HW]
    want[:, (n_mem - 1) * HW:n_mem * HW] = rows
    want[:, po + 2 * k:] = bank[:, po + k:po + (n_ptr - 1) * k]
    want[:, po + k:po + 2 * k] = new_ptr.reshape(B, k, 64).bfloat16()
    ops.bank_shift(bank, HW, n_mem, n_ptr, k, rows, new_ptr)
    assert torch.equal(bank, want)
    src = [rows, new_ptr, bank[:, :HW], torch.arange(5, device=dev), _rand((B, 1, 16, 16), dev, 24)]
    out = ops.clone_many(src)
    assert all(torch.equal(a, b) and a.data_ptr() != b.data_ptr() for a, b in zip(src, out))
    # copy_many: small unaligned tensors (bytes) and pairs of equally laid out dense views (channel-last) share the launch
    chl = _rand((1, 8, 8, 32), dev, 25).permute(0, 3, 1, 2)              # [1,32,8,8] view of NHWC memory
    tiny, odd = _rand((B, 1), dev, 26), _rand((7,), dev, 27)[1:]          # 4*B bytes; a 4-byte aligned, 24-byte tensor
    dsts = [torch.empty_like(chl), torch.empty_like(tiny), torch.empty_like(odd)]
    assert dsts[0].stride() == chl.stride()
    before = ops.lib().vls_launch_count()
    ops.copy_many([chl, tiny, odd], dsts)
    assert ops.lib().vls_launch_count() - before == 1, "one multi-copy launch, no fallback"
    assert all(torch.equal(a, b) for a, b in zip([chl, tiny, odd], dsts))


def test_cc_matches_reference_kernel(dev):
    """Pin: the reference's own connected_components.cu (compiled unmodified into oracle/_ref/ in the build
    container by oracle/build_ref.py) against our kernel and the C oracle, on the same masks."""
    from oracle import build_ref
    from oracle import cc as cc_oracle
    from video_llava_seg_b200.utils.misc import get_connected_components

    ref = build_ref.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref/ref_cc.so was not built (needs /root/reference in the build container)")
    g = torch.Generator().manual_seed(11)
    for shape, dens in (((4, 1, 256, 256), 0.58), ((2, 1, 64, 96), 0.45), ((1, 1, 512, 384), 0.6)):
        m = (torch.rand(shape, generator=g) < dens)
        rl, rc = ref.get_connected_componnets(m.to(torch.uint8).to(dev))
        ours_l, ours_c = get_connected_components(m.to(dev))
        torch.cuda.synchronize()
        assert torch.equal(ours_l, rl) and torch.equal(ours_c, rc)
        ol, oc = cc_oracle.cc_label(m)
        assert torch.equal(ol, rl.cpu()) and torch.equal(oc, rc.cpu())


@pytest.mark.parametrize("Nq,Nk,cluster", [(4096, 28736, 1), (4096, 4096, 1), (4096, 28736, 2)])
def test_attention_is_bitwise_deterministic(dev, vls_lib, Nq, Nk, cluster):
    """Race detector for the TMEM / mbarrier pipeline: same inputs and KV splits must give bit-identical outputs
    run after run, also with a cold L2 and other kernels interleaved (this caught a P/S aliasing hazard in r1)."""
    from video_llava_seg_b200 import ops

    vls_lib.vls_set_tuning(b"attn_cluster", cluster)
    try:
        q = _rand((1, Nq, 256), dev, 41).bfloat16()
        k = _rand((1, Nk, 256), dev, 42).bfloat16()
        vt = _rand((1, 256, Nk), dev, 43).bfloat16()
        flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
        base = ops.attention_d256(q, k, vt).clone()
        for it in range(120):
            if it % 3 == 0:
                flush.zero_()
            if it % 7 == 0:
                (q.float() @ k.float().transpose(1, 2)).sum()
            out = ops.attention_d256(q, k, vt)
            assert torch.equal(out, base), f"run {it} differs from run 0"
    finally:
        vls_lib.vls_set_tuning(b"attn_cluster", 1)
