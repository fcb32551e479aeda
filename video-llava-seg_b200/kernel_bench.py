"""Roofline micro-benchmark of the bandwidth-bound kernels at shapes larger than L2 (126 MB), timed with CUDA events
through the C ABI; achieved GB/s = ALGORITHMIC bytes / time against the measured HBM copy bandwidth.  Used by bench.py
(`roofline_hbm`, rank 0, N = 1) and by tools/bench_kernels.py (which adds `--once` for ncu captures)."""
import torch

from . import _lib, synth
from ._lib import check, ptr, stream


def _blobby(g, n, h, w, thr=0.0):
    z = torch.randn(n, 1, h // 8, w // 8, generator=g)
    z = torch.nn.functional.interpolate(z, size=(h, w), mode="bilinear", align_corners=False)
    return (z + 0.15 * torch.randn(n, 1, h, w, generator=g)) > thr


def cases(dev, quick=True):
    """Yields (name, launch closure, algorithmic bytes).  quick: the three shapes the north-star names (CC production
    shape, CC stress shape, fuser dwconv) + LayerNorm; otherwise also noise input / hole filling."""
    lib = _lib.lib()
    g = torch.Generator().manual_seed(0)
    # 592 = 148 SMs x 2 resident CTAs x 2: whole waves of the one-CTA-per-image kernel (512 images are 1.73 waves)
    shapes = [(512, 256, 256, "blobby"), (592, 256, 256, "blobby"), (64, 1024, 1024, "blobby")] + \
        ([] if quick else [(64, 1024, 1024, "noise")])
    for (n, h, w, kind) in shapes:
        m = (_blobby(g, n, h, w) if kind == "blobby" else torch.rand(n, 1, h, w, generator=g) < 0.55).to(dev).to(torch.uint8)
        nb = lib.vls_cc_workspace_bytes(n, h, w)
        ws = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
        labels = torch.empty((n, 1, h, w), dtype=torch.int32, device=dev)
        counts = torch.empty_like(labels)
        # connected components: 9 B/pixel (1 in + 4 labels + 4 areas)
        yield (f"cc_label N={n} {h}x{w} {kind}",
               lambda m=m, n=n, h=h, w=w, labels=labels, counts=counts, ws=ws, nb=nb: check(
                   lib.vls_cc_label(ptr(m), n, h, w, ptr(labels), ptr(counts), ptr(ws), nb, stream())), n * h * w * 9)
    if not quick:
        s = (torch.nn.functional.avg_pool2d(torch.randn(256, 1, 256, 256, generator=g), 5, 1, 2) * 3).to(dev)
        nb = lib.vls_fill_holes_workspace_bytes(256, 256, 256)
        ws = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
        yield ("fill_holes N=256 256x256",
               lambda: check(lib.vls_fill_holes(ptr(s), 256, 256, 256, 8, 0.1, ptr(ws), nb, stream())), 256 * 256 * 256 * 4)
    # CXBlock dwconv7x7 + LN2d: 1 KB in + 0.5 KB out per pixel
    sd = synth.init_state_dict(0)
    B = 64
    x = torch.randn(B, 4096, 256, generator=g).to(dev)
    dw_w = sd["memory_encoder.fuser.layers.0.dwconv.weight"].reshape(256, 49).t().contiguous().to(dev)
    dw_b = sd["memory_encoder.fuser.layers.0.dwconv.bias"].to(dev)
    ln_w, ln_b = sd["memory_encoder.fuser.layers.0.norm.weight"].to(dev), sd["memory_encoder.fuser.layers.0.norm.bias"].to(dev)
    out = torch.empty(B, 4096, 256, dtype=torch.bfloat16, device=dev)
    yield (f"dwconv7_ln B={B} 64x64x256",
           lambda: check(lib.vls_dwconv7_ln(ptr(x), B, 64, 64, ptr(dw_w), ptr(dw_b), ptr(ln_w), ptr(ln_b), 1e-6, ptr(out),
                                            stream())), B * 4096 * 256 * 6)
    yield (f"layernorm256 rows={B * 4096}",
           lambda: check(lib.vls_layernorm256(ptr(x), B * 4096, ptr(ln_w), ptr(ln_b), 1e-5, 0, ptr(out), stream())),
           B * 4096 * 256 * 6)


def hbm_rooflines(dev, peak_gbs, quick=True, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    results = []
    for name, fn, alg_bytes in cases(dev, quick):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        results.append({"bound": "hbm", "kernel": name, "ms": round(ms, 4), "alg_MB": round(alg_bytes / 1e6, 1),
                        "achieved": round(gbs, 1), "peak": peak_gbs, "unit": "GB/s", "frac": round(gbs / peak_gbs, 3)})
    return results
