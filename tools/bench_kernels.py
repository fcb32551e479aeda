"""Roofline micro-benchmark of the bandwidth-bound kernels at shapes larger than L2 (126 MB), timed with CUDA
events through the C ABI; achieved GB/s = ALGORITHMIC bytes / time, against MEASURED_PEAKS.json hbm_gbs.
    python tools/bench_kernels.py [--json out.json]
With `--once` every kernel is launched a single time (for `ncu --set full`)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, build_sam, synth
from video_llava_seg_b200._lib import check, ptr, stream
from video_llava_seg_b200.utils.misc import fill_holes_in_mask_scores, get_connected_components

once = "--once" in sys.argv
dev = torch.device("cuda:0")
lib = _lib.lib()
peak = 6550.1
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
results = []


def timeit(name, fn, alg_bytes, iters=10):
    if once:
        fn(); torch.cuda.synchronize(); return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    results.append(dict(kernel=name, ms=round(ms, 4), alg_MB=round(alg_bytes / 1e6, 1), GBps=round(gbs, 1),
                        frac_of_hbm_peak=round(gbs / peak, 3)))
    print(results[-1], flush=True)


g = torch.Generator().manual_seed(0)
# ---- connected components: 9 B/pixel (1 in + 4 labels + 4 areas)
def blobby(n, h, w, thr=0.0):
    z = torch.randn(n, 1, h // 8, w // 8, generator=g)
    z = torch.nn.functional.interpolate(z, size=(h, w), mode="bilinear", align_corners=False)
    return (z + 0.15 * torch.randn(n, 1, h, w, generator=g)) > thr
for (n, h, w, kind) in ((64, 1024, 1024, "blobby"), (64, 1024, 1024, "noise"), (512, 256, 256, "blobby")):
    m = (blobby(n, h, w) if kind == "blobby" else torch.rand(n, 1, h, w, generator=g) < 0.55).to(dev).to(torch.uint8)
    nb = lib.vls_cc_workspace_bytes(n, h, w)
    ws = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
    labels = torch.empty((n, 1, h, w), dtype=torch.int32, device=dev)
    counts = torch.empty_like(labels)
    timeit(f"cc_label N={n} {h}x{w} {kind}",
           lambda: check(lib.vls_cc_label(ptr(m), n, h, w, ptr(labels), ptr(counts), ptr(ws), nb, stream())), n * h * w * 9)
    del m, ws, labels, counts
# fused hole filling: reads 4 B/pixel, writes only changed pixels
s = (torch.nn.functional.avg_pool2d(torch.randn(256, 1, 256, 256, generator=g), 5, 1, 2) * 3).to(dev)
nb = lib.vls_fill_holes_workspace_bytes(256, 256, 256)
ws = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
timeit("fill_holes N=256 256x256", lambda: check(lib.vls_fill_holes(ptr(s), 256, 256, 256, 8, 0.1, ptr(ws), nb, stream())),
       256 * 256 * 256 * 4)
del s
# ---- CXBlock dwconv7x7 + LN2d: 1 KB in + 0.5 KB out per pixel
sd = synth.init_state_dict(0)
B = 64
x = torch.randn(B, 4096, 256, generator=g).to(dev)
dw_w = sd["memory_encoder.fuser.layers.0.dwconv.weight"].reshape(256, 49).t().contiguous().to(dev)
dw_b = sd["memory_encoder.fuser.layers.0.dwconv.bias"].to(dev)
ln_w, ln_b = sd["memory_encoder.fuser.layers.0.norm.weight"].to(dev), sd["memory_encoder.fuser.layers.0.norm.bias"].to(dev)
out = torch.empty(B, 4096, 256, dtype=torch.bfloat16, device=dev)
timeit(f"dwconv7_ln B={B} 64x64x256",
       lambda: check(lib.vls_dwconv7_ln(ptr(x), B, 64, 64, ptr(dw_w), ptr(dw_b), ptr(ln_w), ptr(ln_b), 1e-6, ptr(out), stream())),
       B * 4096 * 256 * 6)
timeit(f"layernorm256 rows={B * 4096}",
       lambda: check(lib.vls_layernorm256(ptr(x), B * 4096, ptr(ln_w), ptr(ln_b), 1e-5, 0, ptr(out), stream())),
       B * 4096 * 256 * 6)
del x, out
# ---- memory encoder end to end at B=16 (mask down-sampler from low-res logits + fuser + out_proj)
enc = build_sam.load_prefixed(build_sam.build_memory_encoder(), sd, "memory_encoder.").to(dev).eval()
Bm = 16
vf = torch.randn(4096, Bm, 256, generator=g).to(dev)
low = (torch.randn(Bm, 1, 256, 256, generator=g) * 2).to(dev)
gate = torch.zeros(Bm, device=dev)
with torch.inference_mode():
    timeit(f"mem_encoder (fused low-res path) B={Bm}",
           lambda: enc.encode_from_low_res(vf, low, False, 20.0, -10.0, gate, sd["no_obj_embed_spatial"]),
           Bm * (4096 * 256 * 4 + 256 * 256 * 4 + 2 * 4096 * 64 * 2))
if not once:
    for a in sys.argv:
        if a.endswith(".json"):
            json.dump(dict(hbm_peak_GBps=peak, results=results), open(a, "w"), indent=1)
