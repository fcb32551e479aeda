"""Dev probe: per-phase clock64 trace of cc_small_kernel (library built with VLS_EXTRA_NVCC_FLAGS=-DCC_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib
from video_llava_seg_b200._lib import check, ptr, stream

dev = torch.device("cuda:0")
lib = _lib.lib()
g = torch.Generator().manual_seed(0)
def blobby(n, h, w):
    z = torch.randn(n, 1, h // 8, w // 8, generator=g)
    z = torch.nn.functional.interpolate(z, size=(h, w), mode="bilinear", align_corners=False)
    return (z + 0.15 * torch.randn(n, 1, h, w, generator=g)) > 0
for kind, n in (("blobby", 1), ("blobby", 512), ("noise", 512)):
    m = (blobby(n, 256, 256) if kind == "blobby" else torch.rand(n, 1, 256, 256, generator=g) < 0.55).to(dev).to(torch.uint8)
    labels = torch.empty((n, 1, 256, 256), dtype=torch.int32, device=dev)
    counts = torch.empty_like(labels)
    print(f"--- cc_label {kind} N={n}", flush=True)
    check(lib.vls_cc_label(ptr(m), n, 256, 256, ptr(labels), ptr(counts), None, 0, stream()))
    torch.cuda.synchronize()
s = (torch.nn.functional.avg_pool2d(torch.randn(256, 1, 256, 256, generator=g), 5, 1, 2) * 3).to(dev)
print("--- fill_holes N=256", flush=True)
check(lib.vls_fill_holes(ptr(s), 256, 256, 256, 8, 0.1, None, 0, stream()))
torch.cuda.synchronize()
