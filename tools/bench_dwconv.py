"""Dev probe: the depth-wise 7x7 + LayerNorm2d strip kernels (vls_set_tuning dwconv_tma = 0 / 1 / 2) on the > L2 stress shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, synth
from video_llava_seg_b200._lib import check, ptr, stream

dev = torch.device("cuda:0")
lib = _lib.lib()
g = torch.Generator().manual_seed(0)
sd = synth.init_state_dict(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.randn(B, 4096, 256, generator=g).to(dev)
dw_w = sd["memory_encoder.fuser.layers.0.dwconv.weight"].reshape(256, 49).t().contiguous().to(dev)
dw_b = sd["memory_encoder.fuser.layers.0.dwconv.bias"].to(dev)
ln_w, ln_b = sd["memory_encoder.fuser.layers.0.norm.weight"].to(dev), sd["memory_encoder.fuser.layers.0.norm.bias"].to(dev)
out = torch.empty(B, 4096, 256, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
outs = {}
for variant in (0, 1, 2):
    check(lib.vls_set_tuning(b"dwconv_tma", variant))
    fn = lambda: check(lib.vls_dwconv7_ln(ptr(x), B, 64, 64, ptr(dw_w), ptr(dw_b), ptr(ln_w), ptr(ln_b), 1e-6, ptr(out), stream()))
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    ms = ts[len(ts) // 2]
    outs[variant] = out.float().clone()
    print(f"variant {variant}: {ms * 1e3:.1f} us  {B * 4096 * 256 * 6 / ms / 1e6:.0f} GB/s  max |diff vs variant 0| {(outs[variant] - outs[0]).abs().max().item():.3e}")
