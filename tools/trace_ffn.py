"""Dev probe: phase timeline (SM clock cycles) of the fused FFN / layer-tail kernel for CTA (0,0,0), first epilogue thread."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, ops
lib = _lib.lib()
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
B, M = 1, 4096
t = rn(B, M, 256).bfloat16()
ao = rn(B, M, 64).bfloat16()
w0, b0 = rn(256, 64, sc=1 / 8).bfloat16(), rn(256, sc=0.1)
lw, lb = 1 + rn(256, sc=0.1), rn(256, sc=0.05)
w1, b1 = rn(2048, 256, sc=1 / 16).bfloat16(), rn(2048, sc=0.1)
w2, b2 = rn(256, 2048, sc=1 / 45).bfloat16(), rn(256, sc=0.1)
x = rn(B, M, 256)
names = ["start", "g0_done seen", "pass1 done (x_mid, stats)", "pass2 done (t -> smem)", "d1_full(0)", "d1_full(1)", "d1_full(2)",
         "d1_full(3)", "h(3) written", "y_full", "cluster sync A", "partials pushed", "cluster sync B", "reduced + x written",
         "t_out written", "end"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for which in ("ffn", "tail"):
    call = (lambda: ops.ffn_fused(t, w1, b1, w2, b2, x)) if which == "ffn" else \
        (lambda: ops.mem_attn_layer_tail(ao, w0, b0, lw, lb, w1, b1, w2, b2, x, lw, lb))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ev = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    us = sorted(a.elapsed_time(b) for a, b in ev)[5] * 1e3
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.vls_ffn_trace(buf.data_ptr())
    call()
    torch.cuda.synchronize()
    lib.vls_ffn_trace(None)
    st = buf.cpu().tolist()
    print(f"==== {which}: median {us:.1f} us (cold L2)")
    prev = st[0]
    for n, v in zip(names, st):
        if v:
            print(f"  {n:28s} +{v - st[0]:7d}  (d {v - prev:6d})")
            prev = v
