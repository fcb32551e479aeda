"""Dev probe: attention kernel timing (cross + self shapes) for each tuning variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, ops
lib = _lib.lib()
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
def run(Nq, Nk, cl, iters=20):
    lib.vls_set_tuning(b"attn_cluster", cl)
    q = (torch.randn(1, Nq, 256, generator=g)).to(dev).bfloat16()
    k = (torch.randn(1, Nk, 256, generator=g)).to(dev).bfloat16()
    ld = (Nk + 63) // 64 * 64
    vt = torch.randn(1, 256, ld, generator=g).to(dev).bfloat16()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), vt[:, :, :Nk].transpose(1, 2).float())
    out = ops.attention_d256(q, k, vt)
    err = (out.float() - ref).abs().max().item()
    for _ in range(3): ops.attention_d256(q, k, vt, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): ops.attention_d256(q, k, vt, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    print(f"Nq={Nq} Nk={Nk} cluster={cl}: {ms*1e3:.1f} us (incl. combine) {4*Nq*Nk*256/ms/1e9:.0f} TFLOP/s  max err {err:.2e}", flush=True)
for dbg in [int(a) for a in sys.argv[1:]] or [0]:
    print("dbg", dbg)
    for cl in (1, 2):
        run(4096, 28736, cl)
        run(4096, 4096, cl)
