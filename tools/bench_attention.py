"""Dev probe: attention kernel timing (memory cross-attention + self-attention shapes) for each variant:
dv=256 (r1 kernel, V^T), dv=64 with V as bank rows (MN-major operand) and dv=64 with a transposed copy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, ops
lib = _lib.lib()
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(B, Nq, Nk, dv, v_rows, iters=20, splits=0):
    q = (torch.randn(B, Nq, 256, generator=g)).to(dev).bfloat16()
    k = (torch.randn(B, Nk, 256, generator=g)).to(dev).bfloat16()
    v = torch.randn(B, Nk, dv, generator=g).to(dev).bfloat16()
    ld = (Nk + 63) // 64 * 64
    vt = torch.zeros(B, dv, ld, device=dev, dtype=torch.bfloat16)
    vt[:, :, :Nk] = v.transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(q[:1].float(), k[:1].float(), v[:1].float())
    call = (lambda out=None: ops.attention_qk256(q, k, v, True, out=out, splits=splits)) if v_rows else \
        (lambda out=None: ops.attention_qk256(q, k, vt, False, out=out, splits=splits))
    out = call()
    err = (out[:1].float() - ref).abs().max().item()
    for _ in range(3):
        call(out)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_()
        a.record()
        call(out)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    ms = ts[len(ts) // 2]
    print(f"B={B} Nq={Nq} Nk={Nk} dv={dv} v_rows={int(v_rows)} splits={splits}: median {ms*1e3:.1f} us min {ts[0]*1e3:.1f} "
          f"(incl. combine) ref-algorithmic {4*B*Nq*Nk*256/ms/1e9:.0f} TFLOP/s executed "
          f"{2*B*Nq*Nk*(256+dv)/ms/1e9:.0f} TFLOP/s  max err {err:.2e}", flush=True)


run(1, 4096, 28736, 256, False)
run(1, 4096, 28736, 64, True)
run(1, 4096, 28736, 64, False)
run(1, 4096, 28736, 64, True, splits=4)
run(1, 4096, 4096, 256, False)
run(8, 4096, 28736, 256, False, iters=6)
run(8, 4096, 28736, 64, True, iters=6)
run(8, 4096, 28736, 64, True, iters=6, splits=-1)
