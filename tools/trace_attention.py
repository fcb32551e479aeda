"""Dev probe: per-tile timeline (SM clock cycles) of the attention kernel's roles for CTA (0,0,0).
usage: trace_attention.py [dv=64|256]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, ops
lib = _lib.lib()
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
Nq, Nk = 4096, 28736
for dv in [int(a) for a in sys.argv[1:]] or [64, 256]:
    q = torch.randn(1, Nq, 256, generator=g).to(dev).bfloat16()
    k = torch.randn(1, Nk, 256, generator=g).to(dev).bfloat16()
    v = torch.randn(1, Nk, dv, generator=g).to(dev).bfloat16()
    vt = v.transpose(1, 2).contiguous()
    call = (lambda out=None: ops.attention_qk256(q, k, v, True, splits=4, out=out)) if dv == 64 else \
        (lambda out=None: ops.attention_qk256(q, k, vt, False, splits=4, out=out))
    out = call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        call(out)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 100
    buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
    lib.vls_attention_trace(buf.data_ptr())
    call(out)
    torch.cuda.synchronize()
    lib.vls_attention_trace(None)
    t = buf.cpu().view(3, 64, 8)
    t0 = t[1, 0, 0].item()
    print(f"==== dv={dv}: {us:.1f} us per launch (fixed 4-way split, 56-57 tiles per CTA)")
    print("tile | producer: K_issue V_issue | mma: kfull S_issued pready vfull PV_issued | softmax: sfull max barrier exp st arrive")
    for j in list(range(0, 4)) + list(range(20, 26)):
        pr = [(x.item() - t0) for x in t[0, j, :2]]
        mm = [(x.item() - t0) for x in t[1, j, :5]]
        sm = [(x.item() - t0) for x in t[2, j, :6]]
        print(j, pr, mm, sm)
    lo, hi = 10, 50
    per = (t[2, hi, 5] - t[2, lo, 5]).item() / (hi - lo)
    print(f"steady-state period per tile: {per:.0f} cycles; implied SM clock {per * 56.1 / us / 1e3:.2f} GHz if the launch were all steady state")
    sm = t[2, lo:hi, :6]
    f = lambda x: f"{x.float().mean().item():.0f}"
    print("softmax: s_full->max", f(sm[:, 1] - sm[:, 0]), "| barrier", f(sm[:, 2] - sm[:, 1]), "| rescale+exp", f(sm[:, 3] - sm[:, 2]),
          "| st", f(sm[:, 4] - sm[:, 3]), "| arrive", f(sm[:, 5] - sm[:, 4]), "| idle until next s_full", f(t[2, lo + 1:hi + 1, 0] - t[2, lo:hi, 5]))
    mm = t[1, lo:hi, :5]
    print("mma: wait k_full(j+1) after PV(j-1) issued", f(t[1, lo + 1:hi + 1, 0] - t[1, lo - 1:hi - 1, 4]), "| S issue", f(mm[:, 1] - mm[:, 0]),
          "| S_issued(j+1)->p_ready(j)", f(t[1, lo:hi, 2] - t[1, lo + 1:hi + 1, 1]), "| wait v_full", f(mm[:, 3] - mm[:, 2]),
          "| PV issue", f(mm[:, 4] - mm[:, 3]))
    print("producer: K issue(j+1) - K issue(j)", f(t[0, lo + 1:hi + 1, 0] - t[0, lo:hi, 0]), "| k_full(j) - K issue(j) (load latency incl. queueing)",
          f(t[1, lo:hi, 0] - t[0, lo:hi, 0]), "| v_full(j) - V issue(j)", f(t[1, lo:hi, 3] - t[0, lo:hi, 1]))
    print("softmax s_full(j) - mma S_issued(j) (S MMA latency)", f(t[2, lo:hi, 0] - t[1, lo:hi, 1]))
