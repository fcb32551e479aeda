"""Dev probe: the two-query-tile memory cross-attention kernel (attn_x2.cu) against the one-tile kernel (attn_tc.cu):
launch time per KV split count with a cold L2, error against fp32 SDPA, and the clock64 trace of CTA (0,0,0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, ops
lib = _lib.lib()
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(B, Nq, Nk, x2, splits=0, iters=20, trace=False, poly=2):
    lib.vls_set_tuning(b"attn_x2", int(x2))
    lib.vls_set_tuning(b"attn_x2_poly", poly)
    q = (torch.randn(B, Nq, 256, generator=g)).to(dev).bfloat16()
    k = (torch.randn(B, Nk, 256, generator=g)).to(dev).bfloat16()
    v = torch.randn(B, Nk, 64, generator=g).to(dev).bfloat16()
    ref = torch.nn.functional.scaled_dot_product_attention(q[:1].float(), k[:1].float(), v[:1].float())
    call = lambda out=None: ops.attention_qk256(q, k, v, True, out=out, splits=splits)
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
    print(f"x2={int(x2)} poly={poly} B={B} Nq={Nq} Nk={Nk} splits={splits}: median {ms*1e3:.1f} us min {ts[0]*1e3:.1f} (incl. combine) "
          f"ref-algorithmic {4*B*Nq*Nk*256/ms/1e9:.0f} TFLOP/s executed {2*B*Nq*Nk*320/ms/1e9:.0f} TFLOP/s  max err {err:.2e}",
          flush=True)
    if trace and x2:
        buf = torch.zeros(4 * 48 * 8, dtype=torch.int64, device=dev)
        lib.vls_attention_trace(buf.data_ptr())
        call(out)
        torch.cuda.synchronize()
        lib.vls_attention_trace(None)
        t = buf.cpu().view(4, 48, 8)
        for sp in range(16):
            c0, n0, c1, n1 = [int(x) for x in t[3, 32 + sp, 4:8]]
            if n1 > n0:
                print(f"  CTA (0,{sp},0): {c1 - c0} cycles in {(n1 - n0) / 1e3:.1f} us -> SM clock {(c1 - c0) / (n1 - n0):.3f} GHz; "
                      f"start {(n0 - int(t[3, 32, 5])) / 1e3:+.1f} us after CTA (0,0,0)")
        t0 = t[1, 0, 0].item()
        n = int((t[2, :, 0] != 0).sum())
        print(f"  trace of CTA (0,0,0): {n} key tiles; columns relative to the first S issue")
        print("  tile | K issue, V issue | mma: p_ready(A) PV_A issued S_A(j+1) issued p_ready(B) PV_B issued S_B(j+1) issued | "
              "softmax A: s_full max exp+st arrive | softmax B: ...")
        for j in list(range(0, 3)) + list(range(max(3, n // 2), min(n, n // 2 + 4))):
            row = [t[0, j, :2], t[1, j, 2:8], t[2, j, :4], t[3, j, :4]]
            print("  ", j, *[[int(x) - t0 for x in r] for r in row])
        lo, hi = 4, n - 2
        if hi > lo + 2:
            per = (t[2, hi, 3] - t[2, lo, 3]).item() / (hi - lo)
            f = lambda x: f"{x.float().mean().item():.0f}"
            print(f"  steady-state period per key tile (both query tiles): {per:.0f} cycles")
            for gi, nm in ((2, "A"), (3, "B")):
                s = t[gi, lo:hi]
                print(f"  softmax {nm}: ld+max {f(s[:, 1] - s[:, 0])} | exp+st {f(s[:, 2] - s[:, 1])} | wait st+arrive {f(s[:, 3] - s[:, 2])}"
                      f" | idle until next s_full {f(t[gi, lo + 1:hi + 1, 0] - s[:, 3])}")
            m = t[1, lo:hi]
            print(f"  mma: wait p_ready(A) after S_B issued {f(m[:, 2] - t[1, lo - 1:hi - 1, 7])} | PV_A issue {f(m[:, 3] - m[:, 2])} | "
                  f"S_A issue {f(m[:, 4] - m[:, 3])} | wait p_ready(B) {f(m[:, 5] - m[:, 4])} | PV_B issue {f(m[:, 6] - m[:, 5])} | "
                  f"S_B issue {f(m[:, 7] - m[:, 6])}")
            sa, sb = t[2, lo:hi, 0], t[3, lo:hi, 0]            # S_A(j), S_B(j) complete (seen by the softmax threads)
            pa, pb = t[0, lo:hi, 2], t[0, lo:hi, 3]            # PV_A(j), PV_B(j) complete
            print(f"  pipe: S_A(j) done -> S_B(j) done {f(sb - sa)} | S_B(j) done -> PV_A(j) done {f(pa - sb)} | PV_A(j) -> S_A(j+1) done "
                  f"{f(t[2, lo + 1:hi + 1, 0] - pa)} | S_A(j+1) -> PV_B(j) done {f(pb - t[2, lo + 1:hi + 1, 0])} | PV_B(j) -> S_B(j+1) done "
                  f"{f(t[3, lo + 1:hi + 1, 0] - pb)}")
            print(f"  PV_A(j) done - p_ready(A) seen {f(pa - m[:, 2])} | PV_B(j) done - p_ready(B) seen {f(pb - m[:, 5])}")
            print(f"  K tile issue interval {f(t[0, lo + 1:hi + 1, 0] - t[0, lo:hi, 0])}")


mode = sys.argv[1] if len(sys.argv) > 1 else "all"
run(1, 4096, 28736, False)
for pl in (0, 1, 2, 3):
    run(1, 4096, 28736, True, trace=True, poly=pl)
if mode == "all":
    for s in (6, 8, 9, 10, 12):
        run(1, 4096, 28736, True, splits=s)
    run(8, 4096, 28736, False, iters=6)
    run(8, 4096, 28736, True, iters=6)
    run(8, 4096, 28736, True, iters=6, splits=8)
    run(2, 4096, 28736, True, iters=10)
    run(1, 4096, 12352, True)
    run(1, 4096, 4100, True)
