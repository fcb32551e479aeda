"""Dev stress: the attention kernel must be bitwise deterministic run to run (same inputs, same splits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, ops
lib = _lib.lib(); dev = "cuda:0"
g = torch.Generator().manual_seed(0)
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
dbg_list = [int(a) for a in sys.argv[1:]] or [0]
for (Nq, Nk, cl, dbg) in [(4096, 4096, 1, d) for d in dbg_list] + [(4096, 28736, 1, d) for d in dbg_list]:
    if True:
        lib.vls_set_tuning(b"attn_cluster", cl)
            print("dbg", dbg, end=" ")
        q = torch.randn(1, Nq, 256, generator=g).to(dev).bfloat16()
        k = torch.randn(1, Nk, 256, generator=g).to(dev).bfloat16()
        ld = (Nk + 63) // 64 * 64
        vt = torch.zeros(1, 256, ld, device=dev, dtype=torch.bfloat16)
        vt[:, :, :Nk] = torch.randn(1, 256, Nk, generator=g).to(dev).bfloat16()
        ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), vt[:, :, :Nk].transpose(1, 2).float())
        base = ops.attention_d256(q, k, vt).clone()
        torch.cuda.synchronize()
        print(Nq, Nk, "cl", cl, "err vs torch", (base.float() - ref).abs().max().item(), flush=True)
        nbad = 0
        for it in range(int(os.environ.get('ITERS', '150'))):
            if it % 3 == 0:
                flush.zero_()
            if it % 7 == 0:
                torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), vt[:, :, :Nk].transpose(1, 2).float())
            out = ops.attention_d256(q, k, vt)
            if not torch.equal(out, base):
                nbad += 1
                d = (out.float() - base.float()).abs()
                rows = (d.amax(dim=2)[0] > 0).nonzero().flatten()
                print("  MISMATCH it", it, "max", d.max().item(), "rows", rows[:8].tolist(), "n rows", len(rows),
                      "tiles", sorted(set((rows // 128).tolist()))[:10], flush=True)
                if nbad > 4:
                    break
        print("  mismatches:", nbad)
