"""Dev probe: per-tile timeline (SM clock cycles) of the attention kernel's roles for CTA (0,0,0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, ops
lib = _lib.lib()
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
Nq, Nk = 4096, 28736
cl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib.vls_set_tuning(b"attn_cluster", cl)
q = torch.randn(1, Nq, 256, generator=g).to(dev).bfloat16()
k = torch.randn(1, Nk, 256, generator=g).to(dev).bfloat16()
vt = torch.randn(1, 256, Nk, generator=g).to(dev).bfloat16()
out = ops.attention_d256(q, k, vt)
torch.cuda.synchronize()
buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
lib.vls_attention_trace(buf.data_ptr())
ops.attention_d256(q, k, vt, out=out)
torch.cuda.synchronize()
lib.vls_attention_trace(None)
t = buf.cpu().view(3, 64, 8)
t0 = t[1, 0, 0].item()
print("tile | producer: kempty vempty | mma: kfull S_issued pready vfull PV_issued | softmax: sfull ld+max barrier exp st arrive   (cycles since first k_full)")
for j in list(range(0, 6)) + list(range(20, 30)):
    pr = [(x.item() - t0) for x in t[0, j, :2]]
    mm = [(x.item() - t0) for x in t[1, j, :5]]
    sm = [(x.item() - t0) for x in t[2, j, :6]]
    print(j, pr, mm, sm)
per = (t[2, 40, 5] - t[2, 20, 5]).item() / 20
print("steady-state period per tile (cycles):", per)
sm = t[2, 20:40, :6]
print("softmax stage means: wait s_full->ld/max", (sm[:, 1] - sm[:, 0]).float().mean().item(), "barrier", (sm[:, 2] - sm[:, 1]).float().mean().item(),
      "rescale+exp", (sm[:, 3] - sm[:, 2]).float().mean().item(), "st", (sm[:, 4] - sm[:, 3]).float().mean().item(), "arrive", (sm[:, 5] - sm[:, 4]).float().mean().item(),
      "idle until next s_full", (t[2, 21:41, 0] - t[2, 20:40, 5]).float().mean().item())
mm = t[1, 20:40, :5]
print("mma: S issue", (mm[:, 1] - mm[:, 0]).float().mean().item(), " wait P after S issue(j+1)->pready(j)", (t[1, 20:40, 2] - t[1, 21:41, 1]).float().mean().item(),
      "wait vfull", (mm[:, 3] - mm[:, 2]).float().mean().item(), "PV issue", (mm[:, 4] - mm[:, 3]).float().mean().item())
