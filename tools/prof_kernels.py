"""Dev probe for `ncu --set full`: launches the fused FFN kernel and the dv=64 cross-attention kernel at the steady-state
shapes a few times (usage: python tools/prof_kernels.py [ffn|attn|both])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import ops
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
what = sys.argv[1] if len(sys.argv) > 1 else "both"
if what in ("ffn", "both"):
    t = rn(1, 4096, 256).bfloat16()
    w1, b1 = rn(2048, 256, sc=1 / 16).bfloat16(), rn(2048, sc=0.1)
    w2, b2 = rn(256, 2048, sc=1 / 45).bfloat16(), rn(256, sc=0.1)
    x = rn(1, 4096, 256)
    for _ in range(4):
        ops.ffn_fused(t, w1, b1, w2, b2, x)
if what in ("attn", "both"):
    q, k, v = rn(1, 4096, 256).bfloat16(), rn(1, 28736, 256).bfloat16(), rn(1, 28736, 64).bfloat16()
    for _ in range(4):
        ops.attention_qk256(q, k, v, True)
torch.cuda.synchronize()
print("done")
