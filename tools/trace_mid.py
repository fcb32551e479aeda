"""Dev probe: phase timeline (SM clock cycles) of the memory-attention "mid" kernel for CTA (0,0,0), first epilogue thread."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, build_sam as B, synth
lib = _lib.lib()
dev = "cuda:0"
sd = synth.init_state_dict(0)
m = B.load_prefixed(B.build_memory_attention(), sd, "memory_attention.").to(dev).eval()
g = torch.Generator().manual_seed(0)
nq, nk = 4096, 7 * 4096 + 64
curr = (torch.randn(nq, 1, 256, generator=g) * 0.5).to(dev)
cpos = (torch.randn(nq, 1, 256, generator=g) * 0.5).to(dev)
mem = (torch.randn(nk, 1, 64, generator=g) * 0.5).bfloat16().to(dev)
mpos = (torch.randn(nk, 1, 64, generator=g) * 0.5).to(dev)
for _ in range(3):
    m(curr, mem, cpos, mpos, 64)
torch.cuda.synchronize()
lib.vls_set_tuning(b"tail_fused", 0)   # the layer tail shares the trace buffer: keep it out
lib.vls_set_tuning(b"ffn_fused", 0)
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lib.vls_ffn_trace(buf.data_ptr())
m(curr, mem, cpos, mpos, 64)
torch.cuda.synchronize()
lib.vls_ffn_trace(None)
st = buf.cpu().tolist()
names = ["start", "out-proj MMAs done", "residual slice landed", "x written, stats pushed", "cluster sync 1", "t all-gathered",
         "cluster sync 2", "q-proj MMAs done", "q stored", "end"]
prev = st[0]
for n, v in zip(names, st):
    if v:
        print(f"  {n:28s} +{v - st[0]:7d}  (d {v - prev:6d})")
        prev = v
