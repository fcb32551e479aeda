"""Dev probe: stage timeline (SM clock cycles) of the mask decoder's token-side cluster kernel (dec_tok.cu), thread 0 of CTA (0,0),
for the five launches of one decoder call (SELF / CROSS+MLP per layer, final CROSS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import _lib, build_sam as B, synth
lib = _lib.lib()
dev = "cuda:0"
sd = synth.init_state_dict(0)
dec = B.load_prefixed(B.build_mask_decoder(), sd, "sam_mask_decoder.").to(dev).eval()
pe_mod = B.load_prefixed(B.build_prompt_encoder(), sd, "sam_prompt_encoder.").to(dev).eval()
pe = pe_mod.get_dense_pe()
g = torch.Generator().manual_seed(0)
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 1
emb = torch.randn(Bn, 256, 64, 64, generator=g).to(dev)
s0 = torch.randn(Bn, 32, 256, 256, generator=g).to(dev)
s1 = torch.randn(Bn, 64, 128, 128, generator=g).to(dev)
sparse = torch.randn(Bn, 2, 256, generator=g).to(dev)
dense = pe_mod.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(Bn, -1, 64, 64)
call = lambda: dec(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                   multimask_output=True, repeat_image=False, high_res_features=[s0, s1])
for _ in range(3):
    call()
torch.cuda.synchronize()
buf = torch.zeros(24 * 8, dtype=torch.int64, device=dev)
lib.vls_dec_trace(buf.data_ptr())
call()
torch.cuda.synchronize()
lib.vls_dec_trace(None)
st = buf.cpu().view(8, 24).tolist()
names = {0: "start", 1: "rows + operands in smem", 2: "qkv tiles", 3: "self-attn + bcast", 4: "cluster sync 1", 5: "o-proj + bcast",
         6: "cluster sync 2", 7: "LN1 (+ops/store)", 8: "q tile", 9: "t2i attention", 10: "merge + bcast", 11: "cluster sync 3",
         12: "o-proj + bcast", 13: "cluster sync 4", 14: "LN2 (+ops/store)", 15: "mlp1 tiles", 16: "mlp2 tiles + push",
         17: "cluster sync 5", 18: "reduce + bcast", 19: "cluster sync 6", 20: "LN3, store, i2t k/v"}
for k, row in enumerate(st):
    if not any(row):
        continue
    print(f"==== launch {k}")
    t0 = prev = row[0]
    print(f"  globaltimer: {row[23] - row[22]} ns for {row[21] - row[0]} cycles -> {(row[21] - row[0]) / max(row[23] - row[22], 1):.3f} GHz")
    for i, v in enumerate(row[:21]):
        if v:
            print(f"  {names[i]:28s} +{v - t0:7d}  (d {v - prev:6d})")
            prev = v

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        call()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev[len(ev) * 2 // 3:]:
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name[:70]}")
