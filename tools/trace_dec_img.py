"""Dev probe: phase timeline (SM clock cycles) of the mask decoder's image-side cluster kernel (dec_img.cu), thread 64 of CTA (0,0,0);
the LAST dec_img launch of one decoder call is what remains in the buffer."""
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
emb = torch.randn(1, 256, 64, 64, generator=g).to(dev)
s0 = torch.randn(1, 32, 256, 256, generator=g).to(dev)
s1 = torch.randn(1, 64, 128, 128, generator=g).to(dev)
sparse = torch.randn(1, 2, 256, generator=g).to(dev)
dense = pe_mod.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(1, -1, 64, 64)
call = lambda: dec(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                   multimask_output=True, repeat_image=False, high_res_features=[s0, s1])
for _ in range(3):
    call()
torch.cuda.synchronize()
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lib.vls_ffn_trace(buf.data_ptr())
call()
torch.cuda.synchronize()
lib.vls_ffn_trace(None)
st = buf.cpu().tolist()
names = ["start", "params / token k,v staged", "attention tile written", "out-proj MMAs + residual ready", "epilogue 1 (stats pushed)",
         "cluster sync 1", "LN4, keys stored, t panel + bulk copies", "next-projection MMAs done", "planes stored", "end"]
prev = st[0]
for n, v in zip(names, st):
    if v:
        print(f"  {n:42s} +{v - st[0]:7d}  (d {v - prev:6d})")
        prev = v
