"""The LLaVA-side caller of the mask decoder (llava/model/seg_head/sam2.py: SegmentationHeadSAM2),
re-hosted on libvls_b200: `[SEG]` hidden states -> proj_token -> sparse prompt -> per-frame
MaskDecoder (repeat_image=True, multimask_output=False) -> max over the Q queries -> resize.
The image encoder is injected (any module returning the reference feature dict) or bypassed by passing
precomputed `backbone_feats` / `high_res_feats`."""
import torch
from torch import nn

from . import ops


class SegmentationHeadSAM2(nn.Module):
    def __init__(self, n_token_dims, n_seg_queries, sam2_model):
        """sam2_model: a video_llava_seg_b200 SAM2Base/SAM2VideoPredictor (supplies prompt encoder, mask decoder,
        no_mem_embed and, optionally, image_encoder) -- what the reference takes from
        SAM2ImagePredictor.from_pretrained(...).model (sam2.py:15-24)."""
        super().__init__()
        self.n_seg_queries = n_seg_queries
        self.proj_token = nn.Linear(n_token_dims, 256 * n_seg_queries)
        self.sam2 = sam2_model
        self._w = None
        self.register_load_state_dict_post_hook(lambda m, keys: setattr(m, "_w", None))

    def _apply(self, fn, *a, **kw):
        self._w = None
        return super()._apply(fn, *a, **kw)

    def project_tokens(self, seg_tokens):
        """[M, n_token_dims] -> [M*Q, 1, 256] sparse prompt embeddings (sam2.py:74-76,88)."""
        if self._w is None:
            self._w = (self.proj_token.weight.detach().to(torch.bfloat16).contiguous(),
                       self.proj_token.bias.detach().float().contiguous())
        y = ops.linear_f32(seg_tokens.float(), self._w[0], self._w[1])           # [M, Q*256]
        return y.reshape(-1, self.n_seg_queries, 256).reshape(-1, 1, 256)

    @torch.inference_mode()
    def decode(self, backbone_feats, high_res_feats, seg_tokens, out_size=None):
        """backbone_feats [T,256,H,W] (WITHOUT no_mem_embed), high_res_feats ([T,32,4H,4W], [T,64,2H,2W]) already
        through conv_s0/conv_s1, seg_tokens [M,n_token_dims] -> mask logits [M,T,h,w] (sam2.py:96-131)."""
        m = self.sam2
        T = backbone_feats.shape[0]
        sparse = self.project_tokens(seg_tokens)                                   # [M*Q,1,256]
        n = sparse.shape[0]
        pe = m.sam_prompt_encoder
        dense = pe.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(n, -1, *pe.image_embedding_size)
        image_pe = pe.get_dense_pe()
        H, W = backbone_feats.shape[-2:]
        rows = backbone_feats.flatten(2).permute(2, 0, 1)                          # [HW,T,256]
        feats = ops.add_rowvec(rows, m.no_mem_embed.detach().reshape(-1))          # + no_mem_embed (sam2.py:44)
        feats = feats.permute(1, 2, 0).reshape(T, 256, H, W)
        out = []
        for t in range(T):
            masks, _, _, _ = m.sam_mask_decoder(
                image_embeddings=feats[t:t + 1], image_pe=image_pe, sparse_prompt_embeddings=sparse,
                dense_prompt_embeddings=dense, multimask_output=False, repeat_image=True,
                high_res_features=[high_res_feats[0][t:t + 1], high_res_feats[1][t:t + 1]])
            if out_size is not None:
                masks = ops.resize_bilinear(masks, out_size)
            out.append(masks)
        masks = torch.cat(out, 1)                                                  # [M*Q,T,h,w]
        masks = masks.reshape(-1, self.n_seg_queries, *masks.shape[1:])
        return masks.max(1).values                                                 # max over the Q queries (sam2.py:127-128)
