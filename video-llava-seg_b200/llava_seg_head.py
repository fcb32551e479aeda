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

    # ------------------------------------------------------------------ reference entry point
    @property
    def has_image_encoder(self):
        return getattr(self.sam2, "image_encoder", None) is not None

    def encode_video_frames(self, video_frames):
        """[T,3,1024,1024] RGB in [0,1] -> (backbone_feats [T,256,64,64] WITHOUT no_mem_embed, [feat_s0, feat_s1])
        through the injected image encoder (sam2.py:34-47; the encoder itself is SURVEY row f-4)."""
        if not self.has_image_encoder:
            raise RuntimeError("this SAM2 model was built without an image encoder: pass backbone_features=[(feats, "
                               "[feat_s0, feat_s1]), ...] to forward(), or build the model with one")
        mean = torch.tensor([0.485, 0.456, 0.406], device=video_frames.device)[None, :, None, None]
        std = torch.tensor([0.229, 0.224, 0.225], device=video_frames.device)[None, :, None, None]
        out = self.sam2.forward_image((video_frames.float() - mean) / std)       # conv_s0 / conv_s1 already applied
        fpn = out["backbone_fpn"]
        return fpn[2], [fpn[0], fpn[1]]

    @torch.inference_mode()
    def forward(self, video_frames, seg_tokens, seg_meta, resize_to_original_dims, **kwargs):
        """SegmentationHeadSAM2.forward (llava/model/seg_head/sam2.py:49-131), same arguments and result:
        video_frames: list (batch) of [T,3,H,W] RGB tensors in [0,1]; seg_tokens: list of [M,n_token_dims] `[SEG]` hidden
        states; seg_meta: list of dicts with `padding`, `resized_image_size`, `orig_image_size`.
        Returns a list of [M,T,H',W'] mask logits (padding removed; resized to the original size if asked).
        Extra keyword `backbone_features`: list of (backbone_feats [T,256,64,64], [feat_s0, feat_s1]) per sample, which
        bypasses the image encoder (precomputed features)."""
        pre = kwargs.get("backbone_features")
        outs = []
        for i, (tok, meta) in enumerate(zip(seg_tokens, seg_meta)):
            feats, high = pre[i] if pre is not None else self.encode_video_frames(video_frames[i])
            masks = self.decode(feats, high, tok, reduce_queries=False)              # [M*Q,T,256,256]
            masks = self.postprocess_masks(masks, meta, resize_to_original_dims)     # per query, as the reference (:116-120)
            masks = masks.reshape(-1, self.n_seg_queries, *masks.shape[1:])
            outs.append(masks.max(1).values)                                         # max over the Q queries AFTER the resize (:127-128)
        return outs

    def postprocess_masks(self, masks, meta_dict, resize_to_original_dims):
        """Up-sample to the 1024^2 model input, remove the padding, optionally resize to the original image size
        (sam2.py:133-182).  masks: [M,T,h,w] logits."""
        masks = ops.resize_bilinear(masks.float(), (1024, 1024))
        left, right, top, bottom = [int(p) for p in meta_dict["padding"]]            # F.pad order: last dim first
        H, W = masks.shape[-2:]
        masks = masks[..., top:H - bottom, left:W - right]
        assert list(masks.shape[-2:]) == list(meta_dict["resized_image_size"]), \
            f"Shape mismatch: {masks.shape}, {meta_dict['resized_image_size']}"
        if not resize_to_original_dims:
            return masks
        return ops.resize_bilinear(masks.contiguous(), tuple(int(x) for x in meta_dict["orig_image_size"]))

    @torch.inference_mode()
    def decode(self, backbone_feats, high_res_feats, seg_tokens, out_size=None, reduce_queries=True):
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
        # ONE decoder call per prompt with the T frames as the batch (the reference loops over frames and batches the
        # prompts, sam2.py:103-114: T launches chains of ~60 small kernels each; here the image-side GEMMs see T*4096 rows)
        out = []
        for n_i in range(n):
            masks, _, _, _ = m.sam_mask_decoder(
                image_embeddings=feats, image_pe=image_pe, sparse_prompt_embeddings=sparse[n_i:n_i + 1].expand(T, -1, -1).contiguous(),
                dense_prompt_embeddings=dense[:1].expand(T, -1, -1, -1), multimask_output=False, repeat_image=False,
                high_res_features=[high_res_feats[0], high_res_feats[1]])             # [T,1,h,w]
            if out_size is not None:
                masks = ops.resize_bilinear(masks, out_size)
            out.append(masks.transpose(0, 1))                                          # [1,T,h,w]
        masks = torch.cat(out, 0)                                                  # [M*Q,T,h,w]
        if not reduce_queries:
            return masks
        masks = masks.reshape(-1, self.n_seg_queries, *masks.shape[1:])
        return masks.max(1).values                                                 # max over the Q queries (sam2.py:127-128)
