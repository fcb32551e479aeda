"""Drop-in `MaskDecoder` (sam2/modeling/sam/mask_decoder.py): same constructor keywords, state_dict
keys and forward/predict_masks signatures; one C call (vls_mask_decoder_forward) runs the two-way
transformer, the ConvTranspose up-scaling with high-res skips and the hyper-network mask product."""
import torch
from torch import nn

from ... import _pack
from ..._lib import LL4, check, lib, ptr, stream
from .._base import PackedModule, batch_shared, ctypes_ref, dtype_code, require_cuda
from ..sam2_utils import MLP, LayerNorm2d


class MaskDecoder(PackedModule):
    def __init__(self, *, transformer_dim, transformer, num_multimask_outputs=3, activation=nn.GELU, iou_head_depth=3,
                 iou_head_hidden_dim=256, use_high_res_features=False, iou_prediction_use_sigmoid=False,
                 dynamic_multimask_via_stability=False, dynamic_multimask_stability_delta=0.05,
                 dynamic_multimask_stability_thresh=0.98, pred_obj_scores=False, pred_obj_scores_mlp=False,
                 use_multimask_token_for_obj_ptr=False):
        super().__init__()
        if (transformer_dim, num_multimask_outputs, iou_head_depth, iou_head_hidden_dim) != (256, 3, 3, 256) or \
                activation is not nn.GELU or not (use_high_res_features and pred_obj_scores and pred_obj_scores_mlp):
            raise NotImplementedError("CUDA path implements the SAM 2.1 decoder: dim 256, 3+1 mask tokens, GELU "
                                      "up-scaling with high-res features, object-score MLP")
        self.transformer_dim, self.transformer = transformer_dim, transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.pred_obj_scores = pred_obj_scores
        self.obj_score_token = nn.Embedding(1, transformer_dim)
        self.use_multimask_token_for_obj_ptr = use_multimask_token_for_obj_ptr
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4), activation(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2), activation())
        self.use_high_res_features = use_high_res_features
        self.conv_s0 = nn.Conv2d(transformer_dim, transformer_dim // 8, kernel_size=1, stride=1)
        self.conv_s1 = nn.Conv2d(transformer_dim, transformer_dim // 4, kernel_size=1, stride=1)
        self.output_hypernetworks_mlps = nn.ModuleList(
            MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens))
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth,
                                       sigmoid_output=iou_prediction_use_sigmoid)
        self.pred_obj_score_head = MLP(transformer_dim, transformer_dim, 1, 3)
        # accepted for config compatibility; the stability fallback is commented out in this fork (:149-150)
        self.dynamic_multimask_via_stability = dynamic_multimask_via_stability
        self.dynamic_multimask_stability_delta = dynamic_multimask_stability_delta
        self.dynamic_multimask_stability_thresh = dynamic_multimask_stability_thresh
        self._pe_key = None

    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output,
                repeat_image, high_res_features=None):
        """-> (masks [B,3|1,4H,4W], iou_pred [B,3|1], sam_tokens_out [B,3|1,256], object_score_logits [B,1])."""
        masks, iou_pred, mask_tokens_out, obj = self.predict_masks(
            image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, repeat_image,
            high_res_features)
        if multimask_output:
            masks, iou_pred = masks[:, 1:, :, :], iou_pred[:, 1:]
        else:
            masks, iou_pred = masks[:, 0:1, :, :], iou_pred[:, 0:1]
        if multimask_output and self.use_multimask_token_for_obj_ptr:
            sam_tokens_out = mask_tokens_out[:, 1:]
        else:
            sam_tokens_out = mask_tokens_out[:, 0:1]
        return masks, iou_pred, sam_tokens_out, obj

    def predict_masks(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                      repeat_image, high_res_features=None):
        require_cuda(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings)
        if high_res_features is None:
            raise RuntimeError("use_high_res_features=True: high_res_features=[feat_s0, feat_s1] is required")
        assert image_pe.size(0) == 1, "image_pe should have size 1 in batch dim (from `get_dense_pe()`)"
        dev = image_embeddings.device
        B, Ns = sparse_prompt_embeddings.shape[0], sparse_prompt_embeddings.shape[1]
        _, C, H, W = image_embeddings.shape
        if not repeat_image:
            assert image_embeddings.shape[0] == B
        pe_key = (image_pe.data_ptr(), tuple(image_pe.shape), image_pe.dtype, getattr(image_pe, "_vls_version", 0))
        if self._packed is None or self._pe_key != pe_key:
            self._packed = _pack.pack_mask_decoder(self._flat_sd(), "", dev, image_pe,
                                                   self.iou_prediction_head.sigmoid_output)
            self._pe_key = pe_key
        w = self._packed[0]
        (feat_s0, s0_bs), (feat_s1, s1_bs) = batch_shared(high_res_features[0]), batch_shared(high_res_features[1])
        emb = image_embeddings
        es = LL4(0 if (repeat_image or emb.shape[0] == 1) else emb.stride(0), emb.stride(1), emb.stride(2), emb.stride(3))
        dn = dense_prompt_embeddings
        ds = LL4(dn.stride(0) if dn.shape[0] > 1 else 0, dn.stride(1), dn.stride(2), dn.stride(3))
        sparse = sparse_prompt_embeddings.float().contiguous()
        masks = torch.empty((B, 4, 4 * H, 4 * W), device=dev, dtype=torch.float32)
        iou = torch.empty((B, 4), device=dev, dtype=torch.float32)
        tok = torch.empty((B, 4, 256), device=dev, dtype=torch.float32)
        obj = torch.empty((B, 1), device=dev, dtype=torch.float32)
        nbytes = lib().vls_mask_decoder_workspace_bytes(B, Ns, H, W)
        ws = self._workspace(nbytes, dev)
        check(lib().vls_mask_decoder_forward(
            ctypes_ref(w), ptr(emb), dtype_code(emb), es, ptr(dn), dtype_code(dn), ds, ptr(sparse), ptr(feat_s0),
            dtype_code(feat_s0), s0_bs, ptr(feat_s1), dtype_code(feat_s1), s1_bs, B, Ns, H, W, ptr(masks), ptr(iou), ptr(tok), ptr(obj),
            ptr(ws), ws.numel(), stream()), "vls_mask_decoder_forward")
        dt = image_embeddings.dtype
        return masks.to(dt), iou.to(dt), tok.to(dt), obj.to(dt)
