"""Tracking core with the method surface of sam2/modeling/sam2_base.py (SAM2Base): forward_image,
_prepare_backbone_features, _forward_sam_heads, _prepare_memory_conditioned_features,
_encode_new_memory, track_step, _use_multimask.  Host orchestration stays Python; every arithmetic
step on the propagation path is a call into libvls_b200 (memory attention, mask decoder, SAM-heads
glue, fused memory encoder).  Differences from the reference that a caller can observe:
  * the [B,M,1024,1024] f32 `high_res_multimasks` are only materialised on request
    (`need_high_res=True`); the memory encoder consumes the low-res logits directly;
  * failures raise -- nothing is silently skipped.
"""
import ctypes
import math

import torch
from torch import nn

from .. import _pack, ops
from .._lib import check, lib, ptr, stream
from ._base import ctypes_ref, require_cuda
from .sam.mask_decoder import MaskDecoder
from .sam.prompt_encoder import PromptEncoder
from .sam.transformer import TwoWayTransformer
from .sam2_utils import MLP, get_1d_sine_pe, select_closest_cond_frames

NO_OBJ_SCORE = -1024.0  # sam2_base.py:19


class SAM2Base(nn.Module):
    def __init__(self, image_encoder, memory_attention, memory_encoder, num_maskmem=7, image_size=512,
                 backbone_stride=16, sigmoid_scale_for_mem_enc=1.0, sigmoid_bias_for_mem_enc=0.0,
                 binarize_mask_from_pts_for_mem_enc=False, use_mask_input_as_output_without_sam=False,
                 max_cond_frames_in_attn=-1, directly_add_no_mem_embed=False, use_high_res_features_in_sam=False,
                 multimask_output_in_sam=False, multimask_min_pt_num=1, multimask_max_pt_num=1,
                 multimask_output_for_tracking=False, use_multimask_token_for_obj_ptr=False,
                 iou_prediction_use_sigmoid=False, memory_temporal_stride_for_eval=1,
                 non_overlap_masks_for_mem_enc=False, use_obj_ptrs_in_encoder=False, max_obj_ptrs_in_encoder=16,
                 add_tpos_enc_to_obj_ptrs=True, proj_tpos_enc_in_obj_ptrs=False, use_signed_tpos_enc_to_obj_ptrs=False,
                 only_obj_ptrs_in_the_past_for_eval=False, pred_obj_scores=False, pred_obj_scores_mlp=False,
                 fixed_no_obj_ptr=False, soft_no_obj_ptr=False, use_mlp_for_obj_ptr_proj=False,
                 no_obj_embed_spatial=False, sam_mask_decoder_extra_args=None, compile_image_encoder=False):
        super().__init__()
        unsupported = []
        if not (use_high_res_features_in_sam and use_obj_ptrs_in_encoder and add_tpos_enc_to_obj_ptrs
                and proj_tpos_enc_in_obj_ptrs and pred_obj_scores and pred_obj_scores_mlp and fixed_no_obj_ptr
                and use_mlp_for_obj_ptr_proj and directly_add_no_mem_embed):
            unsupported.append("the SAM 2.1 head configuration (high-res features, object pointers with projected "
                               "temporal encoding, object-score MLP, fixed no-object pointer, no-mem embedding)")
        if soft_no_obj_ptr:
            unsupported.append("soft_no_obj_ptr")
        if compile_image_encoder:
            unsupported.append("compile_image_encoder (no tracing compilers on this path)")
        if unsupported:
            raise NotImplementedError("libvls_b200 implements " + "; ".join(unsupported))
        self.image_encoder = image_encoder
        self.use_high_res_features_in_sam = use_high_res_features_in_sam
        self.num_feature_levels = 3
        self.use_obj_ptrs_in_encoder = use_obj_ptrs_in_encoder
        self.max_obj_ptrs_in_encoder = max_obj_ptrs_in_encoder
        self.mask_downsample = nn.Conv2d(1, 1, kernel_size=4, stride=4)
        self.add_tpos_enc_to_obj_ptrs = add_tpos_enc_to_obj_ptrs
        self.proj_tpos_enc_in_obj_ptrs = proj_tpos_enc_in_obj_ptrs
        self.use_signed_tpos_enc_to_obj_ptrs = use_signed_tpos_enc_to_obj_ptrs
        self.only_obj_ptrs_in_the_past_for_eval = only_obj_ptrs_in_the_past_for_eval
        self.memory_attention = memory_attention
        self.hidden_dim = image_encoder.neck.d_model if hasattr(image_encoder, "neck") else 256
        self.memory_encoder = memory_encoder
        self.mem_dim = self.memory_encoder.out_proj.weight.shape[0]
        self.num_maskmem = num_maskmem
        self.maskmem_tpos_enc = nn.Parameter(torch.zeros(num_maskmem, 1, 1, self.mem_dim))
        self.no_mem_embed = nn.Parameter(torch.zeros(1, 1, self.hidden_dim))
        self.no_mem_pos_enc = nn.Parameter(torch.zeros(1, 1, self.hidden_dim))
        for p in (self.maskmem_tpos_enc, self.no_mem_embed, self.no_mem_pos_enc):
            nn.init.trunc_normal_(p, std=0.02)
        self.directly_add_no_mem_embed = directly_add_no_mem_embed
        self.sigmoid_scale_for_mem_enc = sigmoid_scale_for_mem_enc
        self.sigmoid_bias_for_mem_enc = sigmoid_bias_for_mem_enc
        self.binarize_mask_from_pts_for_mem_enc = binarize_mask_from_pts_for_mem_enc
        self.non_overlap_masks_for_mem_enc = non_overlap_masks_for_mem_enc
        self.memory_temporal_stride_for_eval = memory_temporal_stride_for_eval
        self.use_mask_input_as_output_without_sam = use_mask_input_as_output_without_sam
        self.multimask_output_in_sam = multimask_output_in_sam
        self.multimask_min_pt_num = multimask_min_pt_num
        self.multimask_max_pt_num = multimask_max_pt_num
        self.multimask_output_for_tracking = multimask_output_for_tracking
        self.use_multimask_token_for_obj_ptr = use_multimask_token_for_obj_ptr
        self.iou_prediction_use_sigmoid = iou_prediction_use_sigmoid
        self.image_size = image_size
        self.backbone_stride = backbone_stride
        self.sam_mask_decoder_extra_args = sam_mask_decoder_extra_args
        self.pred_obj_scores, self.pred_obj_scores_mlp = pred_obj_scores, pred_obj_scores_mlp
        self.fixed_no_obj_ptr, self.soft_no_obj_ptr = fixed_no_obj_ptr, soft_no_obj_ptr
        self.no_obj_ptr = nn.Parameter(torch.zeros(1, self.hidden_dim))
        nn.init.trunc_normal_(self.no_obj_ptr, std=0.02)
        self.use_mlp_for_obj_ptr_proj = use_mlp_for_obj_ptr_proj
        self.no_obj_embed_spatial = None
        if no_obj_embed_spatial:
            self.no_obj_embed_spatial = nn.Parameter(torch.zeros(1, self.mem_dim))
            nn.init.trunc_normal_(self.no_obj_embed_spatial, std=0.02)
        # SAM heads (sam2_base.py:207-255)
        self.sam_prompt_embed_dim = self.hidden_dim
        self.sam_image_embedding_size = image_size // backbone_stride
        s = self.sam_image_embedding_size
        self.sam_prompt_encoder = PromptEncoder(embed_dim=self.hidden_dim, image_embedding_size=(s, s),
                                                input_image_size=(image_size, image_size), mask_in_chans=16)
        self.sam_mask_decoder = MaskDecoder(
            num_multimask_outputs=3,
            transformer=TwoWayTransformer(depth=2, embedding_dim=self.hidden_dim, mlp_dim=2048, num_heads=8),
            transformer_dim=self.hidden_dim, iou_head_depth=3, iou_head_hidden_dim=256,
            use_high_res_features=use_high_res_features_in_sam, iou_prediction_use_sigmoid=iou_prediction_use_sigmoid,
            pred_obj_scores=pred_obj_scores, pred_obj_scores_mlp=pred_obj_scores_mlp,
            use_multimask_token_for_obj_ptr=use_multimask_token_for_obj_ptr, **(sam_mask_decoder_extra_args or {}))
        self.obj_ptr_proj = MLP(self.hidden_dim, self.hidden_dim, self.hidden_dim, 3)
        self.obj_ptr_tpos_proj = nn.Linear(self.hidden_dim, self.mem_dim)
        self.max_cond_frames_in_attn = max_cond_frames_in_attn
        self._consts = None
        # weight-derived constants (obj_ptr_proj, no_obj_ptr, maskmem_tpos_enc, ...) follow a checkpoint load
        self.register_load_state_dict_post_hook(lambda m, keys: setattr(m, "_consts", None))

    # ------------------------------------------------------------------ plumbing
    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, *args, **kwargs):
        raise NotImplementedError("Please use the corresponding methods in SAM2VideoPredictor for inference")

    def _apply(self, fn, *a, **kw):
        self._consts = None
        return super()._apply(fn, *a, **kw)

    def _constants(self):
        """Weight-derived constants of the tracking glue, built once per weight set on the device."""
        dev = self.device
        if self._consts is None or self._consts["device"] != dev:
            sd = {k: v for k, v in self.state_dict().items() if k.startswith(("obj_ptr_proj.", "no_obj_ptr"))}
            ptr_w, ptr_keep = _pack.pack_obj_ptr(sd, dev)
            s = self.sam_image_embedding_size
            pos = _pack.sine_pe_2d(self.mem_dim, s, s).to(dev)                  # maskmem_pos_enc, [64,s,s]
            pos_rows = pos.flatten(1).t().contiguous()                          # [s*s, 64]
            tpos = self.maskmem_tpos_enc.detach().float().reshape(self.num_maskmem, 1, self.mem_dim)
            self._consts = dict(
                device=dev, ptr_w=ptr_w, ptr_keep=ptr_keep, maskmem_pos=pos[None].contiguous(),
                # pos + tpos[num_maskmem - t_pos - 1] for t_pos = 0..num_maskmem-1 (sam2_base.py:581-583)
                mem_pos_rows=[(pos_rows + tpos[self.num_maskmem - t - 1]).contiguous() for t in range(self.num_maskmem)],
                tpos_w=self.obj_ptr_tpos_proj.weight.detach().to(dev, torch.bfloat16).contiguous(),
                tpos_b=self.obj_ptr_tpos_proj.bias.detach().to(dev, torch.float32).contiguous(),
                ptr_pos={}, no_mem_rows=self.no_mem_embed.detach().float().reshape(1, 1, -1), ws=None)
        return self._consts

    def _ptr_pos_rows(self, dist_key, num_frames):
        """Temporal encoding of pointers at the given signed frame distances -> [P*4, 64] f32
        (sam2_base.py:628-643).  Each distance's rows are a weight constant, computed once (sine PE + the
        obj_ptr_tpos_proj linear through vls_linear_f32) and cached; the nearest-15 block repeats every frame."""
        c = self._constants()
        t_diff_max = min(num_frames, self.max_obj_ptrs_in_encoder) - 1
        table = c["ptr_pos"].setdefault(t_diff_max, {})
        missing = [d for d in dist_key if d not in table]
        if missing:
            pos = torch.tensor(missing, device=c["device"], dtype=torch.float32) / t_diff_max
            rows = ops.linear_f32(get_1d_sine_pe(pos, dim=self.hidden_dim), c["tpos_w"], c["tpos_b"])   # [n, 64]
            rep = self.hidden_dim // self.mem_dim
            for i, d in enumerate(missing):
                table[d] = rows[i:i + 1].expand(rep, -1)
        tail_key = ("tail", t_diff_max, dist_key[1:])
        if tail_key not in c["ptr_pos"]:
            if len(c["ptr_pos"]) > 256:
                c["ptr_pos"] = {t_diff_max: table}
            c["ptr_pos"][tail_key] = torch.cat([table[d] for d in dist_key[1:]], 0) if len(dist_key) > 1 else None
        tail = c["ptr_pos"][tail_key]
        return table[dist_key[0]] if tail is None else torch.cat([table[dist_key[0]], tail], 0)

    # ------------------------------------------------------------------ image features (out of the hot path)
    def forward_image(self, img_batch):
        """Image encoder + conv_s0/conv_s1 (sam2_base.py:467-479). Runs in PyTorch: outside the hot path."""
        out = self.image_encoder(img_batch)
        if getattr(self.image_encoder, "applies_high_res_convs", False):   # GraphedImageEncoder: captured with the trunk
            return out
        dec = self.sam_mask_decoder
        out["backbone_fpn"][0] = nn.functional.conv2d(out["backbone_fpn"][0], dec.conv_s0.weight, dec.conv_s0.bias)
        out["backbone_fpn"][1] = nn.functional.conv2d(out["backbone_fpn"][1], dec.conv_s1.weight, dec.conv_s1.bias)
        return out

    def _prepare_backbone_features(self, backbone_out):
        """sam2_base.py:481-495: NxCxHxW -> HWxNxC for the last three levels."""
        backbone_out = backbone_out.copy()
        n = self.num_feature_levels
        maps, poss = backbone_out["backbone_fpn"][-n:], backbone_out["vision_pos_enc"][-n:]
        feat_sizes = [(x.shape[-2], x.shape[-1]) for x in poss]
        vision_feats = [x.flatten(2).permute(2, 0, 1) for x in maps]
        vision_pos = [x.flatten(2).permute(2, 0, 1) for x in poss]
        return backbone_out, vision_feats, vision_pos, feat_sizes

    # ------------------------------------------------------------------ SAM heads
    def _forward_sam_heads(self, backbone_features, point_inputs=None, mask_inputs=None, high_res_features=None,
                           multimask_output=False, need_high_res=True, defer_obj_ptr=False):
        """sam2_base.py:257-413. Returns the same 7-tuple; `high_res_*` entries are None unless need_high_res.
        defer_obj_ptr: the object-pointer MLP is enqueued on a forked stream and obj_ptr is only valid after
        `lib().vls_sam_heads_join(stream())` (the captured frame overlaps it with the memory encoder)."""
        B = backbone_features.size(0)
        dev = backbone_features.device
        require_cuda(backbone_features)
        if point_inputs is not None and "prompt_embedding" in point_inputs:
            assert point_inputs["prompt_embedding"].size(0) == B
        elif point_inputs is not None:
            coords, labels = point_inputs["point_coords"], point_inputs["point_labels"]
            assert coords.size(0) == B and labels.size(0) == B
        elif mask_inputs is not None:
            coords = torch.zeros(B, 1, 2, device=dev)
            labels = -torch.ones(B, 1, dtype=torch.int32, device=dev)
        if mask_inputs is not None:
            assert len(mask_inputs.shape) == 4 and mask_inputs.shape[:2] == (B, 1)
            if mask_inputs.shape[-2:] != self.sam_prompt_encoder.mask_input_size:
                mask_prompt = nn.functional.interpolate(mask_inputs.float(), size=self.sam_prompt_encoder.mask_input_size,
                                                        align_corners=False, mode="bilinear", antialias=True)
            else:
                mask_prompt = mask_inputs
        else:
            mask_prompt = None
        if point_inputs is not None and "prompt_embedding" in point_inputs:
            # [SEG]-token prompt: the projected LLM hidden state is the sparse prompt (llava sam2.py:75-88)
            pe = self.sam_prompt_encoder
            sparse = point_inputs["prompt_embedding"].to(dev).float()
            dense = pe.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(B, -1, *pe.image_embedding_size)
        elif point_inputs is None and mask_prompt is None:
            # no prompt on propagated frames: two `not_a_point` tokens + `no_mask` dense embedding are
            # weight constants (prompt_encoder.py:87-96,178-180) -- skip the per-frame prompt-encoder ops
            pe = self.sam_prompt_encoder
            cache = self._constants().setdefault("no_prompt_sparse", {})
            if B not in cache:       # materialised once per weight set: expand().contiguous() was a 2 us copy kernel per frame
                cache[B] = pe.not_a_point_embed.weight.detach().float().reshape(1, 1, -1).expand(B, 2, -1).contiguous()
            sparse = cache[B]
            dense = pe.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(B, -1, *pe.image_embedding_size)
        else:
            sparse, dense = self.sam_prompt_encoder(points=(coords, labels), boxes=None, masks=mask_prompt)
        masks4, iou4, tok4, obj_logits = self.sam_mask_decoder.predict_masks(
            backbone_features, self.sam_prompt_encoder.get_dense_pe(), sparse, dense, False, high_res_features)
        masks4, iou4, tok4, obj_logits = masks4.float(), iou4.float(), tok4.float(), obj_logits.float()
        c = self._constants()
        hw = masks4.shape[-2] * masks4.shape[-1]
        low = torch.empty((B, 1) + tuple(masks4.shape[-2:]), device=dev, dtype=torch.float32)
        obj_ptr = torch.empty((B, self.hidden_dim), device=dev, dtype=torch.float32)
        best = torch.empty((B,), device=dev, dtype=torch.int32)
        is_obj = torch.empty((B,), device=dev, dtype=torch.float32)
        occluded = torch.empty((B,), device=dev, dtype=torch.float32)
        ws = torch.empty((B * 256 * 3 * 4,), device=dev, dtype=torch.uint8)
        heads_post = lib().vls_sam_heads_post_deferred if defer_obj_ptr else lib().vls_sam_heads_post
        check(heads_post(ctypes_ref(c["ptr_w"]), ptr(masks4), ptr(iou4), ptr(tok4), ptr(obj_logits), B,
                         int(bool(multimask_output)), hw, ptr(low), ptr(obj_ptr), ptr(best), ptr(is_obj),
                         ptr(occluded), ptr(ws), ws.numel(), stream()), "vls_sam_heads_post")
        if multimask_output:
            low_multi, ious = masks4[:, 1:], iou4[:, 1:]
        else:
            low_multi, ious = masks4[:, 0:1], iou4[:, 0:1]
        high_multi = high = None
        if need_high_res:
            gated = torch.where((obj_logits > 0)[:, None, None], low_multi, NO_OBJ_SCORE)
            high_multi = ops.resize_bilinear(gated, (self.image_size, self.image_size))
            high = ops.resize_bilinear(low, (self.image_size, self.image_size))
            low_multi = gated
        self._last_gate = (obj_logits, occluded)
        return low_multi, high_multi, ious, low, high, obj_ptr, obj_logits

    def _use_mask_as_output(self, backbone_features, high_res_features, mask_inputs):
        """Mask prompt used verbatim as the output (sam2_base.py:415-465). Prompt handling, not on the
        propagation path: the anti-aliased down-sampling stays a PyTorch call."""
        out_scale, out_bias = 20.0, -10.0
        m = mask_inputs.float()
        high = m * out_scale + out_bias
        low = nn.functional.interpolate(high, size=(high.size(-2) // 4, high.size(-1) // 4), align_corners=False,
                                        mode="bilinear", antialias=True)
        ious = mask_inputs.new_ones(mask_inputs.size(0), 1).float()
        prompt = nn.functional.conv2d(m, self.mask_downsample.weight, self.mask_downsample.bias, stride=4)
        _, _, _, _, _, obj_ptr, _ = self._forward_sam_heads(backbone_features, mask_inputs=prompt,
                                                            high_res_features=high_res_features, need_high_res=False)
        is_obj = torch.any(mask_inputs.flatten(1).float() > 0.0, dim=1)[..., None].float()
        obj_logits = out_scale * is_obj + out_bias
        obj_ptr = is_obj * obj_ptr + (1 - is_obj) * self.no_obj_ptr
        return low, high, ious, low, high, obj_ptr, obj_logits

    # ------------------------------------------------------------------ memory conditioning
    def _gather_memory(self, frame_idx, output_dict, num_frames, track_in_reverse):
        """Which memories / pointers condition `frame_idx` (sam2_base.py:522-620), as index lists."""
        cond = output_dict["cond_frame_outputs"]
        assert len(cond) > 0
        selected, unselected = select_closest_cond_frames(frame_idx, cond, self.max_cond_frames_in_attn)
        mems = [(0, out) for out in selected.values()]
        stride = self.memory_temporal_stride_for_eval
        for t_pos in range(1, self.num_maskmem):
            t_rel = self.num_maskmem - t_pos
            if t_rel == 1:
                prev = frame_idx + t_rel if track_in_reverse else frame_idx - t_rel
            elif not track_in_reverse:
                prev = ((frame_idx - 2) // stride) * stride - (t_rel - 2) * stride
            else:
                prev = -(-(frame_idx + 2) // stride) * stride + (t_rel - 2) * stride
            out = output_dict["non_cond_frame_outputs"].get(prev, None)
            if out is None:
                out = unselected.get(prev, None)
            if out is not None:
                mems.append((t_pos, out))
        sign = -1 if track_in_reverse else 1
        max_ptrs = min(num_frames, self.max_obj_ptrs_in_encoder)
        if self.only_obj_ptrs_in_the_past_for_eval and not self.training:
            ptr_cond = {t: o for t, o in selected.items() if (t >= frame_idx if track_in_reverse else t <= frame_idx)}
        else:
            ptr_cond = selected
        ptrs = [((frame_idx - t) * sign if self.use_signed_tpos_enc_to_obj_ptrs else abs(frame_idx - t), o["obj_ptr"])
                for t, o in ptr_cond.items()]
        for t_diff in range(1, max_ptrs):
            t = frame_idx + t_diff if track_in_reverse else frame_idx - t_diff
            if t < 0 or (num_frames is not None and t >= num_frames):
                break
            out = output_dict["non_cond_frame_outputs"].get(t, unselected.get(t, None))
            if out is not None:
                ptrs.append((t_diff, out["obj_ptr"]))
        return mems, ptrs

    @staticmethod
    def _mem_rows(out):
        """Memory of one frame as token rows [B, HW, 64] bf16 (kept next to the NCHW copy by the predictor)."""
        rows = out.get("maskmem_rows", None)
        if rows is None:
            f = out["maskmem_features"]
            rows = f.flatten(2).transpose(1, 2).contiguous()
        return rows

    def _prepare_memory_conditioned_features(self, frame_idx, is_init_cond_frame, current_vision_feats,
                                             current_vision_pos_embeds, feat_sizes, output_dict, num_frames,
                                             track_in_reverse=False):
        """Fuse the current frame's features with the memory bank (sam2_base.py:497-674) -> [B,C,H,W]."""
        feats = current_vision_feats[-1]
        B, C = feats.size(1), self.hidden_dim
        H, W = feat_sizes[-1]
        dev = feats.device
        if self.num_maskmem == 0:
            return feats.permute(1, 2, 0).view(B, C, H, W)
        if is_init_cond_frame:
            # directly_add_no_mem_embed (sam2_base.py:651-655); a 1 MB broadcast add on a prompt frame only
            out = ops.add_rowvec(feats, self._constants()["no_mem_rows"].to(dev))
            return out.permute(1, 2, 0).reshape(B, C, H, W)
        mem_parts, pos_parts, n_ptr_tokens = self._gather_bank(frame_idx, output_dict, num_frames, track_in_reverse, B, dev)
        memory = torch.cat(mem_parts, dim=1)                                   # [B, Nk, 64]  (data movement only)
        memory_pos = torch.cat(pos_parts, dim=0)[None].expand(B, -1, -1)       # [B, Nk, 64], batch stride 0
        out = self.memory_attention(curr=current_vision_feats, curr_pos=current_vision_pos_embeds,
                                    memory=memory.transpose(0, 1), memory_pos=memory_pos.transpose(0, 1),
                                    num_obj_ptr_tokens=n_ptr_tokens)
        return out.permute(1, 2, 0).reshape(B, C, H, W)

    def _gather_bank(self, frame_idx, output_dict, num_frames, track_in_reverse, B, dev):
        """The memory bank of `frame_idx` as lists of [B, n_i, 64] memories and [n_i, 64] positional rows in the
        reference's key order (sam2_base.py:533-646), plus the number of pointer tokens at the end."""
        c = self._constants()
        mems, ptrs = self._gather_memory(frame_idx, output_dict, num_frames, track_in_reverse)
        mem_parts = [self._mem_rows(o).to(dev, non_blocking=True) for _, o in mems]
        pos_parts = [c["mem_pos_rows"][t_pos] for t_pos, _ in mems]
        n_ptr_tokens = 0
        if ptrs:
            k = self.hidden_dim // self.mem_dim
            dists = tuple(int(d) for d, _ in ptrs)
            ptr_stack = torch.stack([p for _, p in ptrs], dim=1)              # [B, P, 256]
            mem_parts.append(ptr_stack.reshape(B, len(ptrs) * k, self.mem_dim).to(mem_parts[0].dtype))
            pos_parts.append(self._ptr_pos_rows(dists, num_frames))
            n_ptr_tokens = len(ptrs) * k
        return mem_parts, pos_parts, n_ptr_tokens

    # ------------------------------------------------------------------ memory encoding
    def _encode_new_memory(self, current_vision_feats, feat_sizes, pred_masks_high_res, object_score_logits,
                           is_mask_from_pts):
        """Reference-signature variant taking the [B,1,1024,1024] mask (sam2_base.py:676-724)."""
        B, C = current_vision_feats[-1].size(1), self.hidden_dim
        H, W = feat_sizes[-1]
        pix = current_vision_feats[-1].permute(1, 2, 0).reshape(B, C, H, W)
        if self.non_overlap_masks_for_mem_enc and not self.training:
            pred_masks_high_res = self._apply_non_overlapping_constraints(pred_masks_high_res)
        binarize = self.binarize_mask_from_pts_for_mem_enc and is_mask_from_pts and not self.training
        enc = self.memory_encoder
        feats, _ = enc._encode(pix, None, pred_masks_high_res, 4 if binarize else 1, self.sigmoid_scale_for_mem_enc,
                               self.sigmoid_bias_for_mem_enc, self._occluded_gate(object_score_logits),
                               self.no_obj_embed_spatial, want_rows=False, out_dtype=torch.float32)
        pos = self._constants()["maskmem_pos"].expand(B, -1, -1, -1)
        return feats, [pos]

    def _occluded_gate(self, object_score_logits):
        """(1 - is_obj) per object for the occlusion embedding (sam2_base.py:716-722)."""
        if self.no_obj_embed_spatial is None:
            return None
        last = getattr(self, "_last_gate", None)
        if last is not None and last[0] is object_score_logits:
            return last[1]            # produced by the heads-post kernel of this very frame
        return (object_score_logits.reshape(-1) <= 0).float().contiguous()

    def _encode_new_memory_low_res(self, current_vision_feats, low_res_masks, object_score_logits, is_mask_from_pts):
        """Fast path: memory straight from the [B,1,4H,4W] logits -> (NCHW bf16, rows bf16, [pos])."""
        if self.non_overlap_masks_for_mem_enc and not self.training and low_res_masks.size(0) > 1:
            raise NotImplementedError("non_overlap_masks_for_mem_enc needs the high-res path (_encode_new_memory)")
        binarize = self.binarize_mask_from_pts_for_mem_enc and is_mask_from_pts and not self.training
        nchw, rows = self.memory_encoder.encode_from_low_res(
            current_vision_feats[-1], low_res_masks, binarize, self.sigmoid_scale_for_mem_enc,
            self.sigmoid_bias_for_mem_enc, self._occluded_gate(object_score_logits), self.no_obj_embed_spatial)
        pos = self._constants()["maskmem_pos"].expand(low_res_masks.size(0), -1, -1, -1)
        return nchw, rows, [pos]

    # ------------------------------------------------------------------ one tracking step
    def track_step(self, frame_idx, is_init_cond_frame, current_vision_feats, current_vision_pos_embeds, feat_sizes,
                   point_inputs, mask_inputs, output_dict, num_frames, track_in_reverse=False, run_mem_encoder=True,
                   prev_sam_mask_logits=None, need_high_res=False):
        """sam2_base.py:812-877. `pred_masks_high_res` is None unless need_high_res (the predictor drops it)."""
        current_out = {"point_inputs": point_inputs, "mask_inputs": mask_inputs}
        high_res_features = [x.permute(1, 2, 0).reshape(x.size(1), x.size(2), *s)
                             for x, s in zip(current_vision_feats[:-1], feat_sizes[:-1])]
        if mask_inputs is not None and self.use_mask_input_as_output_without_sam:
            pix = current_vision_feats[-1].permute(1, 2, 0).reshape(-1, self.hidden_dim, *feat_sizes[-1])
            sam_outputs = self._use_mask_as_output(pix, high_res_features, mask_inputs)
            low_for_mem = None
        else:
            pix = self._prepare_memory_conditioned_features(
                frame_idx=frame_idx, is_init_cond_frame=is_init_cond_frame,
                current_vision_feats=current_vision_feats[-1:], current_vision_pos_embeds=current_vision_pos_embeds[-1:],
                feat_sizes=feat_sizes[-1:], output_dict=output_dict, num_frames=num_frames,
                track_in_reverse=track_in_reverse)
            if prev_sam_mask_logits is not None:
                assert point_inputs is not None and mask_inputs is None
                mask_inputs = prev_sam_mask_logits
            multimask = self._use_multimask(is_init_cond_frame, point_inputs)
            sam_outputs = self._forward_sam_heads(pix, point_inputs=point_inputs, mask_inputs=mask_inputs,
                                                  high_res_features=high_res_features, multimask_output=multimask,
                                                  need_high_res=need_high_res)
            low_for_mem = sam_outputs[3]
        _, _, _, low_res_masks, high_res_masks, obj_ptr, object_score_logits = sam_outputs
        current_out["pred_masks"] = low_res_masks
        current_out["pred_masks_high_res"] = high_res_masks
        current_out["obj_ptr"] = obj_ptr
        current_out["object_score_logits"] = object_score_logits
        current_out["maskmem_features"] = current_out["maskmem_pos_enc"] = current_out["maskmem_rows"] = None
        if run_mem_encoder and self.num_maskmem > 0:
            if low_for_mem is not None:
                nchw, rows, pos = self._encode_new_memory_low_res(current_vision_feats, low_for_mem, object_score_logits,
                                                                  point_inputs is not None)
                current_out["maskmem_rows"] = rows
            else:
                nchw, pos = self._encode_new_memory(current_vision_feats, feat_sizes, high_res_masks,
                                                    object_score_logits, point_inputs is not None)
            current_out["maskmem_features"], current_out["maskmem_pos_enc"] = nchw, pos
        return current_out

    def _use_multimask(self, is_init_cond_frame, point_inputs):
        """sam2_base.py:879-887."""
        if point_inputs is not None and "prompt_embedding" in point_inputs:
            return False  # embedding prompts decode a single mask, like the LLaVA seg head (llava sam2.py:111)
        num_pts = 0 if point_inputs is None else point_inputs["point_labels"].size(1)
        return (self.multimask_output_in_sam and (is_init_cond_frame or self.multimask_output_for_tracking)
                and (self.multimask_min_pt_num <= num_pts <= self.multimask_max_pt_num))

    def _apply_non_overlapping_constraints(self, pred_masks):
        """Keep only the top-scoring object per pixel (sam2_base.py:889-907); off in every shipped config.
        Cross-object, per-video post-processing outside the per-object hot path: PyTorch ops."""
        if pred_masks.size(0) == 1:
            return pred_masks
        top = torch.argmax(pred_masks, dim=0, keepdim=True)
        keep = top == torch.arange(pred_masks.size(0), device=pred_masks.device)[:, None, None, None]
        return torch.where(keep, pred_masks, torch.clamp(pred_masks, max=-10.0))
