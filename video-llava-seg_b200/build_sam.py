"""Model factory for the hot-path modules with the hyper-parameters every sam2.1_hiera_{t,s,b+,l}.yaml
shares (sam2/configs/sam2.1/*.yaml:26-116; the four files differ only in the image-encoder trunk/neck:
IMAGE_ENCODER_VARIANTS below).  Hydra is not needed: the configuration is spelled out here."""
import torch

from .modeling.memory_attention import MemoryAttention, MemoryAttentionLayer
from .modeling.memory_encoder import CXBlock, Fuser, MaskDownSampler, MemoryEncoder
from .modeling.position_encoding import PositionEmbeddingSine
from .modeling.sam.mask_decoder import MaskDecoder
from .modeling.sam.prompt_encoder import PromptEncoder
from .modeling.sam.transformer import RoPEAttention, TwoWayTransformer


# sam2/configs/sam2.1/sam2.1_hiera_{t,s,b+,l}.yaml:6-24 (trunk keywords; the neck's channel list follows from them)
IMAGE_ENCODER_VARIANTS = {
    "t": dict(embed_dim=96, num_heads=1, stages=(1, 2, 7, 2), global_att_blocks=(5, 7, 9),
              window_pos_embed_bkg_spatial_size=(7, 7)),
    "s": dict(embed_dim=96, num_heads=1, stages=(1, 2, 11, 2), global_att_blocks=(7, 10, 13),
              window_pos_embed_bkg_spatial_size=(7, 7)),
    "b+": dict(embed_dim=112, num_heads=2),
    "l": dict(embed_dim=144, num_heads=2, stages=(2, 6, 36, 4), global_att_blocks=(23, 33, 43),
              window_pos_embed_bkg_spatial_size=(7, 7), window_spec=(8, 4, 16, 8)),
}


def build_image_encoder(variant="b+"):
    """Hiera trunk + FPN neck of sam2.1_hiera_<variant>.yaml; PyTorch modules."""
    from .modeling.backbones.hieradet import Hiera
    from .modeling.backbones.image_encoder import FpnNeck, ImageEncoder

    trunk = Hiera(**IMAGE_ENCODER_VARIANTS[variant])
    neck = FpnNeck(position_encoding=PositionEmbeddingSine(num_pos_feats=256, normalize=True, scale=None, temperature=10000),
                   d_model=256, backbone_channel_list=list(trunk.channel_list), fpn_top_down_levels=[2, 3],
                   fpn_interp_model="nearest")
    return ImageEncoder(trunk=trunk, neck=neck, scalp=1)


def build_memory_attention():
    layer = MemoryAttentionLayer(
        activation="relu", dim_feedforward=2048, dropout=0.1, pos_enc_at_attn=False, d_model=256,
        pos_enc_at_cross_attn_keys=True, pos_enc_at_cross_attn_queries=False,
        self_attention=RoPEAttention(rope_theta=10000.0, feat_sizes=[32, 32], embedding_dim=256, num_heads=1,
                                     downsample_rate=1, dropout=0.1),
        cross_attention=RoPEAttention(rope_theta=10000.0, feat_sizes=[32, 32], rope_k_repeat=True, embedding_dim=256,
                                      num_heads=1, downsample_rate=1, dropout=0.1, kv_in_dim=64))
    return MemoryAttention(d_model=256, pos_enc_at_input=True, layer=layer, num_layers=4)


def build_memory_encoder():
    return MemoryEncoder(
        out_dim=64, position_encoding=PositionEmbeddingSine(num_pos_feats=64, normalize=True, scale=None, temperature=10000),
        mask_downsampler=MaskDownSampler(kernel_size=3, stride=2, padding=1),
        fuser=Fuser(layer=CXBlock(dim=256, kernel_size=7, padding=3, layer_scale_init_value=1e-6, use_dwconv=True),
                    num_layers=2))


def build_mask_decoder(**extra):
    return MaskDecoder(num_multimask_outputs=3,
                       transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                       transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256, use_high_res_features=True,
                       iou_prediction_use_sigmoid=True, pred_obj_scores=True, pred_obj_scores_mlp=True,
                       use_multimask_token_for_obj_ptr=True, **extra)


def build_prompt_encoder(image_size=1024, backbone_stride=16):
    s = image_size // backbone_stride
    return PromptEncoder(embed_dim=256, image_embedding_size=(s, s), input_image_size=(image_size, image_size),
                         mask_in_chans=16)


def load_prefixed(module, sd, prefix):
    """Strictly load the `prefix.*` slice of a flat reference-style state_dict into `module`."""
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    module.load_state_dict(sub, strict=True)
    return module


SAM21_MODEL_KWARGS = dict(  # sam2/configs/sam2.1/sam2.1_hiera_*.yaml:84-116 (identical for t/s/b+/l)
    num_maskmem=7, image_size=1024, sigmoid_scale_for_mem_enc=20.0, sigmoid_bias_for_mem_enc=-10.0,
    use_mask_input_as_output_without_sam=True, directly_add_no_mem_embed=True, no_obj_embed_spatial=True,
    use_high_res_features_in_sam=True, multimask_output_in_sam=True, iou_prediction_use_sigmoid=True,
    use_obj_ptrs_in_encoder=True, add_tpos_enc_to_obj_ptrs=True, proj_tpos_enc_in_obj_ptrs=True,
    use_signed_tpos_enc_to_obj_ptrs=True, only_obj_ptrs_in_the_past_for_eval=True, pred_obj_scores=True,
    pred_obj_scores_mlp=True, fixed_no_obj_ptr=True, multimask_output_for_tracking=True,
    use_multimask_token_for_obj_ptr=True, multimask_min_pt_num=0, multimask_max_pt_num=1,
    use_mlp_for_obj_ptr_proj=True, compile_image_encoder=False)


def build_sam2_video_predictor(image_encoder=None, state_dict=None, device="cuda", apply_postprocessing=True,
                               image_encoder_dtype=None, **overrides):
    """Counterpart of sam2/build_sam.py:79-118.  `image_encoder`: a variant name ("t", "s", "b+", "l": the Hiera + FPN
    encoder of that configuration is built), any module returning the reference's {"backbone_fpn", "vision_pos_enc"}
    dict, or None when frames come with precomputed features.  `state_dict`: reference checkpoint["model"]
    (image_encoder.* keys are loaded into the encoder if there is one, otherwise ignored).
    `image_encoder_dtype` (e.g. torch.bfloat16): after loading, run forward_image in that precision under a CUDA graph
    (GraphedImageEncoder); None keeps the f32 eager module, as the reference."""
    from .sam2_video_predictor import SAM2VideoPredictor

    kw = dict(SAM21_MODEL_KWARGS)
    if apply_postprocessing:  # build_sam.py:93-102
        kw.update(binarize_mask_from_pts_for_mem_enc=True, fill_hole_area=8,
                  sam_mask_decoder_extra_args=dict(dynamic_multimask_via_stability=True,
                                                   dynamic_multimask_stability_delta=0.05,
                                                   dynamic_multimask_stability_thresh=0.98))
    kw.update(overrides)
    if isinstance(image_encoder, str):
        image_encoder = build_image_encoder(image_encoder)
    model = SAM2VideoPredictor(image_encoder=image_encoder, memory_attention=build_memory_attention(),
                               memory_encoder=build_memory_encoder(), **kw)
    if state_dict is not None:
        own = model.state_dict()
        sd = {k: v for k, v in state_dict.items() if k in own}
        missing = [k for k in own if k not in sd]
        if missing:
            raise RuntimeError(f"checkpoint is missing hot-path keys, e.g. {missing[:5]}")
        model.load_state_dict(sd, strict=True)
    model = model.to(device).eval()
    if image_encoder_dtype is not None and model.image_encoder is not None:
        from .modeling.backbones.image_encoder import GraphedImageEncoder

        model.image_encoder = GraphedImageEncoder(model.image_encoder, model.sam_mask_decoder.conv_s0,
                                                  model.sam_mask_decoder.conv_s1, dtype=image_encoder_dtype)
    return model
