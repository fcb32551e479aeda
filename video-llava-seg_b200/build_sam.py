"""Model factory for the hot-path modules with the hyper-parameters every sam2.1_hiera_{t,s,b+,l}.yaml
shares (sam2/configs/sam2.1/*.yaml:26-116; the four files differ only in the image-encoder trunk/neck,
which is outside this package).  Hydra is not needed: the configuration is spelled out here."""
import torch

from .modeling.memory_attention import MemoryAttention, MemoryAttentionLayer
from .modeling.memory_encoder import CXBlock, Fuser, MaskDownSampler, MemoryEncoder
from .modeling.position_encoding import PositionEmbeddingSine
from .modeling.sam.mask_decoder import MaskDecoder
from .modeling.sam.prompt_encoder import PromptEncoder
from .modeling.sam.transformer import RoPEAttention, TwoWayTransformer


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
