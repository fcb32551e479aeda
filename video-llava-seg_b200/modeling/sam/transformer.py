"""Parameter containers for the attention blocks of sam2/modeling/sam/transformer.py (Attention :215-253,
RoPEAttention :289-309, TwoWayAttentionBlock :137-179, TwoWayTransformer :44-88).  They hold the
q/k/v/out projections under the reference's names; the math runs in libvls_b200 (see MaskDecoder and
MemoryAttention, which own the forward passes)."""
from torch import nn

from ..sam2_utils import MLP


class Attention(nn.Module):
    def __init__(self, embedding_dim, num_heads, downsample_rate=1, dropout=0.0, kv_in_dim=None):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.kv_in_dim = kv_in_dim if kv_in_dim is not None else embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(self.kv_in_dim, self.internal_dim)
        self.v_proj = nn.Linear(self.kv_in_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)
        self.dropout_p = dropout


class RoPEAttention(Attention):
    """Attention whose q/k are rotated by the axial 2-D RoPE.  `feat_sizes` is accepted for config
    compatibility; like the reference (:325-328) the table follows the actual sqrt(Nq) grid."""

    def __init__(self, *args, rope_theta=10000.0, rope_k_repeat=False, feat_sizes=(32, 32), **kwargs):
        super().__init__(*args, **kwargs)
        if rope_theta != 10000.0 or self.num_heads != 1 or self.internal_dim != 256:
            raise NotImplementedError("CUDA path implements the SAM 2.1 RoPE attention: theta 1e4, one head of 256")
        self.rope_theta, self.rope_k_repeat, self.feat_sizes = rope_theta, rope_k_repeat, tuple(feat_sizes)


class TwoWayAttentionBlock(nn.Module):
    def __init__(self, embedding_dim, num_heads, mlp_dim=2048, activation=nn.ReLU, attention_downsample_rate=2,
                 skip_first_layer_pe=False):
        super().__init__()
        self.self_attn = Attention(embedding_dim, num_heads)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.cross_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.mlp = MLP(embedding_dim, mlp_dim, embedding_dim, num_layers=2, activation=activation)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe


class TwoWayTransformer(nn.Module):
    def __init__(self, depth, embedding_dim, num_heads, mlp_dim, activation=nn.ReLU, attention_downsample_rate=2):
        super().__init__()
        if (depth, embedding_dim, num_heads, mlp_dim, attention_downsample_rate) != (2, 256, 8, 2048, 2):
            raise NotImplementedError("CUDA path implements the SAM two-way transformer: depth 2, dim 256, 8 heads, "
                                      "MLP 2048, down-sample rate 2")
        self.depth, self.embedding_dim, self.num_heads, self.mlp_dim = depth, embedding_dim, num_heads, mlp_dim
        self.layers = nn.ModuleList(
            TwoWayAttentionBlock(embedding_dim, num_heads, mlp_dim, activation, attention_downsample_rate, i == 0)
            for i in range(depth))
        self.final_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)
