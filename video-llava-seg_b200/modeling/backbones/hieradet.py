"""Hiera trunk (sam2/modeling/backbones/hieradet.py:161-317, utils.py:15-95) hosted in
PyTorch next to the hot path: same constructor keywords, `state_dict` keys and outputs as the reference, so a
reference checkpoint's `image_encoder.trunk.*` slice loads strictly.  No new kernels here (first step of f-4): the
module is meant to run in bf16 under a CUDA graph (image_encoder.GraphedImageEncoder), every block being
    x -> LN -> [pad to windows] -> fused qkv -> (max-pooled q at a stage change) -> SDPA -> proj -> + shortcut -> LN -> MLP(GELU)
with tokens kept as [B, H, W, C] (channels last) throughout.

Behaviour pinned to the reference by tests/golden/image_encoder.npz (make_golden.py::image_encoder_case), including the
details that are easy to lose: windows are zero-padded AFTER norm1 and the padded tokens take part in attention as
keys; q is pooled inside each window; the first block of a stage still uses the previous stage's window size; the
background positional embedding is bicubically resized and a tiled window embedding is added."""
from functools import partial

import torch
import torch.nn.functional as F
from torch import nn


class _MLP(nn.Module):
    """`layers.{i}` container + forward (the trunk's two-layer GELU MLP; sam2_utils.py:112-140)."""

    def __init__(self, dims, act=nn.GELU):
        super().__init__()
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.act = act()

    def forward(self, x):
        for i, lin in enumerate(self.layers):
            x = lin(x)
            if i + 1 < len(self.layers):
                x = self.act(x)
        return x


class PatchEmbed(nn.Module):
    """7x7 / stride-4 patchify convolution, output channels last (utils.py:62-95)."""

    def __init__(self, kernel_size=(7, 7), stride=(4, 4), padding=(3, 3), in_chans=3, embed_dim=768):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)

    def forward(self, x):
        return self.proj(x).permute(0, 2, 3, 1)


def _to_windows(x, ws):
    """[B,H,W,C] -> ([B*nh*nw, ws, ws, C], (nh, nw)); zero padding at the bottom / right (utils.py:15-38)."""
    B, H, W, C = x.shape
    ph, pw = (-H) % ws, (-W) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    nh, nw = (H + ph) // ws, (W + pw) // ws
    x = x.reshape(B, nh, ws, nw, ws, C).transpose(2, 3)
    return x.reshape(B * nh * nw, ws, ws, C), (nh, nw)


def _from_windows(xw, grid, hw):
    """Inverse of _to_windows for windows of any (already pooled) size; crops the padding (utils.py:41-59)."""
    nh, nw = grid
    ws, C = xw.shape[1], xw.shape[-1]
    B = xw.shape[0] // (nh * nw)
    x = xw.reshape(B, nh, nw, ws, ws, C).transpose(2, 3).reshape(B, nh * ws, nw * ws, C)
    return x[:, :hw[0], :hw[1]]


class MultiScaleAttention(nn.Module):
    """Fused-qkv attention over [B', H, W, C] token grids; q optionally max-pooled (hieradet.py:37-82)."""

    def __init__(self, dim, dim_out, num_heads, q_pool=None):
        super().__init__()
        self.dim, self.dim_out, self.num_heads, self.q_pool = dim, dim_out, num_heads, q_pool
        self.qkv = nn.Linear(dim, dim_out * 3)
        self.proj = nn.Linear(dim_out, dim_out)

    def forward(self, x):
        B, H, W, _ = x.shape
        nh, hd = self.num_heads, self.dim_out // self.num_heads
        q, k, v = self.qkv(x).reshape(B, H * W, 3, nh, hd).unbind(2)
        if self.q_pool is not None:
            q = self.q_pool(q.reshape(B, H, W, nh * hd).permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
            H, W = q.shape[1], q.shape[2]
            q = q.reshape(B, H * W, nh, hd)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return self.proj(o.transpose(1, 2).reshape(B, H, W, nh * hd))


class MultiScaleBlock(nn.Module):
    """hieradet.py:85-158."""

    def __init__(self, dim, dim_out, num_heads, mlp_ratio=4.0, drop_path=0.0, norm_layer="LayerNorm", q_stride=None,
                 act_layer=nn.GELU, window_size=0):
        super().__init__()
        if drop_path:
            raise NotImplementedError("inference-only trunk: stochastic depth is not implemented")
        if isinstance(norm_layer, str):
            norm_layer = partial(getattr(nn, norm_layer), eps=1e-6)
        self.dim, self.dim_out, self.window_size, self.q_stride = dim, dim_out, window_size, q_stride
        self.norm1 = norm_layer(dim)
        self.pool = nn.MaxPool2d(kernel_size=q_stride, stride=q_stride, ceil_mode=False) if q_stride else None
        self.attn = MultiScaleAttention(dim, dim_out, num_heads=num_heads, q_pool=self.pool)
        self.norm2 = norm_layer(dim_out)
        self.mlp = _MLP([dim_out, int(dim_out * mlp_ratio), dim_out], act_layer)
        if dim != dim_out:
            self.proj = nn.Linear(dim, dim_out)

    def forward(self, x):
        t = self.norm1(x)
        if self.dim != self.dim_out:      # stage change: the shortcut is projected (and pooled like q)
            x = self.proj(t)
            if self.pool is not None:
                x = self.pool(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        if self.window_size > 0:
            tw, grid = _to_windows(t, self.window_size)
            a = _from_windows(self.attn(tw), grid, x.shape[1:3])   # pooled windows are window_size / stride wide
        else:
            a = self.attn(t)
        x = x + a
        return x + self.mlp(self.norm2(x))


class Hiera(nn.Module):
    """Returns the per-stage feature maps [B, C_s, H_s, W_s], highest resolution first (hieradet.py:161-317)."""

    def __init__(self, embed_dim=96, num_heads=1, drop_path_rate=0.0, q_pool=3, q_stride=(2, 2), stages=(2, 3, 16, 3),
                 dim_mul=2.0, head_mul=2.0, window_pos_embed_bkg_spatial_size=(14, 14), window_spec=(8, 4, 14, 7),
                 global_att_blocks=(12, 16, 20), weights_path=None, return_interm_layers=True):
        super().__init__()
        if weights_path is not None:
            raise NotImplementedError("load weights with load_state_dict (no path manager in this package)")
        assert len(stages) == len(window_spec)
        self.window_spec, self.q_stride = tuple(window_spec), tuple(q_stride)
        self.stage_ends = [sum(stages[:i]) - 1 for i in range(1, len(stages) + 1)]
        assert 0 <= q_pool <= len(self.stage_ends) - 1
        self.q_pool_blocks = [e + 1 for e in self.stage_ends[:-1]][:q_pool]
        self.return_interm_layers = return_interm_layers
        self.global_att_blocks = tuple(global_att_blocks) if global_att_blocks is not None else ()
        self.patch_embed = PatchEmbed(embed_dim=embed_dim)
        self.window_pos_embed_bkg_spatial_size = tuple(window_pos_embed_bkg_spatial_size)
        self.pos_embed = nn.Parameter(torch.zeros(1, embed_dim, *self.window_pos_embed_bkg_spatial_size))
        self.pos_embed_window = nn.Parameter(torch.zeros(1, embed_dim, self.window_spec[0], self.window_spec[0]))
        self.blocks = nn.ModuleList()
        dim, heads, stage = embed_dim, num_heads, 0
        for i in range(sum(stages)):
            # the window size lags one block behind the stage change (the first block of a stage attends in the previous
            # stage's windows and pools q down to the new resolution)
            ws = 0 if i in self.global_att_blocks else self.window_spec[stage]
            dim_out = dim
            if i - 1 in self.stage_ends:
                dim_out, heads, stage = int(dim * dim_mul), int(heads * head_mul), stage + 1
            self.blocks.append(MultiScaleBlock(dim=dim, dim_out=dim_out, num_heads=heads, drop_path=0.0,
                                               q_stride=self.q_stride if i in self.q_pool_blocks else None, window_size=ws))
            dim = dim_out
        self.channel_list = ([self.blocks[e].dim_out for e in self.stage_ends[::-1]] if return_interm_layers
                             else [self.blocks[-1].dim_out])
        self._pos_cache = {}

    def _get_pos_embed(self, hw):
        """bicubic-resized background embedding + tiled window embedding, [1,H,W,C] (hieradet.py:265-273)."""
        key = (tuple(hw), self.pos_embed.dtype, self.pos_embed.device, self.pos_embed._version, self.pos_embed_window._version)
        if self._pos_cache.get("key") != key:
            bkg = F.interpolate(self.pos_embed.float(), size=tuple(hw), mode="bicubic")
            win = self.pos_embed_window.float()
            reps = [a // b for a, b in zip(bkg.shape, win.shape)]
            self._pos_cache = {"key": key, "val": (bkg + win.tile(reps)).permute(0, 2, 3, 1).to(self.pos_embed.dtype)}
        return self._pos_cache["val"]

    def forward(self, x):
        x = self.patch_embed(x)
        x = x + self._get_pos_embed(x.shape[1:3])
        outs = []
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if i == self.stage_ends[-1] or (i in self.stage_ends and self.return_interm_layers):
                outs.append(x.permute(0, 3, 1, 2))
        return outs

    def get_num_layers(self):
        return len(self.blocks)
