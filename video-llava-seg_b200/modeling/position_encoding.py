"""Positional-encoding tables.  PositionEmbeddingSine (position_encoding.py:16-112) is a per-shape
constant in the reference too (it caches by (H, W)); PositionEmbeddingRandom (:115-159) belongs to the
prompt encoder, which stays a tiny PyTorch module next to the hot path."""
import math

import numpy as np
import torch
from torch import nn

from .._pack import sine_pe_2d


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats, temperature=10000, normalize=True, scale=None):
        super().__init__()
        assert num_pos_feats % 2 == 0, "Expecting even model width"
        if not normalize or (scale is not None and scale != 2 * math.pi):
            raise NotImplementedError("only the normalised, 2*pi-scaled variant used by SAM 2 is provided")
        self.num_pos_feats, self.temperature, self.normalize, self.scale = num_pos_feats, temperature, True, 2 * math.pi
        self.cache = {}

    @torch.no_grad()
    def forward(self, x):
        key = (x.shape[-2], x.shape[-1], x.device)
        if key not in self.cache:
            self.cache[key] = sine_pe_2d(self.num_pos_feats, x.shape[-2], x.shape[-1], self.temperature).to(x.device)
        return self.cache[key][None].repeat(x.shape[0], 1, 1, 1)


class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats=64, scale=None):
        super().__init__()
        scale = 1.0 if scale is None or scale <= 0.0 else scale
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    def _pe_encoding(self, coords):
        coords = (2 * coords - 1).to(self.positional_encoding_gaussian_matrix.dtype)
        coords = 2 * np.pi * (coords @ self.positional_encoding_gaussian_matrix)
        return torch.cat([torch.sin(coords), torch.cos(coords)], dim=-1)

    def forward(self, size):
        h, w = size
        dev = self.positional_encoding_gaussian_matrix.device
        ys = (torch.arange(h, device=dev, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, device=dev, dtype=torch.float32) + 0.5) / w
        grid = torch.stack([xs[None, :].expand(h, w), ys[:, None].expand(h, w)], dim=-1)
        return self._pe_encoding(grid).permute(2, 0, 1)

    def forward_with_coords(self, coords_input, image_size):
        coords = coords_input.clone()
        coords[:, :, 0] = coords[:, :, 0] / image_size[1]
        coords[:, :, 1] = coords[:, :, 1] / image_size[0]
        return self._pe_encoding(coords.to(torch.float))
