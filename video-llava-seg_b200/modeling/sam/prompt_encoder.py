"""PromptEncoder (sam2/modeling/sam/prompt_encoder.py): 6 220 parameters, a handful of [B,P,256] ops that
feed the decoder.  It is adjacent to, not on, the hot path and stays PyTorch;
state_dict keys, constructor and forward contract match the reference."""
import torch
from torch import nn

from ..position_encoding import PositionEmbeddingRandom
from ..sam2_utils import LayerNorm2d


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim, image_embedding_size, input_image_size, mask_in_chans, activation=nn.GELU):
        super().__init__()
        self.embed_dim, self.input_image_size, self.image_embedding_size = embed_dim, input_image_size, image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings = 4  # neg / pos point, box top-left / bottom-right
        self.point_embeddings = nn.ModuleList(nn.Embedding(1, embed_dim) for _ in range(self.num_point_embeddings))
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans // 4), activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans),
            activation(), nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1))
        self.no_mask_embed = nn.Embedding(1, embed_dim)
        self._dense_pe = None
        self._pe_version = 0
        self.register_load_state_dict_post_hook(lambda m, keys: m._drop_dense_pe())

    def _drop_dense_pe(self):
        """An in-place load_state_dict keeps data_ptr, so the cache cannot be keyed on pointers alone."""
        self._dense_pe = None
        self._pe_version += 1

    def _apply(self, fn, *a, **kw):
        self._dense_pe = None
        return super()._apply(fn, *a, **kw)

    def get_dense_pe(self):
        """[1, embed_dim, h, w]; cached so the decoder's PE-dependent constants are packed only once."""
        g = self.pe_layer.positional_encoding_gaussian_matrix
        key = (g.data_ptr(), g.device)
        if self._dense_pe is None or self._dense_pe[0] != key:
            pe = self.pe_layer(self.image_embedding_size).unsqueeze(0)
            pe._vls_version = self._pe_version       # MaskDecoder keys its PE-dependent packed constants on it
            self._dense_pe = (key, pe)
        return self._dense_pe[1]

    def _embed_points(self, points, labels, pad):
        points = points + 0.5
        if pad:
            points = torch.cat([points, torch.zeros((points.shape[0], 1, 2), device=points.device)], dim=1)
            labels = torch.cat([labels, -torch.ones((labels.shape[0], 1), device=labels.device, dtype=labels.dtype)], dim=1)
        emb = self.pe_layer.forward_with_coords(points, self.input_image_size)
        emb = torch.where((labels == -1)[..., None], self.not_a_point_embed.weight.expand_as(emb), emb)
        for i in range(self.num_point_embeddings):
            emb = emb + (labels == i)[..., None].to(emb.dtype) * self.point_embeddings[i].weight
        return emb

    def _embed_boxes(self, boxes):
        corners = self.pe_layer.forward_with_coords((boxes + 0.5).reshape(-1, 2, 2), self.input_image_size)
        add = torch.stack([self.point_embeddings[2].weight[0], self.point_embeddings[3].weight[0]], 0)
        return corners + add[None]

    def _embed_masks(self, masks):
        x = masks
        for m in self.mask_downscaling:
            if isinstance(m, LayerNorm2d):
                u = x.mean(1, keepdim=True)
                s = (x - u).pow(2).mean(1, keepdim=True)
                x = m.weight[:, None, None] * ((x - u) / torch.sqrt(s + m.eps)) + m.bias[:, None, None]
            else:
                x = m(x)
        return x

    def forward(self, points, boxes, masks):
        bs = points[0].shape[0] if points is not None else boxes.shape[0] if boxes is not None else \
            masks.shape[0] if masks is not None else 1
        dev = self.point_embeddings[0].weight.device
        sparse = torch.empty((bs, 0, self.embed_dim), device=dev)
        if points is not None:
            sparse = torch.cat([sparse, self._embed_points(points[0], points[1], pad=(boxes is None))], dim=1)
        if boxes is not None:
            sparse = torch.cat([sparse, self._embed_boxes(boxes)], dim=1)
        if masks is not None:
            dense = self._embed_masks(masks)
        else:
            dense = self.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(bs, -1, *self.image_embedding_size)
        return sparse, dense
