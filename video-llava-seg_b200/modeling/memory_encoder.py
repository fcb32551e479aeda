"""Drop-in `MemoryEncoder` (sam2/modeling/memory_encoder.py): MaskDownSampler / CXBlock / Fuser are
parameter containers with the reference's state_dict keys; MemoryEncoder.forward runs the whole module
in one C call (vls_mem_encoder_forward)."""
import copy
import math

import torch
from torch import nn

from .. import _pack
from .._lib import LL4, check, lib, ptr, stream
from ._base import PackedModule, ctypes_ref, dtype_code, require_cuda
from .sam2_utils import LayerNorm2d


class MaskDownSampler(nn.Module):
    def __init__(self, embed_dim=256, kernel_size=4, stride=4, padding=0, total_stride=16, activation=nn.GELU):
        super().__init__()
        num_layers = int(math.log2(total_stride) // math.log2(stride))
        assert stride ** num_layers == total_stride
        if (embed_dim, kernel_size, stride, padding, total_stride) != (256, 3, 2, 1, 16) or activation is not nn.GELU:
            raise NotImplementedError("CUDA path implements the SAM 2.1 mask down-sampler: k3 s2 p1, stride 16, GELU")
        self.encoder = nn.Sequential()
        cin = 1
        for _ in range(num_layers):
            cout = cin * stride ** 2
            self.encoder.append(nn.Conv2d(cin, cout, kernel_size=kernel_size, stride=stride, padding=padding))
            self.encoder.append(LayerNorm2d(cout))
            self.encoder.append(activation())
            cin = cout
        self.encoder.append(nn.Conv2d(cin, embed_dim, kernel_size=1))


class CXBlock(nn.Module):
    def __init__(self, dim, kernel_size=7, padding=3, drop_path=0.0, layer_scale_init_value=1e-6, use_dwconv=True):
        super().__init__()
        if (dim, kernel_size, padding) != (256, 7, 3) or not use_dwconv:
            raise NotImplementedError("CUDA path implements the SAM 2.1 fuser block: depth-wise 7x7 on 256 channels")
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=kernel_size, padding=padding, groups=dim)
        self.norm = LayerNorm2d(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None


class Fuser(nn.Module):
    def __init__(self, layer, num_layers, dim=None, input_projection=False):
        super().__init__()
        if num_layers != 2 or input_projection:
            raise NotImplementedError("CUDA path implements the SAM 2.1 fuser: two CXBlocks, no input projection")
        self.proj = nn.Identity()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])


class MemoryEncoder(PackedModule):
    def __init__(self, out_dim, mask_downsampler, fuser, position_encoding, in_dim=256):
        super().__init__()
        if (out_dim, in_dim) != (64, 256):
            raise NotImplementedError("CUDA path implements in_dim 256 -> out_dim 64")
        self.mask_downsampler = mask_downsampler
        self.pix_feat_proj = nn.Conv2d(in_dim, in_dim, kernel_size=1)
        self.fuser = fuser
        self.position_encoding = position_encoding
        self.out_proj = nn.Conv2d(in_dim, out_dim, kernel_size=1)

    def _pack(self, dev):
        if self._packed is None:
            self._packed = _pack.pack_mem_encoder(self._flat_sd(), "", dev)
        return self._packed[0]

    def forward(self, pix_feat, masks, skip_mask_sigmoid=False):
        """pix_feat [B,256,H,W], masks [B,1,16H,16W] -> {"vision_features": [B,64,H,W], "vision_pos_enc": [[B,64,H,W]]}
        (memory_encoder.py:158-181)."""
        require_cuda(masks)
        dev = masks.device
        pix_feat = pix_feat.to(dev)
        B, C, H, W = pix_feat.shape
        assert masks.shape == (B, 1, 16 * H, 16 * W), "mask must be 16x the feature resolution"
        w = self._pack(dev)
        m = masks.float().contiguous()
        ps = LL4(pix_feat.stride(0), pix_feat.stride(1), pix_feat.stride(2), pix_feat.stride(3))
        out = torch.empty((B, 64, H, W), device=dev, dtype=pix_feat.dtype)
        nbytes = lib().vls_mem_encoder_workspace_bytes(B, H, W)
        ws = self._workspace(nbytes, dev)
        check(lib().vls_mem_encoder_forward(
            ctypes_ref(w), ptr(pix_feat), dtype_code(pix_feat), 0, ps, ptr(m), 0 if skip_mask_sigmoid else 1, 1.0, 0.0,
            None, B, H, W, ptr(out), dtype_code(out), None, ptr(ws), ws.numel(), stream()), "vls_mem_encoder_forward")
        pos = self.position_encoding(out).to(out.dtype)
        return {"vision_features": out, "vision_pos_enc": [pos]}

    def encode_from_low_res(self, vision_feat_rows, low_res_logits, binarize, sigmoid_scale, sigmoid_bias,
                            occluded_gate=None, no_obj_embed=None, want_rows=True):
        """Fast path used by the predictor (replaces sam2_base.py:372-378 + :676-724): takes the [B,1,4H,4W]
        low-res logits and fuses the x4 bilinear up-sampling, sigmoid*scale+bias (or binarisation) into the first
        convolution, so the [B,1,1024,1024] f32 mask never exists.  vision_feat_rows: [HW,B,256] seq-first.
        Returns (features NCHW bf16 [B,64,H,W], features rows bf16 [B,HW,64] or None)."""
        require_cuda(vision_feat_rows, low_res_logits)
        dev = low_res_logits.device
        T, B, C = vision_feat_rows.shape
        H = W = int(round(math.sqrt(T)))
        assert low_res_logits.shape == (B, 1, 4 * H, 4 * W)
        if self._packed is None or (no_obj_embed is not None and not self._packed[0].no_obj_embed):
            self._packed = _pack.pack_mem_encoder(self._flat_sd(), "", dev, no_obj_embed)
        w = self._packed[0]
        vf = vision_feat_rows if vision_feat_rows.stride(2) == 1 else vision_feat_rows.contiguous()
        ps = LL4(vf.stride(0), vf.stride(1), 0, 0)
        lo = low_res_logits.float().contiguous()
        out = torch.empty((B, 64, H, W), device=dev, dtype=torch.bfloat16)
        rows = torch.empty((B, T, 64), device=dev, dtype=torch.bfloat16) if want_rows else None
        nbytes = lib().vls_mem_encoder_workspace_bytes(B, H, W)
        ws = self._workspace(nbytes, dev)
        check(lib().vls_mem_encoder_forward(
            ctypes_ref(w), ptr(vf), dtype_code(vf), 1, ps, ptr(lo), 3 if binarize else 2, float(sigmoid_scale),
            float(sigmoid_bias), ptr(occluded_gate), B, H, W, ptr(out), 1, ptr(rows), ptr(ws), ws.numel(), stream()),
            "vls_mem_encoder_forward")
        return out, rows
