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

    def _encode(self, pix_nchw, pix_rows, mask, mode, scale, bias, gate, no_obj_embed, want_rows, out_dtype):
        """One C call for the whole module.  Exactly one of pix_nchw [B,256,H,W] / pix_rows [HW,B,256] is given;
        mask is [B,1,16H,16W] (modes 0/1/4) or the low-res logits [B,1,4H,4W] (modes 2/3)."""
        require_cuda(mask)
        dev = mask.device
        if pix_nchw is not None:
            pix = pix_nchw.to(dev)
            B, _, H, W = pix.shape
            layout, ps = 0, LL4(pix.stride(0), pix.stride(1), pix.stride(2), pix.stride(3))
        else:
            pix = pix_rows if pix_rows.stride(2) == 1 else pix_rows.contiguous()
            T, B, _ = pix.shape
            H = W = int(round(math.sqrt(T)))
            layout, ps = 1, LL4(pix.stride(0), pix.stride(1), 0, 0)
        f = 4 if mode in (2, 3) else 16
        assert mask.shape == (B, 1, f * H, f * W), f"mask must be [B,1,{f}H,{f}W], got {tuple(mask.shape)}"
        if self._packed is None or (no_obj_embed is not None and not self._packed[0].no_obj_embed):
            self._packed = _pack.pack_mem_encoder(self._flat_sd(), "", dev, no_obj_embed)
        w = self._packed[0]
        m = mask.float().contiguous()
        out = torch.empty((B, 64, H, W), device=dev, dtype=out_dtype)
        rows = torch.empty((B, H * W, 64), device=dev, dtype=torch.bfloat16) if want_rows else None
        nbytes = lib().vls_mem_encoder_workspace_bytes(B, H, W)
        ws = self._workspace(nbytes, dev)
        check(lib().vls_mem_encoder_forward(
            ctypes_ref(w), ptr(pix), dtype_code(pix), layout, ps, ptr(m), int(mode), float(scale), float(bias),
            ptr(gate), B, H, W, ptr(out), dtype_code(out), ptr(rows), ptr(ws), ws.numel(), stream()),
            "vls_mem_encoder_forward")
        return out, rows

    def forward(self, pix_feat, masks, skip_mask_sigmoid=False):
        """pix_feat [B,256,H,W], masks [B,1,16H,16W] -> {"vision_features": [B,64,H,W], "vision_pos_enc": [[B,64,H,W]]}
        (memory_encoder.py:158-181)."""
        out, _ = self._encode(pix_feat, None, masks, 0 if skip_mask_sigmoid else 1, 1.0, 0.0, None, None, False,
                              pix_feat.dtype)
        pos = self.position_encoding(out).to(out.dtype)
        return {"vision_features": out, "vision_pos_enc": [pos]}

    def encode_from_low_res(self, vision_feat_rows, low_res_logits, binarize, sigmoid_scale, sigmoid_bias,
                            occluded_gate=None, no_obj_embed=None, want_rows=True):
        """Fast path used by the predictor (replaces sam2_base.py:372-378 + :676-724): takes the [B,1,4H,4W]
        low-res logits and fuses the x4 bilinear up-sampling and sigmoid*scale+bias (or binarisation) into the first
        convolution, so the [B,1,1024,1024] f32 mask never exists.  vision_feat_rows: [HW,B,256] seq-first.
        Returns (features NCHW bf16 [B,64,H,W], features rows bf16 [B,HW,64] or None)."""
        return self._encode(None, vision_feat_rows, low_res_logits, 3 if binarize else 2, sigmoid_scale, sigmoid_bias,
                            occluded_gate, no_obj_embed, want_rows, torch.bfloat16)
