"""Image encoder = trunk + FPN neck (sam2/modeling/backbones/image_encoder.py:14-136), hosted in PyTorch (SURVEY
section 8 row f-4, first step: bf16 + CUDA graph, no new kernels), plus GraphedImageEncoder: the whole
`SAM2Base.forward_image` (trunk, neck, conv_s0 / conv_s1; sam2_base.py:467-479) captured once per frame shape."""
import torch
import torch.nn.functional as F
from torch import nn


class FpnNeck(nn.Module):
    """1x1 lateral convolutions + top-down nearest/bilinear x2 sums on the levels in `fpn_top_down_levels`; returns the
    feature maps and their sine position encodings, highest resolution first (image_encoder.py:44-136)."""

    def __init__(self, position_encoding, d_model, backbone_channel_list, kernel_size=1, stride=1, padding=0,
                 fpn_interp_model="bilinear", fuse_type="sum", fpn_top_down_levels=None):
        super().__init__()
        assert fuse_type in ("sum", "avg")
        self.position_encoding = position_encoding
        self.backbone_channel_list, self.d_model = list(backbone_channel_list), d_model
        self.convs = nn.ModuleList()
        for c in self.backbone_channel_list:
            lateral = nn.Sequential()
            lateral.add_module("conv", nn.Conv2d(c, d_model, kernel_size=kernel_size, stride=stride, padding=padding))
            self.convs.append(lateral)
        self.fpn_interp_model, self.fuse_type = fpn_interp_model, fuse_type
        self.fpn_top_down_levels = list(range(len(self.convs)) if fpn_top_down_levels is None else fpn_top_down_levels)

    def forward(self, xs):
        n = len(self.convs) - 1
        assert len(xs) == n + 1
        out, pos, carried = [None] * (n + 1), [None] * (n + 1), None
        for i in range(n, -1, -1):                      # lowest resolution first
            lateral = self.convs[n - i](xs[i])
            if i in self.fpn_top_down_levels and carried is not None:
                up = F.interpolate(carried.float(), scale_factor=2.0, mode=self.fpn_interp_model,
                                   align_corners=None if self.fpn_interp_model == "nearest" else False, antialias=False)
                carried = lateral + up.to(lateral.dtype)
                if self.fuse_type == "avg":
                    carried = carried / 2
            else:
                carried = lateral
            out[i] = carried
            pos[i] = self.position_encoding(carried).to(carried.dtype)
        return out, pos


class ImageEncoder(nn.Module):
    def __init__(self, trunk, neck, scalp=0):
        super().__init__()
        self.trunk, self.neck, self.scalp = trunk, neck, scalp
        assert list(trunk.channel_list) == list(neck.backbone_channel_list), \
            f"Channel dims of trunk and neck do not match. Trunk: {trunk.channel_list}, neck: {neck.backbone_channel_list}"

    def forward(self, sample):
        feats, pos = self.neck(self.trunk(sample))
        if self.scalp > 0:                              # drop the lowest-resolution level(s)
            feats, pos = feats[:-self.scalp], pos[:-self.scalp]
        return {"vision_features": feats[-1], "vision_pos_enc": pos, "backbone_fpn": feats}


class GraphedImageEncoder(nn.Module):
    """`forward_image` in a reduced precision under a CUDA graph.  Holds a bf16 copy of an ImageEncoder plus the mask
    decoder's conv_s0 / conv_s1, a static input buffer per batch shape, and replays one captured graph per shape.  The
    returned tensors are the graph's static outputs: consume (or copy) them before the next call -- the predictor copies
    them into its own static frame inputs (graphed.py) or uses them within the frame.

    Returns the dict of SAM2Base.forward_image with conv_s0 / conv_s1 ALREADY applied (`applies_high_res_convs`)."""

    applies_high_res_convs = True

    def __init__(self, encoder, conv_s0, conv_s1, dtype=torch.bfloat16, use_graph=True):
        super().__init__()
        import copy

        self.encoder = copy.deepcopy(encoder).to(dtype).eval()
        self.conv_s0 = copy.deepcopy(conv_s0).to(dtype).eval()
        self.conv_s1 = copy.deepcopy(conv_s1).to(dtype).eval()
        self.dtype, self.use_graph = dtype, use_graph
        self.neck = self.encoder.neck                    # SAM2Base reads image_encoder.neck.d_model
        self._graphs = {}

    def _eager(self, x):
        out = self.encoder(x)
        out["backbone_fpn"][0] = self.conv_s0(out["backbone_fpn"][0])
        out["backbone_fpn"][1] = self.conv_s1(out["backbone_fpn"][1])
        out["vision_features"] = out["backbone_fpn"][-1]
        return out

    @torch.inference_mode()      # the static buffers are inference tensors whoever calls first
    def forward(self, img_batch):
        if not (self.use_graph and img_batch.is_cuda):
            return self._eager(img_batch.to(self.dtype))
        key = (tuple(img_batch.shape), img_batch.device)
        g = self._graphs.get(key)
        if g is None:
            static_in = torch.zeros(img_batch.shape, dtype=self.dtype, device=img_batch.device)
            static_in.copy_(img_batch)
            side = torch.cuda.Stream(device=img_batch.device)
            side.wait_stream(torch.cuda.current_stream(img_batch.device))
            with torch.cuda.stream(side):
                for _ in range(2):                        # warm up: cuDNN / cuBLAS plans, SDPA back-end selection
                    self._eager(static_in)
            torch.cuda.current_stream(img_batch.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._eager(static_in)
            g = self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = g
        static_in.copy_(img_batch, non_blocking=True)     # dtype conversion (e.g. f32 / uint8-normalised -> bf16) included
        graph.replay()
        return {"vision_features": static_out["vision_features"], "vision_pos_enc": list(static_out["vision_pos_enc"]),
                "backbone_fpn": list(static_out["backbone_fpn"])}
