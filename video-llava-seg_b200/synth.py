"""Deterministic synthetic weights and backbone features for the SAM2 propagation hot path.

There is no network for checkpoints or datasets, so tests, bench.py and the golden-vector
generator all draw the *same* seeded random-init weights (keys/shapes identical to the
reference state_dict for everything outside `image_encoder.*`) and the same synthetic clips
from the CPU generator here.  torch's CPU Philox/MT generator is bit-reproducible across hosts
for a fixed torch build, which is what lets golden vectors produced by the reference in the
build container be compared on the GPU box.

The image encoder (Hiera + FPN) is outside the hot path: clips
are given as the tensors `SAM2Base.forward_image` + `_prepare_backbone_features` would produce
(sam2_base.py:467-495): vision_feat / vision_pos [4096,1,256], feat_s0 [1,32,256,256] and
feat_s1 [1,64,128,128] (already through conv_s0/conv_s1).
"""
import math

import torch
import torch.nn.functional as F


def _attn_shapes(p, emb, internal, kv_in=None):
    kv_in = kv_in or emb
    return {
        f"{p}.q_proj.weight": (internal, emb), f"{p}.q_proj.bias": (internal,),
        f"{p}.k_proj.weight": (internal, kv_in), f"{p}.k_proj.bias": (internal,),
        f"{p}.v_proj.weight": (internal, kv_in), f"{p}.v_proj.bias": (internal,),
        f"{p}.out_proj.weight": (emb, internal), f"{p}.out_proj.bias": (emb,),
    }


def _ln(p, c):
    return {f"{p}.weight": (c,), f"{p}.bias": (c,)}


def _mlp(p, dims):
    out = {}
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        out[f"{p}.layers.{i}.weight"] = (b, a)
        out[f"{p}.layers.{i}.bias"] = (b,)
    return out


def hot_path_param_shapes():
    """Reference state_dict keys -> shapes for everything except image_encoder.* (identical for
    sam2.1_hiera_{t,s,b+,l}.yaml; verified against the reference in tests/golden/make_golden.py)."""
    s = {
        "maskmem_tpos_enc": (7, 1, 1, 64), "no_mem_embed": (1, 1, 256), "no_mem_pos_enc": (1, 1, 256),
        "no_obj_ptr": (1, 256), "no_obj_embed_spatial": (1, 64),
        "mask_downsample.weight": (1, 1, 4, 4), "mask_downsample.bias": (1,),
    }
    for i in range(4):
        lp = f"memory_attention.layers.{i}"
        s.update(_attn_shapes(lp + ".self_attn", 256, 256))
        s.update(_attn_shapes(lp + ".cross_attn_image", 256, 256, 64))
        s.update({lp + ".linear1.weight": (2048, 256), lp + ".linear1.bias": (2048,),
                  lp + ".linear2.weight": (256, 2048), lp + ".linear2.bias": (256,)})
        for n in ("norm1", "norm2", "norm3"):
            s.update(_ln(f"{lp}.{n}", 256))
    s.update(_ln("memory_attention.norm", 256))
    e = "memory_encoder.mask_downsampler.encoder"
    ch = [1, 4, 16, 64, 256]
    for i in range(4):
        s[f"{e}.{3 * i}.weight"] = (ch[i + 1], ch[i], 3, 3)
        s[f"{e}.{3 * i}.bias"] = (ch[i + 1],)
        s.update(_ln(f"{e}.{3 * i + 1}", ch[i + 1]))
    s[e + ".12.weight"] = (256, 256, 1, 1)
    s[e + ".12.bias"] = (256,)
    s["memory_encoder.pix_feat_proj.weight"] = (256, 256, 1, 1)
    s["memory_encoder.pix_feat_proj.bias"] = (256,)
    for i in range(2):
        lp = f"memory_encoder.fuser.layers.{i}"
        s.update({lp + ".gamma": (256,), lp + ".dwconv.weight": (256, 1, 7, 7), lp + ".dwconv.bias": (256,),
                  lp + ".pwconv1.weight": (1024, 256), lp + ".pwconv1.bias": (1024,),
                  lp + ".pwconv2.weight": (256, 1024), lp + ".pwconv2.bias": (256,)})
        s.update(_ln(lp + ".norm", 256))
    s["memory_encoder.out_proj.weight"] = (64, 256, 1, 1)
    s["memory_encoder.out_proj.bias"] = (64,)
    pe = "sam_prompt_encoder"
    s[pe + ".pe_layer.positional_encoding_gaussian_matrix"] = (2, 128)
    for i in range(4):
        s[f"{pe}.point_embeddings.{i}.weight"] = (1, 256)
    s[pe + ".not_a_point_embed.weight"] = (1, 256)
    s.update({pe + ".mask_downscaling.0.weight": (4, 1, 2, 2), pe + ".mask_downscaling.0.bias": (4,),
              pe + ".mask_downscaling.3.weight": (16, 4, 2, 2), pe + ".mask_downscaling.3.bias": (16,),
              pe + ".mask_downscaling.6.weight": (256, 16, 1, 1), pe + ".mask_downscaling.6.bias": (256,)})
    s.update(_ln(pe + ".mask_downscaling.1", 4))
    s.update(_ln(pe + ".mask_downscaling.4", 16))
    s[pe + ".no_mask_embed.weight"] = (1, 256)
    md = "sam_mask_decoder"
    for i in range(2):
        lp = f"{md}.transformer.layers.{i}"
        s.update(_attn_shapes(lp + ".self_attn", 256, 256))
        s.update(_attn_shapes(lp + ".cross_attn_token_to_image", 256, 128))
        s.update(_attn_shapes(lp + ".cross_attn_image_to_token", 256, 128))
        s.update(_mlp(lp + ".mlp", [256, 2048, 256]))
        for n in ("norm1", "norm2", "norm3", "norm4"):
            s.update(_ln(f"{lp}.{n}", 256))
    s.update(_attn_shapes(md + ".transformer.final_attn_token_to_image", 256, 128))
    s.update(_ln(md + ".transformer.norm_final_attn", 256))
    s.update({md + ".iou_token.weight": (1, 256), md + ".mask_tokens.weight": (4, 256),
              md + ".obj_score_token.weight": (1, 256),
              md + ".output_upscaling.0.weight": (256, 64, 2, 2), md + ".output_upscaling.0.bias": (64,),
              md + ".output_upscaling.3.weight": (64, 32, 2, 2), md + ".output_upscaling.3.bias": (32,),
              md + ".conv_s0.weight": (32, 256, 1, 1), md + ".conv_s0.bias": (32,),
              md + ".conv_s1.weight": (64, 256, 1, 1), md + ".conv_s1.bias": (64,)})
    s.update(_ln(md + ".output_upscaling.1", 64))
    for i in range(4):
        s.update(_mlp(f"{md}.output_hypernetworks_mlps.{i}", [256, 256, 256, 32]))
    s.update(_mlp(md + ".iou_prediction_head", [256, 256, 256, 4]))
    s.update(_mlp(md + ".pred_obj_score_head", [256, 256, 256, 1]))
    s.update(_mlp("obj_ptr_proj", [256, 256, 256, 256]))
    s["obj_ptr_tpos_proj.weight"] = (64, 256)
    s["obj_ptr_tpos_proj.bias"] = (64,)
    return s


_EMBED_KEYS = ("point_embeddings", "not_a_point_embed", "no_mask_embed", "iou_token", "mask_tokens",
               "obj_score_token")
_TRUNC02_KEYS = ("maskmem_tpos_enc", "no_mem_embed", "no_mem_pos_enc", "no_obj_ptr", "no_obj_embed_spatial")


def init_state_dict(seed: int = 0, obj_score_bias: float = 0.75):
    """Seeded random init in the style of torch defaults (U(+-1/sqrt(fan_in)) for Linear/Conv,
    N(0,1) embeddings, trunc-normal(0.02) tokens) but with NON-trivial LayerNorm affines and
    CXBlock layer-scale gamma ~ 0.1 so that every kernel on the path is numerically exercised
    (the reference's default gamma=1e-6 makes the fuser an identity).
    `obj_score_bias` shifts the object-score head so the object gate (sam2_base.py:360) has a
    margin well above bf16 noise."""
    g = torch.Generator().manual_seed(seed)
    shapes = hot_path_param_shapes_cached()
    sd = {}
    for k, shp in shapes.items():
        parent = k.rsplit(".", 1)[0]
        is_ln = parent.rsplit(".", 1)[-1].startswith("norm") or _is_ln2d(k)
        if k in _TRUNC02_KEYS:
            v = (torch.randn(shp, generator=g) * 0.02).clamp_(-0.04, 0.04)
        elif k.endswith("positional_encoding_gaussian_matrix") or any(e in k for e in _EMBED_KEYS):
            v = torch.randn(shp, generator=g)
        elif k.endswith(".gamma"):
            v = 0.1 + 0.03 * torch.randn(shp, generator=g)
        elif is_ln and k.endswith(".weight"):
            v = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif is_ln and k.endswith(".bias"):
            v = 0.05 * torch.randn(shp, generator=g)
        else:
            w = shapes[parent + ".weight"]
            fan_in = 1
            for d in w[1:]:
                fan_in *= d  # Linear: in; Conv: in/groups*k*k; ConvTranspose [in,out,k,k]: out*k*k (torch)
            v = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
        sd[k] = v.float()
    sd["sam_mask_decoder.pred_obj_score_head.layers.2.bias"] += obj_score_bias
    return sd


_SHAPES = None


def hot_path_param_shapes_cached():
    global _SHAPES
    if _SHAPES is None:
        _SHAPES = hot_path_param_shapes()
    return _SHAPES


_LN2D = ("mask_downsampler.encoder.1.", "mask_downsampler.encoder.4.", "mask_downsampler.encoder.7.",
         "mask_downsampler.encoder.10.", "mask_downscaling.1.", "mask_downscaling.4.", "output_upscaling.1.")


def _is_ln2d(k):
    return any(t in k for t in _LN2D)


def sine_pos_256(h=64, w=64):
    """The FPN neck's PositionEmbeddingSine(256) output for one [256,h,w] level
    (backbones/image_encoder.py:133, position_encoding.py:78-112), as [h*w,1,256]."""
    half, eps, scale, temp = 128, 1e-6, 2 * math.pi, 10000.0
    y = torch.arange(1, h + 1, dtype=torch.float32).view(h, 1).expand(h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32).view(1, w).expand(h, w)
    y = y / (y[-1:, :] + eps) * scale
    x = x / (x[:, -1:] + eps) * scale
    dim_t = temp ** (2 * (torch.arange(half, dtype=torch.float32) // 2) / half)
    px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(h * w, 1, 256).contiguous()


class SyntheticClip:
    """A seeded synthetic clip of backbone features: a smooth static scene, a feature-space
    'object' (a Gaussian blob moving 20 px / 10 px per frame in 1024-space, as in SURVEY.md
    section 8d config 1) and small per-frame noise.  Frame t is generated on demand on CPU."""

    def __init__(self, seed: int, num_frames: int, feat: int = 64):
        self.seed, self.num_frames, self.feat = seed, num_frames, feat
        g = torch.Generator().manual_seed(1000003 * seed + 17)
        self._scene = self._smooth(g, 256, feat, 8) * 0.6
        self._s0 = self._smooth(g, 32, 4 * feat, 16) * 0.4
        self._s1 = self._smooth(g, 64, 2 * feat, 16) * 0.4
        self._obj_dir = torch.randn(256, generator=g) * 0.9
        self._obj_s0 = torch.randn(32, generator=g) * 0.5
        self._obj_s1 = torch.randn(64, generator=g) * 0.5
        self.pos = sine_pos_256(feat, feat)

    @staticmethod
    def _smooth(g, c, size, low):
        z = torch.randn(1, c, low, low, generator=g)
        return F.interpolate(z, size=(size, size), mode="bicubic", align_corners=False)[0]

    def _blob(self, t, size):
        cx = (300.0 + 20.0 * t) / 1024.0 * size
        cy = (500.0 + 10.0 * t) / 1024.0 * size
        sig = 80.0 / 1024.0 * size
        yy = torch.arange(size, dtype=torch.float32).view(size, 1) + 0.5
        xx = torch.arange(size, dtype=torch.float32).view(1, size) + 0.5
        return torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sig * sig))

    def frame(self, t: int, batch: int = 1):
        g = torch.Generator().manual_seed(7919 * self.seed + 104729 * t + 3)
        f = self.feat
        vf = self._scene + self._obj_dir[:, None, None] * self._blob(t, f) \
            + 0.05 * torch.randn(256, f, f, generator=g)
        s0 = self._s0 + self._obj_s0[:, None, None] * self._blob(t, 4 * f) \
            + 0.02 * torch.randn(32, 4 * f, 4 * f, generator=g)
        s1 = self._s1 + self._obj_s1[:, None, None] * self._blob(t, 2 * f) \
            + 0.02 * torch.randn(64, 2 * f, 2 * f, generator=g)
        return {
            "vision_feat": vf.flatten(1).t().reshape(f * f, 1, 256).expand(-1, batch, -1).contiguous(),
            "vision_pos": self.pos.expand(-1, batch, -1).contiguous(),
            "feat_s0": s0[None].expand(batch, -1, -1, -1).contiguous(),
            "feat_s1": s1[None].expand(batch, -1, -1, -1).contiguous(),
        }

    def point_prompt(self, batch: int = 1):
        """One positive click per object at (300+40*o, 500) in 1024-space."""
        pts = torch.tensor([[[300.0 + 40.0 * o, 500.0]] for o in range(batch)])
        return {"point_coords": pts, "point_labels": torch.ones(batch, 1, dtype=torch.int32)}


# ----------------------------------------------------------------------------- image encoder
def init_image_encoder_state_dict(variant: str = "t", seed: int = 0):
    """Seeded weights for the Hiera + FPN image encoder of sam2.1_hiera_<variant>.yaml, keyed like the reference's
    `image_encoder.*` slice WITHOUT the prefix (keys / shapes come from the module itself; the strict load into the
    reference module in tests/golden/make_golden.py verifies them).  U(+-1/sqrt(fan_in)) matrices, non-trivial
    LayerNorm affines, N(0, 0.1) positional embeddings: every block contributes at the 1e-1 level."""
    from .build_sam import build_image_encoder

    shapes = {k: tuple(v.shape) for k, v in build_image_encoder(variant).state_dict().items()}
    g = torch.Generator().manual_seed(7_000_003 * (seed + 1) + sum(map(ord, variant)))
    sd = {}
    for k in sorted(shapes):
        shp = shapes[k]
        leaf = k.rsplit(".", 2)[-2] if k.count(".") else k
        if "pos_embed" in k:
            v = 0.1 * torch.randn(shp, generator=g)
        elif leaf.startswith("norm") and k.endswith(".weight"):
            v = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif leaf.startswith("norm") and k.endswith(".bias"):
            v = 0.05 * torch.randn(shp, generator=g)
        else:
            w = shapes[k.rsplit(".", 1)[0] + ".weight"]
            fan_in = 1
            for d in w[1:]:
                fan_in *= d
            v = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
        sd[k] = v.float()
    return sd


def synthetic_frames(num_frames: int, size: int = 1024, seed: int = 1, start: int = 0):
    """ImageNet-normalised synthetic video frames [T,3,size,size] f32: a low-frequency
    bicubic noise field plus a Gaussian blob (sigma 80 px at 1024) whose centre moves from (300, 500) by (20, 10) px per
    frame, and a little per-frame pixel noise."""
    g = torch.Generator().manual_seed(1_000_003 * seed + 29)
    scene = F.interpolate(torch.randn(1, 3, 12, 12, generator=g), size=(size, size), mode="bicubic", align_corners=False)[0]
    scene = 0.5 + 0.18 * scene
    colour = torch.tensor([0.9, 0.2, -0.4]).view(3, 1, 1)
    yy = torch.arange(size, dtype=torch.float32).view(size, 1) + 0.5
    xx = torch.arange(size, dtype=torch.float32).view(1, size) + 0.5
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    out = []
    for t in range(start, start + num_frames):
        gt = torch.Generator().manual_seed(7919 * seed + 104729 * t + 11)
        cx, cy, sig = (300.0 + 20.0 * t) / 1024.0 * size, (500.0 + 10.0 * t) / 1024.0 * size, 80.0 / 1024.0 * size
        blob = torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sig * sig))
        img = (scene + 0.45 * colour * blob + 0.01 * torch.randn(3, size, size, generator=gt)).clamp_(0.0, 1.0)
        out.append((img - mean) / std)
    return torch.stack(out, 0)
