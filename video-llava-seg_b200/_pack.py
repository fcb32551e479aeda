"""ctypes mirrors of the weight structs in include/vls_b200.h and the packers that turn a flat
state_dict with the reference's key names (build_sam.py:141-151 loads those strictly) into the
device-resident bf16/f32 layouts the kernels expect.  Packing happens once per weight set; the
returned objects keep every device tensor alive."""
import ctypes
import math

import torch

from . import ops
from ._lib import c_void_p, c_int, ptr

PF = ctypes.POINTER(ctypes.c_float)


def _fields(*names):
    return [(n, c_void_p) for n in names]


class MemAttnLayer(ctypes.Structure):
    _fields_ = _fields("sa_qk_w", "sa_qk_b", "sa_v_w", "sa_v_b", "sa_o_w", "sa_o_b", "ca_q_w", "ca_q_b", "ca_ov_w",
                       "ca_ov_b", "l1_w", "l1_b", "l2_w", "l2_b", "n1_w", "n1_b", "n2_w", "n2_b", "n3_w", "n3_b")


class MemAttnWeights(ctypes.Structure):
    _fields_ = [("num_layers", c_int), ("layers", MemAttnLayer * 8), ("norm_w", c_void_p), ("norm_b", c_void_p),
                ("ca_k_w_all", c_void_p), ("ca_k_b_all", c_void_p),
                ("rope_cos", c_void_p), ("rope_sin", c_void_p), ("rope_len", c_int)]


class AttnW(ctypes.Structure):
    _fields_ = _fields("q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")


class DecLayer(ctypes.Structure):
    _fields_ = [("self_attn", AttnW), ("t2i", AttnW), ("i2t", AttnW)] + _fields(
        "img_w", "img_b", "img_pe_add", "mlp1_w", "mlp1_b", "mlp2_w", "mlp2_b", "n1_w", "n1_b", "n2_w", "n2_b", "n3_w",
        "n3_b", "n4_w", "n4_b")


class MaskDecoderWeights(ctypes.Structure):
    _fields_ = [("layers", DecLayer * 2), ("final_t2i", AttnW)] + _fields(
        "final_img_w", "final_img_b", "final_pe_add", "nf_w", "nf_b", "out_tokens", "up1_w", "up1_b", "up_ln_w",
        "up_ln_b", "up2_w", "up2_b", "up2_wh") + [("hyper_w", c_void_p * 3), ("hyper_b", c_void_p * 3), ("iou_w", c_void_p * 3),
                                        ("iou_b", c_void_p * 3), ("obj_w", c_void_p * 3), ("obj_b", c_void_p * 3),
                                        ("iou_sigmoid", c_int)]


class ObjPtrWeights(ctypes.Structure):
    _fields_ = [("w", c_void_p * 3), ("b", c_void_p * 3), ("no_obj_ptr", c_void_p)]


class CxBlock(ctypes.Structure):
    _fields_ = _fields("dw_w", "dw_b", "ln_w", "ln_b", "pw1_w", "pw1_b", "pw2_w", "pw2_b")


class MemEncoderWeights(ctypes.Structure):
    _fields_ = _fields("c1_w", "c1_b", "ln1_w", "ln1_b", "c2_w", "c2_b", "ln2_w", "ln2_b", "c3_w", "c3_b", "ln3_w",
                       "ln3_b", "c3_wh", "c4_w", "c4_b", "ln4_w", "ln4_b", "c5_w", "c5_b", "pix_w", "pix_b") + [
                           ("cx", CxBlock * 2)] + _fields("out_w", "out_b", "no_obj_embed")


class _Keep:
    """Uploads tensors and remembers them so the raw pointers in the ctypes struct stay valid."""

    def __init__(self, device):
        self.device = device
        self.tensors = []

    def h(self, t):  # bf16 weight
        t = t.detach().to(self.device, torch.float32).to(torch.bfloat16).contiguous()
        self.tensors.append(t)
        return t.data_ptr()

    def f(self, t):  # f32 vector / table
        t = t.detach().to(self.device, torch.float32).contiguous()
        self.tensors.append(t)
        return t.data_ptr()

    def keep(self, t):
        self.tensors.append(t)
        return t.data_ptr()


def axial_rope_tables(nq, device, dim=256, theta=10000.0):
    """cos/sin [nq, dim/2] of compute_axial_cis (position_encoding.py:168-184) for a square grid.
    A model constant (depends only on the grid size), computed once per size and cached by the caller."""
    side = int(round(math.sqrt(nq)))
    if side * side != nq:
        raise ValueError(f"RoPE needs a square token grid, got {nq} tokens")
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    t = torch.arange(nq, dtype=torch.float32)
    ang = torch.cat([torch.outer((t % side).float(), freqs),
                     torch.outer(torch.div(t, side, rounding_mode="floor").float(), freqs)], dim=-1)
    return torch.cos(ang).to(device).contiguous(), torch.sin(ang).to(device).contiguous()


def sine_pe_2d(num_pos_feats, h, w, temperature=10000.0):
    """PositionEmbeddingSine table [C,h,w] (position_encoding.py:78-112): a per-shape constant."""
    half, eps, scale = num_pos_feats // 2, 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32).view(h, 1).expand(h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32).view(1, w).expand(h, w)
    y = y / (y[-1:, :] + eps) * scale
    x = x / (x[:, -1:] + eps) * scale
    dim_t = temperature ** (2 * (torch.arange(half, dtype=torch.float32) // 2) / half)
    px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).permute(2, 0, 1).contiguous()


# ----------------------------------------------------------------------------- memory attention
def pack_mem_attn(sd, prefix, device):
    k = _Keep(device)
    w = MemAttnWeights()
    n = 0
    while f"{prefix}layers.{n}.norm1.weight" in sd:
        n += 1
    if not 1 <= n <= 8:
        raise ValueError(f"memory attention with {n} layers is not supported")
    w.num_layers = n
    for i in range(n):
        p = f"{prefix}layers.{i}."
        L = w.layers[i]
        sa, ca = p + "self_attn.", p + "cross_attn_image."
        if tuple(sd[sa + "q_proj.weight"].shape) != (256, 256) or tuple(sd[ca + "k_proj.weight"].shape) != (256, 64) \
                or tuple(sd[p + "linear1.weight"].shape) != (2048, 256):
            raise NotImplementedError("only the SAM 2.1 memory-attention geometry (d_model 256, kv_in 64, FFN 2048, "
                                      "one head) is implemented")
        L.sa_qk_w = k.h(torch.cat([sd[sa + "q_proj.weight"], sd[sa + "k_proj.weight"]], 0))
        L.sa_qk_b = k.f(torch.cat([sd[sa + "q_proj.bias"], sd[sa + "k_proj.bias"]], 0))
        L.sa_v_w, L.sa_v_b = k.h(sd[sa + "v_proj.weight"]), k.f(sd[sa + "v_proj.bias"])
        L.sa_o_w, L.sa_o_b = k.h(sd[sa + "out_proj.weight"]), k.f(sd[sa + "out_proj.bias"])
        L.ca_q_w, L.ca_q_b = k.h(sd[ca + "q_proj.weight"]), k.f(sd[ca + "q_proj.bias"])
        # value + output projection folded (softmax rows sum to one): the kernel attends over the raw 64-d memory
        wo, wv = sd[ca + "out_proj.weight"].double(), sd[ca + "v_proj.weight"].double()
        L.ca_ov_w = k.h((wo @ wv).float())
        L.ca_ov_b = k.f((wo @ sd[ca + "v_proj.bias"].double() + sd[ca + "out_proj.bias"].double()).float())
        L.l1_w, L.l1_b = k.h(sd[p + "linear1.weight"]), k.f(sd[p + "linear1.bias"])
        L.l2_w, L.l2_b = k.h(sd[p + "linear2.weight"]), k.f(sd[p + "linear2.bias"])
        for j in (1, 2, 3):
            setattr(L, f"n{j}_w", k.f(sd[f"{p}norm{j}.weight"]))
            setattr(L, f"n{j}_b", k.f(sd[f"{p}norm{j}.bias"]))
    w.norm_w, w.norm_b = k.f(sd[prefix + "norm.weight"]), k.f(sd[prefix + "norm.bias"])
    ca = [f"{prefix}layers.{i}.cross_attn_image." for i in range(n)]
    w.ca_k_w_all = k.h(torch.stack([sd[c + "k_proj.weight"] for c in ca], 0))
    w.ca_k_b_all = k.f(torch.stack([sd[c + "k_proj.bias"] for c in ca], 0))
    return w, k


# ----------------------------------------------------------------------------- mask decoder
def _attn(k, sd, p, dst, parts="qkvo"):
    for c, name in (("q", "q_proj"), ("k", "k_proj"), ("v", "v_proj"), ("o", "out_proj")):
        if c in parts:
            setattr(dst, c + "_w", k.h(sd[f"{p}{name}.weight"]))
            setattr(dst, c + "_b", k.f(sd[f"{p}{name}.bias"]))


def _pe_times_wt(pe_rows, wt):
    """pe_rows [T,256] f32 (device) x wt [N,256] -> [T,N] f32 through the tcgen05 GEMM.  pe is split
    into bf16 hi + lo parts so the constant keeps ~16 mantissa bits (weights are bf16 on the path anyway)."""
    hi = pe_rows.to(torch.bfloat16)
    lo = (pe_rows - hi.float()).to(torch.bfloat16)
    wb = wt.to(pe_rows.device, torch.float32).to(torch.bfloat16).contiguous()
    out = ops.gemm(hi.contiguous(), wb, out_dtype=torch.float32)
    return ops.gemm(lo.contiguous(), wb, residual=out, out_dtype=torch.float32)


def pack_mask_decoder(sd, prefix, device, image_pe, iou_sigmoid=True):
    """image_pe: [1,256,H,W] dense positional encoding (prompt_encoder.get_dense_pe())."""
    k = _Keep(device)
    w = MaskDecoderWeights()
    if tuple(sd[prefix + "mask_tokens.weight"].shape) != (4, 256) or (prefix + "obj_score_token.weight") not in sd \
            or (prefix + "pred_obj_score_head.layers.2.weight") not in sd:
        raise NotImplementedError("only the SAM 2.1 decoder head layout (4 mask tokens, obj-score MLP) is implemented")
    pe_rows = image_pe.detach().to(device, torch.float32).flatten(2)[0].t().contiguous()  # [T,256]
    T = pe_rows.shape[0]
    tp = prefix + "transformer."
    for i in range(2):
        p = f"{tp}layers.{i}."
        L = w.layers[i]
        _attn(k, sd, p + "self_attn.", L.self_attn)
        _attn(k, sd, p + "cross_attn_token_to_image.", L.t2i, "qo")
        _attn(k, sd, p + "cross_attn_image_to_token.", L.i2t, "kvo")
        t2i, i2t = p + "cross_attn_token_to_image.", p + "cross_attn_image_to_token."
        L.img_w = k.h(torch.cat([sd[t2i + "k_proj.weight"], sd[t2i + "v_proj.weight"], sd[i2t + "q_proj.weight"]], 0))
        L.img_b = k.f(torch.cat([sd[t2i + "k_proj.bias"], sd[t2i + "v_proj.bias"], sd[i2t + "q_proj.bias"]], 0))
        add = torch.zeros((T, 384), device=device, dtype=torch.float32)
        add[:, 0:128] = _pe_times_wt(pe_rows, sd[t2i + "k_proj.weight"])
        add[:, 256:384] = _pe_times_wt(pe_rows, sd[i2t + "q_proj.weight"])
        L.img_pe_add = k.keep(add)
        L.mlp1_w, L.mlp1_b = k.h(sd[p + "mlp.layers.0.weight"]), k.f(sd[p + "mlp.layers.0.bias"])
        L.mlp2_w, L.mlp2_b = k.h(sd[p + "mlp.layers.1.weight"]), k.f(sd[p + "mlp.layers.1.bias"])
        for j in (1, 2, 3, 4):
            setattr(L, f"n{j}_w", k.f(sd[f"{p}norm{j}.weight"]))
            setattr(L, f"n{j}_b", k.f(sd[f"{p}norm{j}.bias"]))
    fa = tp + "final_attn_token_to_image."
    _attn(k, sd, fa, w.final_t2i, "qo")
    w.final_img_w = k.h(torch.cat([sd[fa + "k_proj.weight"], sd[fa + "v_proj.weight"]], 0))
    w.final_img_b = k.f(torch.cat([sd[fa + "k_proj.bias"], sd[fa + "v_proj.bias"]], 0))
    add = torch.zeros((T, 256), device=device, dtype=torch.float32)
    add[:, 0:128] = _pe_times_wt(pe_rows, sd[fa + "k_proj.weight"])
    w.final_pe_add = k.keep(add)
    w.nf_w, w.nf_b = k.f(sd[tp + "norm_final_attn.weight"]), k.f(sd[tp + "norm_final_attn.bias"])
    w.out_tokens = k.f(torch.cat([sd[prefix + "obj_score_token.weight"], sd[prefix + "iou_token.weight"],
                                  sd[prefix + "mask_tokens.weight"]], 0))
    # ConvTranspose2d weight [ci, co, kh, kw] -> GEMM weight [(dy*2+dx)*co + c][ci]
    up = prefix + "output_upscaling."
    w1 = sd[up + "0.weight"]  # [256, 64, 2, 2]
    w.up1_w = k.h(w1.permute(2, 3, 1, 0).reshape(4 * 64, 256))
    w.up1_b = k.f(sd[up + "0.bias"].repeat(4))
    w.up_ln_w, w.up_ln_b = k.f(sd[up + "1.weight"]), k.f(sd[up + "1.bias"])
    w2 = sd[up + "3.weight"]  # [64 ci, 32 co, 2, 2] -> [pos][ci][co]
    w.up2_w = k.f(w2.permute(2, 3, 0, 1).reshape(4, 64, 32))
    w.up2_b = k.f(sd[up + "3.bias"])
    w2g = w2.permute(2, 3, 1, 0).reshape(4 * 32, 64).float()   # [(dy*2+dx)*32 + co][ci]
    w2hi = w2g.to(torch.bfloat16)
    w.up2_wh = k.h(torch.cat([w2hi.float(), w2g - w2hi.float()], 0))
    hp = prefix + "output_hypernetworks_mlps."
    for j in range(3):
        w.hyper_w[j] = k.h(torch.stack([sd[f"{hp}{m}.layers.{j}.weight"] for m in range(4)], 0))
        w.hyper_b[j] = k.f(torch.stack([sd[f"{hp}{m}.layers.{j}.bias"] for m in range(4)], 0))
        w.iou_w[j] = k.h(sd[f"{prefix}iou_prediction_head.layers.{j}.weight"])
        w.iou_b[j] = k.f(sd[f"{prefix}iou_prediction_head.layers.{j}.bias"])
        w.obj_w[j] = k.h(sd[f"{prefix}pred_obj_score_head.layers.{j}.weight"])
        w.obj_b[j] = k.f(sd[f"{prefix}pred_obj_score_head.layers.{j}.bias"])
    w.iou_sigmoid = int(bool(iou_sigmoid))
    return w, k


def pack_obj_ptr(sd, device, prefix="obj_ptr_proj.", no_obj_ptr_key="no_obj_ptr"):
    k = _Keep(device)
    w = ObjPtrWeights()
    for j in range(3):
        w.w[j] = k.h(sd[f"{prefix}layers.{j}.weight"])
        w.b[j] = k.f(sd[f"{prefix}layers.{j}.bias"])
    w.no_obj_ptr = k.f(sd[no_obj_ptr_key].reshape(-1))
    return w, k


# ----------------------------------------------------------------------------- memory encoder
def pack_mem_encoder(sd, prefix, device, no_obj_embed=None):
    k = _Keep(device)
    w = MemEncoderWeights()
    e = prefix + "mask_downsampler.encoder."
    if tuple(sd[e + "0.weight"].shape) != (4, 1, 3, 3) or tuple(sd[e + "9.weight"].shape) != (256, 64, 3, 3):
        raise NotImplementedError("only the SAM 2.1 mask down-sampler (k3 s2 p1, 1-4-16-64-256) is implemented")
    w.c1_w, w.c1_b = k.f(sd[e + "0.weight"].reshape(4, 9)), k.f(sd[e + "0.bias"])
    w.ln1_w, w.ln1_b = k.f(sd[e + "1.weight"]), k.f(sd[e + "1.bias"])
    w.c2_w = k.f(sd[e + "3.weight"].permute(2, 3, 1, 0).reshape(9, 4, 16))   # [co,ci,ky,kx] -> [tap][ci][co]
    w.c2_b, w.ln2_w, w.ln2_b = k.f(sd[e + "3.bias"]), k.f(sd[e + "4.weight"]), k.f(sd[e + "4.bias"])
    w.c3_w = k.f(sd[e + "6.weight"].permute(2, 3, 1, 0).reshape(9, 16, 64))
    w.c3_b, w.ln3_w, w.ln3_b = k.f(sd[e + "6.bias"]), k.f(sd[e + "7.weight"]), k.f(sd[e + "7.bias"])
    w.c3_wh = k.h(sd[e + "6.weight"].permute(0, 2, 3, 1).reshape(64, 9 * 16))  # [co][(ky*3+kx)*16+ci]
    w.c4_w = k.h(sd[e + "9.weight"].permute(0, 2, 3, 1).reshape(256, 9 * 64))  # [co][(ky*3+kx)*64+ci]
    w.c4_b, w.ln4_w, w.ln4_b = k.f(sd[e + "9.bias"]), k.f(sd[e + "10.weight"]), k.f(sd[e + "10.bias"])
    w.c5_w, w.c5_b = k.h(sd[e + "12.weight"].reshape(256, 256)), k.f(sd[e + "12.bias"])
    w.pix_w, w.pix_b = k.h(sd[prefix + "pix_feat_proj.weight"].reshape(256, 256)), k.f(sd[prefix + "pix_feat_proj.bias"])
    for i in range(2):
        p = f"{prefix}fuser.layers.{i}."
        cx = w.cx[i]
        cx.dw_w = k.f(sd[p + "dwconv.weight"].reshape(256, 49).t())           # [49][256]
        cx.dw_b, cx.ln_w, cx.ln_b = k.f(sd[p + "dwconv.bias"]), k.f(sd[p + "norm.weight"]), k.f(sd[p + "norm.bias"])
        cx.pw1_w, cx.pw1_b = k.h(sd[p + "pwconv1.weight"]), k.f(sd[p + "pwconv1.bias"])
        gamma = sd[p + "gamma"].float() if (p + "gamma") in sd else torch.ones(256)
        cx.pw2_w = k.h(gamma[:, None].to(sd[p + "pwconv2.weight"].device) * sd[p + "pwconv2.weight"].float())
        cx.pw2_b = k.f(gamma.to(sd[p + "pwconv2.bias"].device) * sd[p + "pwconv2.bias"].float())
    w.out_w, w.out_b = k.h(sd[prefix + "out_proj.weight"].reshape(64, 256)), k.f(sd[prefix + "out_proj.bias"])
    w.no_obj_embed = k.f(no_obj_embed.reshape(-1)) if no_obj_embed is not None else None
    return w, k
