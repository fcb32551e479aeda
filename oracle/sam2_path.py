"""TEST INFRASTRUCTURE ONLY -- CPU restatement (fp32, torch CPU operators) of the reference's
SAM 2.1 per-frame mask-propagation path.  Only tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py may import this file; the product path
(video-llava-seg_b200) never does.

Pinning: this restatement is checked (tests/test_oracle_vs_reference.py, run in the build
container where /root/reference exists) against the *unmodified* reference modules on the same
seeded weights and inputs, and against the golden vectors in tests/golden/ which were produced
by the reference itself (tests/golden/make_golden.py).  The reference has no tests or golden
vectors of its own, so that is the strongest pin available.

All functions are written against a flat ``sd`` dict whose keys are exactly the reference
``state_dict`` keys (build_sam.py:141-151 loads strictly), so a reference model's weights drop in.
Each function cites the reference lines it restates (paths relative to /root/reference).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

NO_OBJ_SCORE = -1024.0  # sam2/modeling/sam2_base.py:19


# --------------------------------------------------------------------------- small pieces
def linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def layer_norm(sd, p, x, eps=1e-5):
    """nn.LayerNorm default eps (memory_attention.py:43-45, sam/transformer.py:162-174)."""
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def layer_norm2d(sd, p, x, eps=1e-6):
    """Channel LayerNorm on NCHW, biased variance (sam2_utils.py:141-153)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[p + ".weight"][None, :, None, None] * x + sd[p + ".bias"][None, :, None, None]


def mlp(sd, p, x, num_layers, sigmoid_output=False):
    """ReLU MLP (sam2_utils.py:112-136)."""
    for i in range(num_layers):
        x = linear(sd, f"{p}.layers.{i}", x)
        if i < num_layers - 1:
            x = F.relu(x)
    return torch.sigmoid(x) if sigmoid_output else x


def sine_pe_1d(pos, dim, temperature=10000.0):
    """sam2_utils.py:64-74: sin half then cos half."""
    pe_dim = dim // 2
    dim_t = torch.arange(pe_dim, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / pe_dim)
    e = pos.unsqueeze(-1) / dim_t
    return torch.cat([e.sin(), e.cos()], dim=-1)


def sine_pe_2d(num_pos_feats, h, w, temperature=10000.0):
    """PositionEmbeddingSine.forward (position_encoding.py:78-112) for one image -> [C,h,w]."""
    half = num_pos_feats // 2
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32).view(h, 1).expand(h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32).view(1, w).expand(h, w)
    y = y / (y[-1:, :] + eps) * scale
    x = x / (x[:, -1:] + eps) * scale
    dim_t = torch.arange(half, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / half)
    px = x[:, :, None] / dim_t
    py = y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).permute(2, 0, 1).contiguous()


def axial_rope_table(end_x, end_y, dim=256, theta=10000.0):
    """compute_axial_cis (position_encoding.py:168-184) as (cos, sin) tables [end_x*end_y, dim/2].

    Pair j<dim/4 rotates by x*theta^(-4j/dim); pair j>=dim/4 by y*theta^(-4(j-dim/4)/dim).
    """
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    t = torch.arange(end_x * end_y, dtype=torch.float32)
    tx = (t % end_x).float()
    ty = torch.div(t, end_x, rounding_mode="floor").float()
    ang = torch.cat([torch.outer(tx, freqs), torch.outer(ty, freqs)], dim=-1)
    return torch.cos(ang), torch.sin(ang)


def apply_rope(x, cos, sin):
    """apply_rotary_enc (position_encoding.py:195-222): complex multiply on adjacent pairs, fp32.

    x: [..., N, D]; cos/sin: [N, D/2] (already repeated for keys, position_encoding.py:213-220).
    """
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    xe, xo = xf[..., 0], xf[..., 1]
    oe = xe * cos - xo * sin
    oo = xe * sin + xo * cos
    return torch.stack((oe, oo), dim=-1).flatten(-2).type_as(x)


def sdpa(q, k, v):
    """F.scaled_dot_product_attention as called at sam/transformer.py:270,344 (scale 1/sqrt(d))."""
    return F.scaled_dot_product_attention(q, k, v)


def _heads(x, n):
    b, t, c = x.shape
    return x.reshape(b, t, n, c // n).transpose(1, 2)


def _unheads(x):
    b, n, t, c = x.shape
    return x.transpose(1, 2).reshape(b, t, n * c)


# --------------------------------------------------------------------------- attention flavours
def attention(sd, p, q, k, v, num_heads):
    """Attention.forward (sam/transformer.py:255-286)."""
    q = _heads(linear(sd, p + ".q_proj", q), num_heads)
    k = _heads(linear(sd, p + ".k_proj", k), num_heads)
    v = _heads(linear(sd, p + ".v_proj", v), num_heads)
    return linear(sd, p + ".out_proj", _unheads(sdpa(q, k, v)))


def rope_attention(sd, p, q, k, v, num_k_exclude_rope=0, rope_k_repeat=False, num_heads=1, theta=10000.0):
    """RoPEAttention.forward (sam/transformer.py:311-360). The table is recomputed for
    sqrt(Nq) x sqrt(Nq) at first call (:325-328), so feat_sizes in the YAML is irrelevant."""
    q = _heads(linear(sd, p + ".q_proj", q), num_heads)
    k = _heads(linear(sd, p + ".k_proj", k), num_heads)
    v = _heads(linear(sd, p + ".v_proj", v), num_heads)
    nq = q.shape[-2]
    side = int(round(math.sqrt(nq)))
    cos, sin = axial_rope_table(side, side, dim=q.shape[-1], theta=theta)
    num_k_rope = k.shape[-2] - num_k_exclude_rope
    q = apply_rope(q, cos, sin)
    if num_k_rope > 0:
        if num_k_rope != nq:
            assert rope_k_repeat and num_k_rope % nq == 0
        r = num_k_rope // nq
        k = torch.cat([apply_rope(k[:, :, :num_k_rope], cos.repeat(r, 1), sin.repeat(r, 1)),
                       k[:, :, num_k_rope:]], dim=2)
    return linear(sd, p + ".out_proj", _unheads(sdpa(q, k, v)))


# --------------------------------------------------------------------------- memory attention
def memory_attention(sd, curr, memory, curr_pos, memory_pos, num_obj_ptr_tokens=0,
                     p="memory_attention", num_layers=4):
    """MemoryAttention.forward (memory_attention.py:119-169) + MemoryAttentionLayer (:58-99).

    curr, curr_pos: [Nq,B,256]; memory, memory_pos: [Nk,B,64]; returns [Nq,B,256].
    pos_enc_at_input=True (0.1*pos, :141); pos_enc_at_attn=False; pos at cross-attn keys only
    (sam2.1_hiera_b+.yaml:35,45,46); FFN ReLU; dropout inert in eval.
    """
    x = (curr + 0.1 * curr_pos).transpose(0, 1)
    mem = memory.transpose(0, 1)
    mem_pos = memory_pos.transpose(0, 1)
    for i in range(num_layers):
        lp = f"{p}.layers.{i}"
        t = layer_norm(sd, lp + ".norm1", x)
        x = x + rope_attention(sd, lp + ".self_attn", t, t, t)
        t = layer_norm(sd, lp + ".norm2", x)
        x = x + rope_attention(sd, lp + ".cross_attn_image", t, mem + mem_pos, mem,
                               num_k_exclude_rope=num_obj_ptr_tokens, rope_k_repeat=True)
        t = layer_norm(sd, lp + ".norm3", x)
        x = x + linear(sd, lp + ".linear2", F.relu(linear(sd, lp + ".linear1", t)))
    return layer_norm(sd, p + ".norm", x).transpose(0, 1)


# --------------------------------------------------------------------------- mask decoder
def two_way_transformer(sd, p, image_embedding, image_pe, point_embedding, depth=2, num_heads=8):
    """TwoWayTransformer.forward / TwoWayAttentionBlock.forward (sam/transformer.py:90-134,181-212)."""
    keys = image_embedding.flatten(2).permute(0, 2, 1)
    key_pe = image_pe.flatten(2).permute(0, 2, 1)
    queries, query_pe = point_embedding, point_embedding
    for i in range(depth):
        lp = f"{p}.layers.{i}"
        if i == 0:  # skip_first_layer_pe
            queries = attention(sd, lp + ".self_attn", queries, queries, queries, num_heads)
        else:
            q = queries + query_pe
            queries = queries + attention(sd, lp + ".self_attn", q, q, queries, num_heads)
        queries = layer_norm(sd, lp + ".norm1", queries)
        q, k = queries + query_pe, keys + key_pe
        queries = layer_norm(sd, lp + ".norm2",
                             queries + attention(sd, lp + ".cross_attn_token_to_image", q, k, keys, num_heads))
        queries = layer_norm(sd, lp + ".norm3", queries + mlp(sd, lp + ".mlp", queries, 2))
        q, k = queries + query_pe, keys + key_pe
        keys = layer_norm(sd, lp + ".norm4",
                          keys + attention(sd, lp + ".cross_attn_image_to_token", k, q, queries, num_heads))
    q, k = queries + query_pe, keys + key_pe
    queries = queries + attention(sd, p + ".final_attn_token_to_image", q, k, keys, num_heads)
    return layer_norm(sd, p + ".norm_final_attn", queries), keys


def mask_decoder(sd, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                 multimask_output, repeat_image, high_res_features=None, p="sam_mask_decoder",
                 num_mask_tokens=4, iou_sigmoid=True, use_multimask_token_for_obj_ptr=True):
    """MaskDecoder.forward / predict_masks (sam/mask_decoder.py:110-245) with pred_obj_scores(_mlp),
    use_high_res_features and iou_prediction_use_sigmoid as in sam2.1 YAMLs (:98,:102-103,:111).
    The stability fallback is commented out in this fork (:149-150)."""
    out_tok = torch.cat([sd[p + ".obj_score_token.weight"], sd[p + ".iou_token.weight"],
                         sd[p + ".mask_tokens.weight"]], dim=0)
    bsz = sparse_prompt_embeddings.size(0)
    tokens = torch.cat((out_tok.unsqueeze(0).expand(bsz, -1, -1), sparse_prompt_embeddings), dim=1)
    src = torch.repeat_interleave(image_embeddings, bsz, dim=0) if repeat_image else image_embeddings
    assert src.shape[0] == bsz
    src = src + dense_prompt_embeddings
    pos_src = torch.repeat_interleave(image_pe, bsz, dim=0)
    b, c, h, w = src.shape
    hs, src = two_way_transformer(sd, p + ".transformer", src, pos_src, tokens)
    iou_token_out = hs[:, 1, :]
    mask_tokens_out = hs[:, 2:2 + num_mask_tokens, :]
    src = src.transpose(1, 2).reshape(b, c, h, w)
    up = p + ".output_upscaling"
    feat_s0, feat_s1 = high_res_features
    x = F.conv_transpose2d(src, sd[up + ".0.weight"], sd[up + ".0.bias"], stride=2) + feat_s1
    x = F.gelu(layer_norm2d(sd, up + ".1", x))
    x = F.gelu(F.conv_transpose2d(x, sd[up + ".3.weight"], sd[up + ".3.bias"], stride=2) + feat_s0)
    hyper = torch.stack([mlp(sd, f"{p}.output_hypernetworks_mlps.{i}", mask_tokens_out[:, i, :], 3)
                         for i in range(num_mask_tokens)], dim=1)
    b, c, h, w = x.shape
    masks = (hyper @ x.view(b, c, h * w)).view(b, -1, h, w)
    iou_pred = mlp(sd, p + ".iou_prediction_head", iou_token_out, 3, sigmoid_output=iou_sigmoid)
    obj_logits = mlp(sd, p + ".pred_obj_score_head", hs[:, 0, :], 3)
    if multimask_output:
        masks, iou_pred = masks[:, 1:], iou_pred[:, 1:]
    else:
        masks, iou_pred = masks[:, 0:1], iou_pred[:, 0:1]
    if multimask_output and use_multimask_token_for_obj_ptr:
        tok = mask_tokens_out[:, 1:]
    else:
        tok = mask_tokens_out[:, 0:1]
    return masks, iou_pred, tok, obj_logits


# --------------------------------------------------------------------------- prompt encoder (adjacent, tiny)
def dense_pe(sd, size=64, p="sam_prompt_encoder"):
    """PromptEncoder.get_dense_pe / PositionEmbeddingRandom.forward (position_encoding.py:138-150)."""
    g = sd[p + ".pe_layer.positional_encoding_gaussian_matrix"]
    grid = torch.ones((size, size), dtype=torch.float32)
    y = (grid.cumsum(0) - 0.5) / size
    x = (grid.cumsum(1) - 0.5) / size
    c = 2 * torch.stack([x, y], dim=-1) - 1
    c = 2 * math.pi * (c @ g)
    return torch.cat([c.sin(), c.cos()], dim=-1).permute(2, 0, 1).unsqueeze(0)


def prompt_points(sd, coords, labels, image_size=1024, p="sam_prompt_encoder"):
    """PromptEncoder._embed_points with pad=True (prompt_encoder.py:79-103) -> sparse [B,P+1,256]."""
    pts = coords + 0.5
    pts = torch.cat([pts, torch.zeros(pts.shape[0], 1, 2)], dim=1)
    lab = torch.cat([labels, -torch.ones(labels.shape[0], 1, dtype=labels.dtype)], dim=1)
    g = sd[p + ".pe_layer.positional_encoding_gaussian_matrix"]
    c = 2 * (pts / image_size).float() - 1
    c = 2 * math.pi * (c @ g)
    e = torch.cat([c.sin(), c.cos()], dim=-1)
    e[lab == -1] = 0.0
    e[lab == -1] += sd[p + ".not_a_point_embed.weight"]
    for i in range(4):
        e[lab == i] += sd[f"{p}.point_embeddings.{i}.weight"]
    return e


def dense_no_mask(sd, bsz, size=64, p="sam_prompt_encoder"):
    """prompt_encoder.py:178-180."""
    return sd[p + ".no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(bsz, -1, size, size)


# --------------------------------------------------------------------------- memory encoder
def memory_encoder(sd, pix_feat, masks, skip_mask_sigmoid=False, p="memory_encoder"):
    """MemoryEncoder.forward (memory_encoder.py:158-181); MaskDownSampler k3 s2 p1 x4 + 1x1 (:17-58,
    YAML :68-72); CXBlock x2 (:62-117); out_proj 256->64; sine PE 64-d (YAML :62-67)."""
    if not skip_mask_sigmoid:
        masks = torch.sigmoid(masks)
    x = masks
    e = p + ".mask_downsampler.encoder"
    for i in range(4):
        x = F.conv2d(x, sd[f"{e}.{3 * i}.weight"], sd[f"{e}.{3 * i}.bias"], stride=2, padding=1)
        x = F.gelu(layer_norm2d(sd, f"{e}.{3 * i + 1}", x))
    x = F.conv2d(x, sd[e + ".12.weight"], sd[e + ".12.bias"])
    x = F.conv2d(pix_feat, sd[p + ".pix_feat_proj.weight"], sd[p + ".pix_feat_proj.bias"]) + x
    for i in range(2):
        lp = f"{p}.fuser.layers.{i}"
        y = F.conv2d(x, sd[lp + ".dwconv.weight"], sd[lp + ".dwconv.bias"], padding=3, groups=x.shape[1])
        y = layer_norm2d(sd, lp + ".norm", y).permute(0, 2, 3, 1)
        y = linear(sd, lp + ".pwconv2", F.gelu(linear(sd, lp + ".pwconv1", y)))
        y = (sd[lp + ".gamma"] * y).permute(0, 3, 1, 2)
        x = x + y
    x = F.conv2d(x, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])
    pos = sine_pe_2d(x.shape[1], x.shape[2], x.shape[3])[None].repeat(x.shape[0], 1, 1, 1)
    return {"vision_features": x, "vision_pos_enc": [pos.to(x.dtype)]}


# --------------------------------------------------------------------------- connected components
def cc_label(mask):
    """Closed form of sam2/csrc/connected_components.cu:62-209 (8-connectivity, block-based
    union-find with atomicMin): label = 1 + min over the component of ((r&~1)*W + (c&~1));
    count = component area.  mask: uint8/bool [N,1,H,W] -> (labels i32, counts i32).
    The plain-C restatement in oracle/cc_oracle.c is the primary oracle; this numpy/scipy one
    cross-checks it."""
    import numpy as np
    from scipy import ndimage

    m = mask.detach().cpu().numpy().astype(bool)
    n, _, h, w = m.shape
    rr, cc = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    blk = ((rr & ~1) * w + (cc & ~1)).astype(np.int64)
    labels = np.zeros(m.shape, np.int32)
    counts = np.zeros(m.shape, np.int32)
    for i in range(n):
        lab, k = ndimage.label(m[i, 0], structure=np.ones((3, 3)))
        if k == 0:
            continue
        idx = np.arange(1, k + 1)
        mn = ndimage.minimum(blk, lab, idx).astype(np.int64)
        ar = ndimage.sum(m[i, 0], lab, idx).astype(np.int64)
        lut_l = np.concatenate([[0], mn + 1]).astype(np.int32)
        lut_c = np.concatenate([[0], ar]).astype(np.int32)
        labels[i, 0] = lut_l[lab]
        counts[i, 0] = lut_c[lab]
    return torch.from_numpy(labels), torch.from_numpy(counts)


def fill_holes_in_mask_scores(mask, max_area, cc=cc_label):
    """sam2/utils/misc.py:312-338 (without the swallow-all except)."""
    labels, areas = cc(mask <= 0)
    is_hole = (labels > 0) & (areas <= max_area)
    return torch.where(is_hole, 0.1, mask)


# --------------------------------------------------------------------------- tracking core
class Cfg:
    """Hot-path hyper-parameters shared by sam2.1_hiera_{t,s,b+,l}.yaml:26-116 + build_sam.py:93-102."""
    image_size = 1024
    feat = 64
    hidden = 256
    mem_dim = 64
    num_maskmem = 7
    max_obj_ptrs = 16
    sigmoid_scale = 20.0
    sigmoid_bias = -10.0
    fill_hole_area = 8


def forward_sam_heads(sd, cfg, pix_feat, high_res, point_inputs, multimask_output):
    """SAM2Base._forward_sam_heads (sam2_base.py:257-413), mask_inputs=None path."""
    bsz = pix_feat.size(0)
    if point_inputs is not None and "prompt_embedding" in point_inputs:
        # [SEG]-token prompt: what the reference predictor computes when
        # sam_prompt_encoder.forward is patched to return this sparse embedding (+ the no-mask dense embedding)
        sparse = point_inputs["prompt_embedding"]
    else:
        if point_inputs is not None:
            coords, labels = point_inputs["point_coords"], point_inputs["point_labels"]
        else:
            coords = torch.zeros(bsz, 1, 2)
            labels = -torch.ones(bsz, 1, dtype=torch.int32)
        sparse = prompt_points(sd, coords, labels, cfg.image_size)
    dense = dense_no_mask(sd, bsz, cfg.feat)
    low, ious, tokens, obj_logits = mask_decoder(sd, pix_feat, dense_pe(sd, cfg.feat), sparse, dense,
                                                 multimask_output, False, high_res)
    is_obj = obj_logits > 0
    low = torch.where(is_obj[:, None, None], low, torch.tensor(NO_OBJ_SCORE)).float()
    high = F.interpolate(low, size=(cfg.image_size, cfg.image_size), mode="bilinear", align_corners=False)
    tok = tokens[:, 0]
    if multimask_output:
        best = torch.argmax(ious, dim=-1)
        bi = torch.arange(bsz)
        low_best, high_best = low[bi, best].unsqueeze(1), high[bi, best].unsqueeze(1)
        if tokens.size(1) > 1:
            tok = tokens[bi, best]
    else:
        low_best, high_best = low, high
    ptr = mlp(sd, "obj_ptr_proj", tok, 3)
    lam = is_obj.float()
    ptr = lam * ptr + (1 - lam) * sd["no_obj_ptr"]
    return dict(low_res_multimasks=low, ious=ious, low_res_masks=low_best, high_res_masks=high_best,
                obj_ptr=ptr, object_score_logits=obj_logits)


def encode_new_memory(sd, cfg, vision_feat, high_res_masks, obj_logits, is_mask_from_pts):
    """SAM2Base._encode_new_memory (sam2_base.py:676-724), eval mode, non_overlap off."""
    bsz = vision_feat.size(1)
    pix = vision_feat.permute(1, 2, 0).reshape(bsz, cfg.hidden, cfg.feat, cfg.feat)
    if is_mask_from_pts:  # binarize_mask_from_pts_for_mem_enc (build_sam.py:99)
        m = (high_res_masks > 0).float()
    else:
        m = torch.sigmoid(high_res_masks)
    m = m * cfg.sigmoid_scale + cfg.sigmoid_bias
    out = memory_encoder(sd, pix, m, skip_mask_sigmoid=True)
    feats = out["vision_features"]
    is_obj = (obj_logits > 0).float()
    feats = feats + (1 - is_obj[..., None, None]) * sd["no_obj_embed_spatial"][..., None, None]
    return feats, out["vision_pos_enc"]


def memory_bank(sd, cfg, frame_idx, output_dict, num_frames):
    """Memory / pointer assembly of SAM2Base._prepare_memory_conditioned_features
    (sam2_base.py:522-663) for forward tracking, eval, stride 1, all cond frames selected."""
    mem, pos = [], []
    cond = output_dict["cond_frame_outputs"]
    items = [(0, o) for o in cond.values()]
    for t_pos in range(1, cfg.num_maskmem):
        t_rel = cfg.num_maskmem - t_pos
        items.append((t_pos, output_dict["non_cond_frame_outputs"].get(frame_idx - t_rel)))
    for t_pos, prev in items:
        if prev is None:
            continue
        mem.append(prev["maskmem_features"].float().flatten(2).permute(2, 0, 1))
        e = prev["maskmem_pos_enc"][-1].flatten(2).permute(2, 0, 1)
        pos.append(e + sd["maskmem_tpos_enc"][cfg.num_maskmem - t_pos - 1])
    bsz = mem[0].shape[1]
    max_ptrs = min(num_frames, cfg.max_obj_ptrs)
    pp = [(frame_idx - t, o["obj_ptr"]) for t, o in cond.items() if t <= frame_idx]
    for t_diff in range(1, max_ptrs):
        t = frame_idx - t_diff
        if t < 0:
            break
        o = output_dict["non_cond_frame_outputs"].get(t)
        if o is not None:
            pp.append((t_diff, o["obj_ptr"]))
    n_ptr_tokens = 0
    if pp:
        pl, ptrs = zip(*pp)
        ptrs = torch.stack(ptrs, dim=0)
        op = sine_pe_1d(torch.tensor(pl).float() / (max_ptrs - 1), cfg.hidden)
        op = linear(sd, "obj_ptr_tpos_proj", op).unsqueeze(1).expand(-1, bsz, cfg.mem_dim)
        k = cfg.hidden // cfg.mem_dim
        ptrs = ptrs.reshape(-1, bsz, k, cfg.mem_dim).permute(0, 2, 1, 3).flatten(0, 1)
        op = op.repeat_interleave(k, dim=0)
        mem.append(ptrs)
        pos.append(op)
        n_ptr_tokens = ptrs.shape[0]
    return torch.cat(mem, dim=0), torch.cat(pos, dim=0), n_ptr_tokens


def track_step(sd, cfg, frame_idx, is_init_cond_frame, feats, point_inputs, output_dict, num_frames,
               run_mem_encoder=True):
    """SAM2Base.track_step (sam2_base.py:726-877) for point / no prompts.

    feats: dict(vision_feat [4096,B,256], vision_pos [4096,B,256], feat_s0 [B,32,256,256],
    feat_s1 [B,64,128,128]) -- the output of forward_image + _prepare_backbone_features.
    """
    vf, vp = feats["vision_feat"], feats["vision_pos"]
    bsz = vf.size(1)
    if is_init_cond_frame:  # directly_add_no_mem_embed (sam2_base.py:651-655)
        pix = (vf + sd["no_mem_embed"]).permute(1, 2, 0).reshape(bsz, cfg.hidden, cfg.feat, cfg.feat)
    else:
        mem, mem_pos, n_ptr = memory_bank(sd, cfg, frame_idx, output_dict, num_frames)
        pix = memory_attention(sd, vf, mem, vp, mem_pos, n_ptr)
        pix = pix.permute(1, 2, 0).reshape(bsz, cfg.hidden, cfg.feat, cfg.feat)
    if point_inputs is not None and "prompt_embedding" in point_inputs:
        multimask = False  # single-mask decode, as the LLaVA head does (llava/model/seg_head/sam2.py:111)
    else:
        num_pts = 0 if point_inputs is None else point_inputs["point_labels"].size(1)
        multimask = 0 <= num_pts <= 1  # _use_multimask (sam2_base.py:879-887) with YAML :109-113
    out = forward_sam_heads(sd, cfg, pix, [feats["feat_s0"], feats["feat_s1"]], point_inputs, multimask)
    cur = dict(pred_masks=out["low_res_masks"], pred_masks_high_res=out["high_res_masks"],
               obj_ptr=out["obj_ptr"], object_score_logits=out["object_score_logits"], ious=out["ious"],
               pix_feat_with_mem=pix, maskmem_features=None, maskmem_pos_enc=None)
    if run_mem_encoder:
        f, pe = encode_new_memory(sd, cfg, vf, out["high_res_masks"], out["object_score_logits"],
                                  point_inputs is not None)
        cur["maskmem_features"], cur["maskmem_pos_enc"] = f, pe
    return cur


def propagate(sd, cfg, frame_feats, point_inputs_frame0, num_frames=None, fill_holes=True, cc=cc_label,
              on_frame=None):
    """init_state + add_new_points_or_box on frame 0 + propagate_in_video
    (sam2_video_predictor.py:173-314, :593-745, :912-978) for B objects prompted on frame 0
    with the same number of clicks each.  `frame_feats(t)` returns the feats dict for frame t
    (already expanded to B).  Returns the per-frame compact outputs (bf16 memory, hole-filled
    low-res masks), i.e. what the reference keeps in inference_state["output_dict"].
    """
    num_frames = num_frames or len(frame_feats)
    get = frame_feats if callable(frame_feats) else (lambda t: frame_feats[t])
    output_dict = {"cond_frame_outputs": OrderedDict(), "non_cond_frame_outputs": OrderedDict()}
    results = []
    # prompt frame: decoder without memory; memory encoder runs in the preflight on the
    # (consolidated) low-res mask re-upsampled to 1024 (sam2_video_predictor.py:533-550)
    f0 = get(0)
    bsz = f0["vision_feat"].size(1)
    outs = []
    for b in range(bsz):  # add_new_points_or_box runs per object with batch_size=1 (:283-298)
        fb = {k: v[:, b:b + 1] if k.startswith("vision") else v[b:b + 1] for k, v in f0.items()}
        pb = {k: v[b:b + 1] for k, v in point_inputs_frame0.items()}
        o = track_step(sd, cfg, 0, True, fb, pb, {}, num_frames, run_mem_encoder=False)
        pm = o["pred_masks"]
        if fill_holes:
            pm = fill_holes_in_mask_scores(pm, cfg.fill_hole_area, cc)
        o["pred_masks"] = pm
        outs.append(o)
    pred = torch.cat([o["pred_masks"] for o in outs], 0)
    ptr = torch.cat([o["obj_ptr"] for o in outs], 0)
    osl = torch.cat([o["object_score_logits"] for o in outs], 0)
    high = F.interpolate(pred, size=(cfg.image_size, cfg.image_size), mode="bilinear", align_corners=False)
    mf, mpe = encode_new_memory(sd, cfg, f0["vision_feat"], high, osl, True)
    cond = dict(maskmem_features=mf.to(torch.bfloat16), maskmem_pos_enc=mpe, pred_masks=pred, obj_ptr=ptr,
                object_score_logits=osl)
    output_dict["cond_frame_outputs"][0] = cond
    results.append(cond)
    if on_frame:
        on_frame(0, cond)
    for t in range(1, num_frames):
        o = track_step(sd, cfg, t, False, get(t), None, output_dict, num_frames, run_mem_encoder=True)
        pm = o["pred_masks"]
        if fill_holes:
            pm = fill_holes_in_mask_scores(pm, cfg.fill_hole_area, cc)
        compact = dict(maskmem_features=o["maskmem_features"].to(torch.bfloat16),
                       maskmem_pos_enc=o["maskmem_pos_enc"], pred_masks=pm, obj_ptr=o["obj_ptr"],
                       object_score_logits=o["object_score_logits"], ious=o["ious"])
        output_dict["non_cond_frame_outputs"][t] = compact
        results.append(compact)
        if on_frame:
            on_frame(t, compact)
    return results
