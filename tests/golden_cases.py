"""Seeded inputs shared by tests/golden/make_golden.py (reference side, build container) and the
parity tests (oracle / CUDA side).  The draw ORDER is part of the fixture contract."""
import torch


def module_inputs():
    g = torch.Generator().manual_seed(1234)
    rn = lambda *s: torch.randn(*s, generator=g)
    d = {}
    nq, b = 256, 2          # 16x16 query grid, 2 memory frames + 2 pointers (8 tokens)
    nk = 2 * nq + 8
    d["curr"], d["curr_pos"] = rn(nq, b, 256) * 0.5, rn(nq, b, 256) * 0.5
    d["mem"], d["mem_pos"] = rn(nk, b, 64) * 0.5, rn(nk, b, 64) * 0.5
    d["n_ptr_tokens"] = 8
    d["emb"] = rn(2, 256, 64, 64) * 0.5
    d["s0"], d["s1"] = rn(2, 32, 256, 256) * 0.3, rn(2, 64, 128, 128) * 0.3
    d["sparse"] = rn(2, 2, 256)
    d["seg"] = rn(3, 1, 256)
    d["pix"] = rn(2, 256, 64, 64) * 0.5
    d["msk"] = torch.sigmoid(rn(2, 1, 1024, 1024) * 3) * 20 - 10
    return d


# ----------------------------------------------------------------------------- predictor API scenarios
def blob_mask(cx=300.0, cy=500.0, radius=90.0, size=1024):
    """Boolean [size,size] disk: the mask prompt of the add_new_mask scenario."""
    yy = torch.arange(size, dtype=torch.float32).view(size, 1) + 0.5
    xx = torch.arange(size, dtype=torch.float32).view(1, size) + 0.5
    return ((xx - cx) ** 2 + (yy - cy) ** 2) <= radius * radius


API_FRAMES = 4
API_CLIP_SEED = 8


def api_scenarios(predictor, init_state, names=None):
    """Drives a predictor with the reference's public API (sam2_video_predictor.py) through the calls r1 left untested --
    add_new_mask (:321), re-click on a tracked frame with prev_sam_mask_logits (:269-297), clear_all_prompts_in_frame
    (:777), non_overlap_masks (_apply_non_overlapping_constraints, sam2_base.py:889), reverse propagation,
    offload_state_to_cpu (:70-75) -- and records what they return.  The SAME function runs the unmodified reference
    (tests/golden/make_golden.py -> tests/golden/api.npz) and this repository's predictor (tests/test_predictor_gpu.py).
    `init_state(**kw)` opens a session on the API_CLIP_SEED clip with API_FRAMES frames."""
    import numpy as np

    rec = {}

    def bits(name, video_res):
        rec[name] = np.packbits((video_res.detach().float().cpu() > 0).numpy().reshape(video_res.shape[0], -1), axis=1)

    def track(tag, st, ambiguous=False, **kw):
        for f, ids, video in predictor.propagate_in_video(st, **kw):
            bits(f"{tag}_f{f}", video)
            key = "cond_frame_outputs" if f in st["output_dict"]["cond_frame_outputs"] else "non_cond_frame_outputs"
            o = st["output_dict"][key][f]
            rec[f"{tag}_ptr{f}"] = o["obj_ptr"].detach().float().cpu().numpy()
            if ambiguous:
                # the non-overlap constraint is a per-pixel ARG-MAX over the objects: where the two objects' scores are
                # within the numerical tolerance of each other the winner is a coin flip -- mark those pixels
                pm = torch.nn.functional.interpolate(o["pred_masks"].detach().float().cpu(), size=tuple(video.shape[-2:]),
                                                     mode="bilinear", align_corners=False)
                mx = pm.max(0).values        # ties only matter where an object is (nearly) foreground
                amb = (((pm[0] - pm[1]).abs() < 3e-2) & (mx > -3e-2)) | (mx.abs() < 1e-2)
                rec[f"{tag}_amb{f}"] = np.packbits(amb.numpy().reshape(1, -1), axis=1)

    want = lambda n: names is None or n in names
    if want("mask"):
        st = init_state()
        f, ids, video = predictor.add_new_mask(st, 0, 1, blob_mask())
        assert f == 0 and list(ids) == [1]
        bits("mask_prompt", video)
        track("mask", st)
    if want("reclick"):
        st = init_state()
        predictor.add_new_points_or_box(st, 0, 1, points=[[300.0, 500.0]], labels=[1])
        track("reclick_first", st)
        # a negative click on an already tracked frame: the decoder also gets the frame's previous logits
        f, ids, video = predictor.add_new_points_or_box(st, 2, 1, points=[[400.0, 520.0]], labels=[0])
        assert f == 2
        bits("reclick_edit", video)
        track("reclick_second", st)
        f, ids, video = predictor.clear_all_prompts_in_frame(st, 2, 1)
        bits("reclick_cleared", video)
        track("reclick_third", st)
    if want("nonoverlap"):
        old = predictor.non_overlap_masks
        predictor.non_overlap_masks = True
        try:
            # two objects prompted with OVERLAPPING mask prompts: with random-init weights two click prompts give nearly the
            # same mask for both objects and the per-pixel arg-max would be a coin flip everywhere
            st = init_state()
            predictor.add_new_mask(st, 0, 1, blob_mask(300.0, 500.0, 90.0))
            f, ids, video = predictor.add_new_mask(st, 0, 2, blob_mask(400.0, 540.0, 110.0))
            bits("nonoverlap_prompt", video)
            track("nonoverlap", st, ambiguous=True)
        finally:
            predictor.non_overlap_masks = old
    if want("reverse"):
        st = init_state()
        predictor.add_new_points_or_box(st, API_FRAMES - 1, 1, points=[[360.0, 530.0]], labels=[1])
        track("reverse", st, reverse=True)
    if want("offload"):
        st = init_state(offload_state_to_cpu=True)
        predictor.add_new_points_or_box(st, 0, 1, box=[200.0, 400.0, 420.0, 620.0])
        track("offload", st)
    return rec


# ----------------------------------------------------------------------------- LLaVA seg head (BASELINE configs[3])
SEG_META = {"padding": [0, 0, 0, 448], "resized_image_size": [576, 1024], "orig_image_size": [480, 854]}


def seg_head_inputs():
    """Seeded inputs of the SegmentationHeadSAM2.forward parity case: 3 frames of backbone features, 2 objects x 2 seg
    queries, 512-d `[SEG]` hidden states, the head's own projection weights."""
    g = torch.Generator().manual_seed(4321)
    rn = lambda *s: torch.randn(*s, generator=g)
    return {"feats": rn(3, 256, 64, 64) * 0.5, "s0": rn(3, 32, 256, 256) * 0.3, "s1": rn(3, 64, 128, 128) * 0.3,
            "tokens": rn(2, 512), "proj_w": rn(512, 512) * 0.05, "proj_b": rn(512) * 0.05, "n_seg_queries": 2}
