"""SURVEY section 8 row f-4 on the GPU: forward_image (Hiera + FPN + conv_s0 / conv_s1, PyTorch-hosted) feeding the
CUDA hot path, FROM PIXELS, against the unmodified reference run on the same synthetic frames and seeded weights
(tests/golden/clip_pixels_t8.npz = BASELINE configs[0]: Hiera-T, 8 frames of 1024^2, one point prompt), and the
bf16 + CUDA-graph hosting of the encoder (GraphedImageEncoder) against the fp32 module."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
T = 8


def _full_state_dict(variant):
    from video_llava_seg_b200 import synth

    sd = dict(synth.init_state_dict(0))
    sd.update({"image_encoder." + k: v for k, v in synth.init_image_encoder_state_dict(variant, 0).items()})
    return sd


def _track(predictor, frames):
    state = predictor.init_state(frames)
    predictor.add_new_points_or_box(state, frame_idx=0, obj_id=1, points=[[300.0, 500.0]], labels=[1])
    out = {}
    for fi, ids, video_res in predictor.propagate_in_video(state):
        key = "cond_frame_outputs" if fi == 0 else "non_cond_frame_outputs"
        out[fi] = state["output_dict"][key][fi]
    return out


@pytest.fixture(scope="module")
def frames():
    from video_llava_seg_b200 import synth

    return synth.synthetic_frames(T, 1024, seed=1)


def test_propagation_from_pixels_matches_reference(vls_lib, frames):
    """fp32 encoder (TF32 off) + the bf16 CUDA hot path: low-res logits within 1e-2 of the reference on every frame.
    With random-init weights the mask logits from pixel features are small (|logit| < 0.5), so the binarised masks are
    compared on the pixels whose reference logit is not within the numerical tolerance of 0."""
    from video_llava_seg_b200 import build_sam

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gold = np.load(os.path.join(GOLD, "clip_pixels_t8.npz"))
    predictor = build_sam.build_sam2_video_predictor("t", _full_state_dict("t"), "cuda:0")
    out = _track(predictor, frames)
    assert sorted(out) == list(range(T))
    for t in range(T):
        pm = out[t]["pred_masks"].float().cpu()
        ref = torch.from_numpy(gold[f"mask_s2_{t}"])
        got = pm[:, :, ::2, ::2]
        keep = (got != 0.1) & (ref != 0.1)                     # hole filling is a discrete decision
        err = (got - ref).abs()[keep].max().item()
        ref_bits = np.unpackbits(gold[f"maskbits_{t}"], axis=1).reshape(1, 1, 256, 256).astype(bool)
        sure = ((ref.abs() > 2e-2) & keep).numpy()             # not within tolerance of 0, not a hole-filling decision
        agree = ((pm > 0).numpy()[:, :, ::2, ::2] == ref_bits[:, :, ::2, ::2])[sure].mean()
        ptr_err = (out[t]["obj_ptr"].cpu() - torch.from_numpy(gold[f"obj_ptr_{t}"])).abs().max().item()
        osl_err = (out[t]["object_score_logits"].cpu() - torch.from_numpy(gold[f"obj_score_{t}"])).abs().max().item()
        md = (out[t]["maskmem_features"].float().cpu()[:, :, ::4, ::4] - torch.from_numpy(gold[f"mem_s4_{t}"])).abs()
        print(f"pixels t={t}: logit err {err:.3e} agreement outside ties {agree:.5f} ({sure.mean():.3f} of the pixels) "
              f"ptr err {ptr_err:.3e} obj score err {osl_err:.3e} mem err mean {md.mean().item():.3e}")
        assert err < 1e-2, (t, err)
        assert agree >= 0.999, (t, agree)
        assert osl_err < 1e-2 and ptr_err < 5e-2, (t, osl_err, ptr_err)
        assert md.mean().item() < 2e-2


@pytest.mark.parametrize("variant", ["t", "b+"])
def test_graphed_bf16_encoder_tracks_the_fp32_module(vls_lib, frames, variant):
    """GraphedImageEncoder (bf16, one CUDA graph per input shape, conv_s0 / conv_s1 captured with the trunk): features
    within bf16 accuracy of the fp32 module, replays are deterministic, and a second frame reuses the captured graph."""
    from video_llava_seg_b200 import build_sam

    torch.backends.cudnn.allow_tf32 = False
    ref_p = build_sam.build_sam2_video_predictor(variant, _full_state_dict(variant), "cuda:0")
    fast_p = build_sam.build_sam2_video_predictor(variant, _full_state_dict(variant), "cuda:0",
                                                  image_encoder_dtype=torch.bfloat16)
    x = frames[:1].to("cuda:0")
    a = ref_p.forward_image(x)
    b = fast_p.forward_image(x)
    b = {k: ([t.float().clone() for t in v] if isinstance(v, list) else v.float().clone()) for k, v in b.items()}
    for lvl in range(3):
        fa, fb = a["backbone_fpn"][lvl], b["backbone_fpn"][lvl]
        assert fa.shape == fb.shape
        rel = (fa - fb).abs().max().item() / fa.abs().max().item()
        mean_rel = (fa - fb).abs().mean().item() / fa.abs().mean().item()
        print(f"{variant} level {lvl}: bf16-vs-fp32 max rel {rel:.3e} mean rel {mean_rel:.3e}")
        assert mean_rel < 2e-2 and rel < 0.15
        assert (a["vision_pos_enc"][lvl] - b["vision_pos_enc"][lvl]).abs().max().item() < 1e-2
    n_graphs = len(fast_p.image_encoder._graphs)
    c = fast_p.forward_image(frames[1:2].to("cuda:0"))
    assert len(fast_p.image_encoder._graphs) == n_graphs == 1
    assert not torch.equal(c["backbone_fpn"][2].float(), b["backbone_fpn"][2])    # new frame, new features
    d = fast_p.forward_image(x)
    assert torch.equal(d["backbone_fpn"][2].float(), b["backbone_fpn"][2])        # replay is deterministic


def test_from_pixels_with_bf16_encoder_end_to_end(vls_lib, frames):
    """The fast configuration end to end (bf16 graphed encoder + CUDA hot path): masks agree with the reference outside
    the near-zero band that bf16 feature noise can flip."""
    from video_llava_seg_b200 import build_sam

    gold = np.load(os.path.join(GOLD, "clip_pixels_t8.npz"))
    predictor = build_sam.build_sam2_video_predictor("t", _full_state_dict("t"), "cuda:0", image_encoder_dtype=torch.bfloat16)
    out = _track(predictor, frames)
    for t in range(T):
        pm = out[t]["pred_masks"].float().cpu()
        ref = torch.from_numpy(gold[f"mask_s2_{t}"])
        err = (pm[:, :, ::2, ::2] - ref).abs()
        sure = (ref.abs() > 0.1) & (ref != 0.1) & (pm[:, :, ::2, ::2] != 0.1)
        agree = ((pm[:, :, ::2, ::2] > 0) == (ref > 0))[sure].float().mean().item()
        print(f"bf16 encoder t={t}: logit err max {err.max().item():.3e} mean {err.mean().item():.3e} agreement {agree:.5f}")
        assert err.mean().item() < 2e-2 and agree >= 0.995
