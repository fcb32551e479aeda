"""Dev probe: whole-clip cost (64 frames: prompt frame + 16 ramp frames with a growing bank + graph capture + steady state)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import build_sam, synth
from video_llava_seg_b200.features import FeatureClip

dev = torch.device("cuda:0")
predictor = build_sam.build_sam2_video_predictor(None, synth.init_state_dict(0), dev)
T = 64
clip = synth.SyntheticClip(100, T)
frames = [clip.frame(t, 1) for t in range(T)]
src = FeatureClip(lambda t: frames[t], T, resident_device=dev)
point = clip.point_prompt(1)["point_coords"][0].tolist()
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    state = predictor.init_state(src)
    predictor.add_new_points_or_box(state, 0, 1, points=point, labels=[1])
    stamps = []
    for f, ids, m in predictor.propagate_in_video(state):
        if f in (0, 15, 16, 17, 18):
            torch.cuda.synchronize()
            stamps.append((f, time.perf_counter() - t0))
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print(f"clip {rep}: {total * 1e3:.1f} ms for {T} frames = {T / total:.1f} frames/s; cumulative ms at frames {[(f, round(t * 1e3, 1)) for f, t in stamps]}", flush=True)
