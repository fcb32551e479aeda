"""Dev probe: (a) C independent clips tracked concurrently on C streams (one predictor each), (b) B objects batched in
one clip.  Reports aggregate frames/s (a) and frame-objects/s (b)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import build_sam, synth
from video_llava_seg_b200.features import FeatureClip

dev = torch.device("cuda:0")
T = 60
sd = synth.init_state_dict(0)

def make(seed, batch):
    clip = synth.SyntheticClip(seed, T)
    frames = [clip.frame(t, 1) for t in range(T)]
    return clip, FeatureClip(lambda t: frames[t], T, resident_device=dev)

for C in (1, 2, 3, 4):
    preds = [build_sam.build_sam2_video_predictor(None, sd, dev) for _ in range(C)]
    streams = [torch.cuda.Stream() for _ in range(C)]
    gens = []
    for i in range(C):
        clip, src = make(100 + i, 1)
        with torch.cuda.stream(streams[i]):
            st = preds[i].init_state(src)
            preds[i].add_new_points_or_box(st, 0, 1, points=clip.point_prompt(1)["point_coords"][0].tolist(), labels=[1])
            g = preds[i].propagate_in_video(st)
            for _ in range(22):
                next(g)
        gens.append(g)
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        for i in range(C):
            with torch.cuda.stream(streams[i]):
                next(gens[i])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"concurrent clips C={C}: {C * n / dt:.1f} frames/s total ({dt / n * 1e3:.3f} ms per round)", flush=True)
    for g in gens:
        g.close()
    del preds, gens

pred = build_sam.build_sam2_video_predictor(None, sd, dev)
for B in (1, 2, 4, 8):
    clip, src = make(100, B)
    st = pred.init_state(src)
    prompt = clip.point_prompt(B)
    for o in range(B):
        pred.add_new_points_or_box(st, 0, o + 1, points=prompt["point_coords"][o].tolist(), labels=[1])
    g = pred.propagate_in_video(st)
    for _ in range(22):
        next(g)
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        next(g)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"batched objects B={B}: {n / dt:.1f} frames/s = {B * n / dt:.1f} frame-objects/s ({dt / n * 1e3:.3f} ms/frame)", flush=True)
    g.close()
