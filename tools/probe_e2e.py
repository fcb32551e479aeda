"""Dev probe: end-to-end (pinned host features) frame time and raw H2D bandwidth (not part of the bench contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import build_sam, synth
from video_llava_seg_b200.features import FeatureClip

dev = torch.device("cuda:0")
h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
print(f"H2D pinned: {10 * 64 / 1024 / (e0.elapsed_time(e1) / 1e3):.1f} GiB/s", flush=True)

predictor = build_sam.build_sam2_video_predictor(None, synth.init_state_dict(0), dev)
T = 60
clip = synth.SyntheticClip(100, T)
frames = [clip.frame(t, 1) for t in range(T)]
for name, kw in (("resident", dict(resident_device=dev)), ("pinned", dict(pinned=True)), ("pinned", dict(pinned=True))):
    src = FeatureClip(lambda t: frames[t], T, **kw)
    state = predictor.init_state(src)
    predictor.add_new_points_or_box(state, 0, 1, points=clip.point_prompt(1)["point_coords"][0].tolist(), labels=[1])
    gen = predictor.propagate_in_video(state)
    for _ in range(22):
        _, _, m = next(gen)
        (m > 0).to(torch.uint8).cpu()
    torch.cuda.synchronize()
    n = 30
    per = []
    t_all = time.perf_counter()
    for _ in range(n):
        t0 = time.perf_counter()
        _, _, m = next(gen)
        t1 = time.perf_counter()
        (m > 0).to(torch.uint8).cpu()
        per.append((t1 - t0, time.perf_counter() - t1))
    wall = time.perf_counter() - t_all
    print(f"{name}: wall {wall / n * 1e3:.3f} ms/frame; next() {sum(p[0] for p in per) / n * 1e3:.3f} ms, readback {sum(p[1] for p in per) / n * 1e3:.3f} ms", flush=True)
    gen.close()
