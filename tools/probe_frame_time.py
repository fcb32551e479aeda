"""Dev probe: host vs device time per steady-state frame (not part of the product or the bench contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import build_sam, synth, _lib
from video_llava_seg_b200.features import FeatureClip

dev = torch.device("cuda:0")
predictor = build_sam.build_sam2_video_predictor(None, synth.init_state_dict(0), dev)
T = 60
clip = synth.SyntheticClip(100, T)
frames = [clip.frame(t, 1) for t in range(T)]
src = FeatureClip(lambda t: frames[t], T, resident_device=dev)
pinned = FeatureClip(lambda t: frames[t], T, pinned=True)
for mode in ("nosync", "sync", "nosync", "pinned-sync", "pinned-sync-eager"):
    if mode.startswith("pinned"):
        src_use = pinned
        predictor.use_cuda_graph = not mode.endswith("eager")
    else:
        src_use = src
    state = predictor.init_state(src_use)
    predictor.add_new_points_or_box(state, 0, 1, points=clip.point_prompt(1)["point_coords"][0].tolist(), labels=[1])
    gen = predictor.propagate_in_video(state)
    for _ in range(20):
        next(gen)
    torch.cuda.synchronize()
    n = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cpu = []
    e0.record()
    t_all = time.perf_counter()
    for _ in range(n):
        t0 = time.perf_counter()
        next(gen)
        if "sync" in mode and mode != "nosync":
            torch.cuda.synchronize()
        cpu.append(time.perf_counter() - t0)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_all
    print(f"{mode}: gpu span {e0.elapsed_time(e1)/n:.3f} ms/frame, cpu loop {sum(cpu)/n*1e3:.3f} ms/frame (min {min(cpu)*1e3:.3f}), wall {wall/n*1e3:.3f}")
    gen.close()
# where does host time go? coarse profile of one frame
import cProfile, pstats
state = predictor.init_state(src)
predictor.add_new_points_or_box(state, 0, 1, points=clip.point_prompt(1)["point_coords"][0].tolist(), labels=[1])
gen = predictor.propagate_in_video(state)
for _ in range(20):
    next(gen)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    next(gen)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr).sort_stats("cumulative")
st.print_stats(28)
