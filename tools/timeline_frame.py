"""Dev probe: real (warm, in-graph) per-kernel timeline of steady-state frames via torch.profiler (CUPTI)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from video_llava_seg_b200 import build_sam, synth
from video_llava_seg_b200.features import FeatureClip

dev = torch.device("cuda:0")
predictor = build_sam.build_sam2_video_predictor(None, synth.init_state_dict(0), dev)
T = 30
B = int(os.environ.get("VLS_TL_OBJECTS", "1"))     # objects tracked jointly (BASELINE configs[2]: 8)
clip = synth.SyntheticClip(100, T)
frames = [clip.frame(t, 1) for t in range(T)]
src = FeatureClip(lambda t: frames[t], T, resident_device=dev)
state = predictor.init_state(src)
for o, pt in enumerate(clip.point_prompt(B)["point_coords"]):
    predictor.add_new_points_or_box(state, 0, o + 1, points=pt.tolist(), labels=[1])
gen = predictor.propagate_in_video(state)
for _ in range(22):
    next(gen)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        next(gen)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
print("cuda events:", len(ev))
# split into frames by large gaps
t0 = ev[0].time_range.start
rows = [(e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:60]) for e in ev]
# a frame's graph replay begins with the first axpy_rows kernel after the staging copies (multi_copy / memcpy) of the host loop
firsts = [i for i, r in enumerate(rows) if "axpy_rows" in r[2] and i > 0 and ("multi_copy" in rows[i - 1][2] or "Memcpy" in rows[i - 1][2])]
print("frame starts", firsts[:8])
a, b = firsts[-2], firsts[-1]
base = rows[a][0]
tot = 0.0
for s, d, n in rows[a:b]:
    print(f"{s - base:9.1f} {d:7.1f}  {n}")
    tot += d
print("frame span us", rows[b][0] - base, "sum of kernel durations", tot)
gen.close()   # (a generator finalised at interpreter exit prints a traceback from torch.no_grad's context manager)
