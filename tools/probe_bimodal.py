"""Dev probe: is the steady-state frame time a property of the graph capture (it varies by ~6 % between captures)?"""
import os, sys, re, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from video_llava_seg_b200 import build_sam, synth
from video_llava_seg_b200.features import FeatureClip

dev = torch.device("cuda:0")
predictor = build_sam.build_sam2_video_predictor(None, synth.init_state_dict(0), dev)
T = 70
clip = synth.SyntheticClip(100, T)
frames = [clip.frame(t, 1) for t in range(T)]
src = FeatureClip(lambda t: frames[t], T, resident_device=dev)
point = clip.point_prompt(1)["point_coords"][0].tolist()
results = []
for rep in range(6):
    state = predictor.init_state(src)
    predictor.add_new_points_or_box(state, 0, 1, points=point, labels=[1])
    gen = predictor.propagate_in_video(state)
    for _ in range(22):
        next(gen)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        next(gen)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            next(gen)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    ax = [i for i, e in enumerate(evs) if "axpy_rows" in e.name]
    firsts = [i for k, i in enumerate(ax) if k % 5 == 0]
    a_, b_ = firsts[-2], firsts[-1]
    base = evs[a_].time_range.start
    line = [(e.time_range.start - base, e.time_range.end - e.time_range.start, re.sub(r"vls::\(anonymous namespace\)::", "", e.name).replace("void ", "")[:30]) for e in evs[a_:b_]]
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            n = re.sub(r"vls::\(anonymous namespace\)::", "", e.name).replace("void ", "")[:34]
            a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += (e.time_range.end - e.time_range.start) / 4
    results.append((ms, agg, line))
    print(f"capture {rep}: {ms:.4f} ms/frame", flush=True)
    gen.close()
results.sort(key=lambda r: r[0])
fast, slow = results[0], results[-1]
print(f"fastest {fast[0]:.4f} vs slowest {slow[0]:.4f}: per-kernel us/frame (slow - fast), |diff| > 1.5")
for n in slow[1]:
    d = slow[1][n][1] - fast[1].get(n, [0, 0.0])[1]
    if abs(d) > 1.5:
        print(f"  {d:+8.1f}  {slow[1][n][1]:8.1f} vs {fast[1].get(n, [0, 0.0])[1]:8.1f}  x{slow[1][n][0] // 4}  {n}")

print("side by side (start us, dur): fast | slow")
for (fs, fd, fn), (ss, sd, sn) in zip(fast[2], slow[2]):
    flag = " <<<" if abs((ss - fs)) > 15 and fn == sn else ""
    print(f"{fs:8.1f} {fd:6.1f} {fn:30s} | {ss:8.1f} {sd:6.1f} {sn:30s}{flag}")
