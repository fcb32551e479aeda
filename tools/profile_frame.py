"""Dev/profiling driver: runs the propagation to steady state, then brackets N frames with
cudaProfilerStart/Stop so that `ncu --profile-from-start off` only instruments those frames.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/profile_frame.py
Not part of the product path or of the bench contract."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import build_sam, synth
from video_llava_seg_b200.features import FeatureClip

frames_to_profile = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
predictor = build_sam.build_sam2_video_predictor(None, synth.init_state_dict(0), dev)
T = 20 + frames_to_profile
clip = synth.SyntheticClip(100, T)
frames = [clip.frame(t, 1) for t in range(T)]
src = FeatureClip(lambda t: frames[t], T, resident_device=dev)
state = predictor.init_state(src)
prompt = clip.point_prompt(batch)
for o in range(batch):
    predictor.add_new_points_or_box(state, 0, o + 1, points=prompt["point_coords"][o].tolist(), labels=[1])
gen = predictor.propagate_in_video(state)
for _ in range(19):
    next(gen)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(frames_to_profile):
    next(gen)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
gen.close()
print("profiled", frames_to_profile, "frames, batch", batch)
