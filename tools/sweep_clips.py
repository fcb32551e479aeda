"""BASELINE configs[4]: throughput sweep over many short clips, sharded by video across the GPUs of one box exactly as
the reference's inference driver shards its video list (llava/inference/main.py:41-49, scripts/infer.sh:1-7): every
rank owns a contiguous chunk of the clip list, tracks its clips one after the other, no collective on the data path.

    python tools/sweep_clips.py --clips 64 --frames 64                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
        tools/sweep_clips.py --clips 512 --frames 64                         # 8 GPUs

Synthetic backbone features are expensive to generate on the CPU (1 GB per 64-frame clip), so a pool of `--pool`
distinct clips is generated once per rank and clip i uses pool[i % pool] (with its own prompt); features stay resident.
Prints one JSON line from rank 0 (clips/s and frames/s of the whole job, max device+host time over ranks)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from video_llava_seg_b200 import build_sam, synth
from video_llava_seg_b200.features import FeatureClip
from video_llava_seg_b200.shard import aggregate_throughput, shard_clips

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=64)
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--pool", type=int, default=2)
ap.add_argument("--streams", type=int, default=1, help="clips tracked concurrently per GPU (one predictor + CUDA stream each)")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sd = synth.init_state_dict(0)
# one predictor (own workspaces / captured graphs) and one stream per concurrently tracked clip: a single clip is a chain
# of small kernels that leaves SMs idle, two or three chains side by side fill them (+29 % / +40 % on one B200)
lanes = [(build_sam.build_sam2_video_predictor(None, sd, dev), torch.cuda.Stream(device=dev)) for _ in range(args.streams)]
for pr, _ in lanes:
    pr.output_mode = "binary"
mine = shard_clips(range(args.clips), world, rank)
pool = []
for p in range(args.pool):
    clip = synth.SyntheticClip(1000 + rank * args.pool + p, args.frames)
    frames = [clip.frame(t, 1) for t in range(args.frames)]
    pool.append((clip, FeatureClip(lambda t, fr=frames: fr[t], args.frames, resident_device=dev)))


def session(pr, i):
    clip, src = pool[i % args.pool]
    st = pr.init_state(src)
    pr.add_new_points_or_box(st, 0, 1, points=clip.point_prompt(1)["point_coords"][0].tolist(), labels=[1])
    return pr.propagate_in_video(st)


def run(clip_ids):
    """Round-robin over the lanes: every lane advances its current clip by one frame per round."""
    todo = list(clip_ids)
    gens = [None] * len(lanes)
    last = None
    while todo or any(g is not None for g in gens):
        for k, (pr, stream) in enumerate(lanes):
            with torch.cuda.stream(stream):
                if gens[k] is None:
                    if not todo:
                        continue
                    gens[k] = session(pr, todo.pop(0))
                try:
                    _, _, last = next(gens[k])      # a consumer would ship the uint8 mask
                except StopIteration:
                    gens[k] = None
    return last


run(range(2 * len(lanes)))      # warm-up: every lane captures its steady-state graph once
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
run(mine)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3
clips_s, ms_max, n = aggregate_throughput(len(mine), ms, dev)
if rank == 0:
    print(json.dumps({"workload": f"{args.clips} clips x {args.frames} frames, 1 object, sharded by clip over {world} GPU(s), {args.streams} concurrent clip(s) per GPU",
                      "clips_per_s": round(clips_s, 2), "frames_per_s": round(clips_s * args.frames, 1),
                      "slowest_rank_ms": round(ms_max, 1), "clips": int(n)}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
