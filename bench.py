#!/usr/bin/env python
"""bench.py -- propagated frames/s of the SAM 2.1 mask-propagation hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                 # our arm (sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W          # CPU arm: oracle port of the reference path
    python bench.py --workload "configs[2]" ...                    # 8 objects tracked jointly
    python bench.py --workload "configs[3]" ...                    # LG-VIS: [SEG] prompt embeddings -> decoder + propagation
    python bench.py --workload "configs[4]" ...                    # clip sweep sharded by video across the ranks

Default workload = BASELINE.json configs[1]: Hiera-B+ propagation at 1024^2 (64x64 tokens), 7-frame memory bank +
16 object pointers (Nk = 28 736), 1 object per GPU, synthetic clip, random-init weights.  The image encoder is
outside the hot path: clips are given as backbone features.  One step = one propagated frame in steady state
(the bank is filled during an untimed 17-frame ramp).  N > 1: one process per GPU (torchrun), one clip per rank,
no collective on the data path ("scaling": "weak"); NCCL only for the barrier / max-over-ranks of the time.

Timing: exactly K steps per window, each step bracketed by CUDA events with an L2 flush in between; the window is
repeated R times back to back (K * R >= 200 by default) and the line reports the MEDIAN window plus every window's
ms/step, so a 30 ms sample is never the whole evidence.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")

UNIT = "frames/s"
RAMP = 17  # prompt frame + 16 propagated frames: full memory bank and 16 pointers afterwards
NQ, NK_STEADY, D, DV = 4096, 7 * 4096 + 64, 256, 64
PERIOD = 32  # distinct synthetic frames per clip; longer clips cycle through them (the tracker state keeps evolving)

WORKLOADS = {
    "configs[1]": dict(objects=1, kind="frames", metric="propagated frames/sec (SAM2.1 Hiera-B+ hot path, 1024^2, 7-frame memory bank, 1 object per GPU)",
                       text="configs[1]: SAM2.1 Hiera-B+ propagation, 64x64 tokens, 7 memories + 16 pointers (Nk=28736), 1 object, synthetic"),
    "configs[2]": dict(objects=8, kind="frames", metric="propagated frames/sec (SAM2.1 hot path, 1024^2, 7-frame memory bank, 8 objects tracked jointly per GPU)",
                       text="configs[2]: SAM2.1 propagation (Hiera-L shares the hot-path shape), 8 objects tracked jointly (batched memory bank and object pointers), Nk=28736, synthetic"),
    "configs[3]": dict(objects=1, kind="clips", frames=32, metric="frames/sec, LG-VIS inference: [SEG] prompt embedding -> SAM2 mask decoder + propagation, 32-frame clips",
                       text="configs[3]: Video-LLaVA-Seg LG-VIS inference: a random 4096-d [SEG] hidden state -> seg-head projection -> prompt embedding on frame 0 -> mask decoder + propagation, 32 frames, random-init weights"),
    "configs[4]": dict(objects=1, kind="clips", frames=64, metric="frames/sec, throughput sweep: 64-frame clips sharded by video across the GPUs",
                       text="configs[4]: throughput sweep, 512 synthetic clips x 64 frames, 1 object, sharded by video across the ranks (contiguous chunks), no collective"),
}


def make_config(key, world):
    """The `config` object of BOTH arms (the driver compares them)."""
    w = WORKLOADS[key]
    return {"workload": w["text"], "objects_per_gpu": w["objects"], "clips_per_gpu": 1, "ramp_frames": RAMP,
            "l2": "flushed between timed steps (256 MiB memset, outside the per-step events)",
            "parallelism": f"{world} independent replica(s), sharded by clip, no collective on the path",
            "step": "one propagated frame in steady state (full bank)" if w["kind"] == "frames" else f"one whole {w['frames']}-frame clip (prompt frame, ramp, steady state)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(burst=float(j.get("bf16_tflops", 1590.0)), sustained=float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1400.0))),
                    hbm=float(j["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def pick_tensor_peak(clocks):
    """r1 verdict: a kernel that ran at the maximum SM clock with no power cap is held against the BURST cuBLAS figure;
    the sustained figure (measured power-capped at ~1.3 GHz) only applies when the window itself was capped."""
    pk = measured_peaks()
    mhz, mx = clocks.get("sm_mhz"), clocks.get("sm_max_mhz")
    capped = "sw_power_cap" in (clocks.get("reasons") or []) or (mhz and mx and mhz < 0.97 * mx)
    which = "sustained" if capped else "burst"
    return pk[which], f"{pk['source']}, {which} bf16 ({'power cap / reduced SM clock seen' if capped else 'SM clock at max, no cap'} in the timed window)"


NVML_SAMPLER = r"""
import sys, time
import pynvml as N
N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
bits = (("hw_slowdown", N.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksThrottleReasonHwThermalSlowdown),
        ("sw_thermal_slowdown", N.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksThrottleReasonSwPowerCap))
while True:
    mhz = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
    r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    f = ["Active" if r & b else "Not Active" for _, b in bits]
    print(f"{time.time()!r}, {mhz}, {mx}, " + ", ".join(f), flush=True)
    time.sleep(float(sys.argv[2]))
"""


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed window by a separate process (NVML through pynvml, one
    light query every 25 ms -- at 4 ms the queries themselves stalled about one 16 MB H2D step per run by 6-20 ms; `nvidia-smi -lms` as the fallback).  A child process, not a thread: a Python sampler
    thread would contend for the GIL with the loop that enqueues the frames, and a polling `nvidia-smi` takes driver
    locks for milliseconds, which showed up as 3-20 ms stalls of individual timed steps."""

    def __init__(self, index):
        self.samples, self.max_mhz, self.t0, self.t1 = [], None, None, None
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            import pynvml  # noqa: F401

            self.cmd, self.stamped = [sys.executable, "-c", NVML_SAMPLER, str(index), os.environ.get("VLS_BENCH_SAMPLE_S", "0.025")], True
        except ImportError:
            self.cmd = ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)]
            self.stamped = False
        self.proc, self.thread = None, None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                stamp = float(f.pop(0)) if self.stamped else time.time()
                mhz = float(f[0])
                self.max_mhz = float(f[1])
                active = [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")]
                self.samples.append((stamp, mhz, active))
            except (ValueError, IndexError):
                pass

    def start(self, ready_timeout=15.0):
        """Spawn the sampler once per process and wait for its first sample: nvmlInit / nvidia-smi start-up holds
        driver locks for 0.1-1 s on a fresh box and must not overlap a timed window."""
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t_end = time.time() + ready_timeout
            while not self.samples and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.01)
        except OSError:
            self.proc = None
        return self

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            self.thread.join(timeout=2)
            self.proc = None

    def __enter__(self):
        return self

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def __exit__(self, *a):
        time.sleep(0.02)      # let the last in-window samples arrive

    def summary(self):
        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        m0, m1 = (0.0, 0.0) if self.stamped else (0.02, 0.04)   # nvidia-smi lines are stamped on arrival
        win = [s for s in self.samples if t0 - m0 <= s[0] <= t1 + m1]
        scope = "timed window"
        if not win:  # window shorter than the sampling period: fall back to the loaded ramp just before it
            win, scope = [s for s in self.samples if s[0] <= t1 + m1][-10:], "ramp + timed window"
        reasons = sorted({r for s in win for r in s[2]})
        return {"sm_mhz": statistics.median([s[1] for s in win]) if win else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(win), "scope": scope}


def make_clip(seed, num_frames):
    from video_llava_seg_b200 import synth

    return synth.SyntheticClip(seed, num_frames)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from video_llava_seg_b200 import _lib, build_sam, synth
    from video_llava_seg_b200.features import FeatureClip
    from video_llava_seg_b200.shard import aggregate_throughput, shard_clips

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py measures the sm_100a kernels: a CUDA device is required (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    wl = WORKLOADS[args.workload]
    B = wl["objects"]
    sampler = ClockSampler(local).start()
    predictor = build_sam.build_sam2_video_predictor(None, synth.init_state_dict(0), dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    config = make_config(args.workload, world)
    method = {}   # how this arm runs and times the workload: kept OUT of `config`, which both arms must share verbatim
    method["execution"] = ("steady-state frames replay one CUDA graph (graphed.py); gpu_launches counts the library kernels inside each replay. "
                           "Frames are software-pipelined: a replay runs frame t from its first cross-attention on plus the feature-only head "
                           "of frame t+1's memory attention (next to frame t's decoder / memory encoder), so every step still executes each "
                           "kernel of the path exactly once")
    method["e2e_path"] = ("pinned host features -> H2D one frame ahead on a copy stream (3 staging sets) -> propagate_in_video(output_mode='binary': fused "
                          "resize+threshold) -> uint8 mask D2H into pinned memory every step on a copy stream; step i's events close over the "
                          "read-back of step i-1 (K steps = K complete read-backs), consumer pipelined by one frame")
    line = {"metric": wl["metric"], "unit": UNIT, "n_gpus": world, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic backbone features + seeded random-init weights",
            "config": config, "method": method}

    clip = make_clip(100 + rank, PERIOD)
    frames = [clip.frame(t, 1) for t in range(PERIOD)]
    prompts = clip.point_prompt(B)["point_coords"]

    def prompt_all(state, pred=None):
        for o in range(B):
            (pred or predictor).add_new_points_or_box(state, 0, o + 1, points=prompts[o].tolist(), labels=[1])

    if wl["kind"] == "clips":
        run_clip_workload(args, line, wl, predictor, frames, prompts, sampler, lib, dev, world, rank, torch, dist,
                          FeatureClip, aggregate_throughput, shard_clips)
        return

    K, W = args.steps, max(args.warmup, 3)
    R = args.repeats if args.repeats > 0 else max(1, math.ceil(200 / K))
    T = RAMP + W + K * R + 1

    def timed_pass(source, d2h, predictor=predictor, R=R, init_kw=None):
        """Ramp + warm-up untimed, then R windows of exactly K steps, each step bracketed by CUDA events, L2 flushed between
        steps.  With d2h the binarised video-resolution mask of EVERY step is read back into pinned host memory inside the
        step's events; the consumer is software-pipelined by one frame (it waits for frame t-1's mask after frame t
        has been enqueued), as a streaming client of propagate_in_video would be."""
        import gc

        host = [torch.empty((B, 1, 1024, 1024), dtype=torch.uint8).pin_memory() for _ in range(2)] if d2h else None
        done = [torch.cuda.Event(), torch.cuda.Event()]
        checksum = 0
        main = torch.cuda.current_stream(dev)
        d2h_stream = torch.cuda.Stream(device=dev) if d2h else None

        def read_back(m, slot):   # with d2h the predictor runs in output_mode "binary": m is already the uint8 mask
            # the copy engine reads the mask back on its own stream while the next frame is tracked (as a streaming client
            # would); step i's timed region ends with a wait for step i-1's copy, so every window of K steps contains K
            # complete read-backs
            ready = torch.cuda.Event()
            ready.record(main)
            d2h_stream.wait_event(ready)
            with torch.cuda.stream(d2h_stream):
                host[slot].copy_(m, non_blocking=True)
                done[slot].record(d2h_stream)
            m.record_stream(d2h_stream)

        predictor.output_mode = "binary" if d2h else "logits"
        with sampler as clocks:
            state = predictor.init_state(source, **(init_kw or {}))
            prompt_all(state, predictor)
            gen = predictor.propagate_in_video(state)
            for j in range(RAMP + W):
                _, _, m = next(gen)
                if d2h:
                    read_back(m, j % 2)
                    done[j % 2].synchronize()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            n = K * R
            starts = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
            stops = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
            launches0 = lib.vls_launch_count()
            out_bytes = 0
            gc.collect()
            gc.disable()          # a generation-2 collection inside a 2 ms step is host noise, not the path
            clocks.mark_start()
            host_wait = 0.0                                  # seconds the host spent blocked on the device (debug)
            t_loop = time.perf_counter()
            for i in range(n):
                flush.zero_()
                starts[i].record()
                _, _, m = next(gen)
                if d2h:
                    read_back(m, i % 2)                      # result of step i -> pinned host memory (copy stream)
                    out_bytes = host[i % 2].numel()
                    main.wait_event(done[(i - 1) % 2])       # ... and the previous step's read-back ends inside this step
                stops[i].record()
                if d2h and i > 0:
                    t_w = time.perf_counter()
                    done[(i - 1) % 2].synchronize()          # consume step i-1's mask on the host
                    host_wait += time.perf_counter() - t_w
                    checksum += int(host[(i - 1) % 2][0, 0, 0, 0])
            t_loop = time.perf_counter() - t_loop
            torch.cuda.synchronize()
            clocks.mark_stop()
            gc.enable()
            predictor.output_mode = "logits"
        launches = lib.vls_launch_count() - launches0
        per_step = [s.elapsed_time(e) for s, e in zip(starts, stops)]
        if os.environ.get("VLS_BENCH_DEBUG"):
            print(f"[bench debug] d2h={d2h} per-step ms: {[round(x, 3) for x in per_step]}", file=sys.stderr, flush=True)
            print(f"[bench debug] d2h={d2h} host: {1e3 * t_loop / n:.3f} ms per iteration enqueued, of which "
                  f"{1e3 * host_wait / n:.3f} ms blocked on the device", file=sys.stderr, flush=True)
        if world > 1:
            dist.barrier()
        windows = []
        for r in range(R):        # every window: units of all ranks / slowest rank's device time
            fps, ms, _ = aggregate_throughput(K, sum(per_step[r * K:(r + 1) * K]), dev)
            windows.append((fps, ms))
        gen.close()
        med = sorted(windows, key=lambda w: w[1])[len(windows) // 2]
        return med[0], med[1], launches // R, clocks.summary(), out_bytes, [round(w[1] / K, 4) for w in windows]

    # (1) device-resident inputs: kernel + host-orchestration throughput
    resident = FeatureClip(lambda t: frames[t], T, resident_device=dev, period=PERIOD)
    # one untimed pass over a short clip first: the predictor keeps every frame's outputs (as the reference
    # does), so a fresh process would otherwise time cudaMalloc growth of the caching allocator, not the path
    warm = predictor.init_state(FeatureClip(lambda t: frames[t], RAMP + 8, resident_device=dev, period=PERIOD))
    prompt_all(warm)
    for _ in predictor.propagate_in_video(warm):
        pass
    del warm
    torch.cuda.synchronize()
    value, ms, launches, clocks, _, win_ms = timed_pass(resident, d2h=False)
    # (2) end to end through the public API with host buffers: H2D of each frame's features, D2H of the mask
    pinned = FeatureClip(lambda t: frames[t], T, pinned=True, period=PERIOD)
    e2e, _, _, _, out_bytes, e2e_win = timed_pass(pinned, d2h=True)
    sampler.stop()

    # (3) informational: whole 64-frame clips (the workload as a user runs it: prompt frame, 16 frames with a growing
    # bank, then steady state; captured graph re-used from clip to clip), wall clock around complete sessions
    def whole_clips(n_clips=3, T_clip=64):
        src = FeatureClip(lambda t: frames[t], T_clip, resident_device=dev, period=PERIOD)
        predictor.output_mode = "logits"
        times = []
        for _ in range(n_clips + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = predictor.init_state(src)
            prompt_all(st)
            for _ in predictor.propagate_in_video(st):
                pass
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        best = min(times[1:])
        return {"frames_per_clip": T_clip, "value": round(T_clip / best, 1), "unit": UNIT, "ms_per_clip": round(best * 1e3, 2),
                "note": "complete sessions incl. prompt frame and 16-frame ramp, resident features, best of %d" % n_clips}

    clip_info = whole_clips()
    pixels_info = None
    if rank == 0 and world == 1 and B == 1 and not args.no_pixels:
        pixels_info = from_pixels(timed_pass, build_sam, synth, dev, torch, RAMP + W + K + 1)
    method["timing"] = f"{R} back-to-back windows of exactly {K} steps; value / ms_per_step / e2e are the MEDIAN window; windows_ms_per_step lists all"
    line.update({"value": round(value, 3), "steps": K, "warmup": W, "ms_per_step": round(ms / K, 4),
                 "windows_ms_per_step": win_ms, "spread": {"min": min(win_ms), "max": max(win_ms), "windows": R},
                 "clocks": clocks, "gpu_launches": int(launches), "whole_clip": clip_info,
                 "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": int(pinned.h2d_bytes_per_frame),
                         "d2h_bytes_per_step": int(out_bytes), "windows_ms_per_step": e2e_win}})
    if pixels_info is not None:
        line["e2e_from_pixels"] = pixels_info
    if B > 1:
        line["frame_objects_per_s"] = round(value * B, 1)
    if rank == 0:
        line["roofline"] = roofline(predictor, resident, prompt_all, lib, torch, B, clocks)
        if world == 1:
            try:
                from video_llava_seg_b200 import kernel_bench

                line["roofline_hbm"] = kernel_bench.hbm_rooflines(dev, measured_peaks()["hbm"])
            except Exception as e:  # informational leg: never lose the bench line over it
                line["roofline_hbm"] = {"unavailable": repr(e)[:200]}
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline(steps=2, objects=B)
                line["gpu_eager_baseline"] = gpu_eager_baseline(dev, objects=B)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def from_pixels(timed_pass, build_sam, synth, dev, torch, T):
    """Informational second end-to-end number: the SAME public call path, but starting
    from PIXELS -- normalised f32 frames in pinned host memory, uploaded frame by frame (offload_video_to_cpu=True), the
    Hiera-B+ + FPN image encoder hosted in PyTorch (bf16, one CUDA graph), then the CUDA hot path and the uint8 mask read
    back every step.  One window; not the roofline basis (the encoder is library code: cuDNN / cuBLAS / SDPA)."""
    try:
        sd = dict(synth.init_state_dict(0))
        sd.update({"image_encoder." + k: v for k, v in synth.init_image_encoder_state_dict("b+", 0).items()})
        pred = build_sam.build_sam2_video_predictor("b+", sd, dev, image_encoder_dtype=torch.bfloat16)
        base = synth.synthetic_frames(8, 1024, seed=1)                      # 8 distinct frames, cycled (12.6 MB each)
        frames = base[torch.arange(T) % 8].contiguous().pin_memory()
        fps, ms, _, _, out_bytes, win = timed_pass(frames, True, predictor=pred, R=1, init_kw={"offload_video_to_cpu": True})
        enc_ms = None
        x = frames[:1].to(dev)
        for _ in range(3):
            pred.forward_image(x)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            pred.forward_image(x)
        b.record()
        torch.cuda.synchronize()
        enc_ms = a.elapsed_time(b) / 10
        del pred
        return {"value": round(fps, 2), "unit": UNIT, "ms_per_step": win[0], "image_encoder_ms": round(enc_ms, 3),
                "h2d_bytes_per_step": int(frames[0].numel() * frames[0].element_size()), "d2h_bytes_per_step": int(out_bytes),
                "note": "Hiera-B+ + FPN image encoder in PyTorch (bf16, CUDA graph) + the CUDA hot path, from pinned f32 frames; "
                        "informational, not the headline metric"}
    except Exception as e:  # informational leg: never lose the bench line over it
        return {"unavailable": repr(e)[:300]}


def run_clip_workload(args, line, wl, predictor, frames, prompts, sampler, lib, dev, world, rank, torch, dist, FeatureClip,
                      aggregate_throughput, shard_clips):
    """configs[3] / configs[4]: a step is one WHOLE clip through the public API (init_state, prompt, propagate_in_video),
    clips sharded by video across the ranks (contiguous chunks, llava/inference/main.py:41-49), no collective.  Timed on
    the device with CUDA events around each rank's chunk (value) and again with pinned host features + uint8 masks read
    back every frame (e2e).  configs[3] prompts with a [SEG]-token embedding through the seg head's projection."""
    T_clip = wl["frames"]
    total = args.sweep_clips if args.workload == "configs[4]" else max(world, args.steps * world)
    mine = list(shard_clips(range(total), world, rank))
    seg_head = None
    if args.workload == "configs[3]":
        from video_llava_seg_b200.llava_seg_head import SegmentationHeadSAM2

        seg_head = SegmentationHeadSAM2(n_token_dims=4096, n_seg_queries=1, sam2_model=predictor).to(dev)
        seg_tokens = torch.randn(64, 4096, generator=torch.Generator().manual_seed(5)).to(dev).bfloat16()

    def session(src, i):
        st = predictor.init_state(src)
        if seg_head is not None:
            emb = seg_head.project_tokens(seg_tokens[i % 64: i % 64 + 1])        # [1,1,256]
            predictor.add_new_prompt_embedding(st, 0, 1, emb[0])
        else:
            predictor.add_new_points_or_box(st, 0, 1, points=prompts[0].tolist(), labels=[1])
        return predictor.propagate_in_video(st)

    def sweep(src, d2h):
        predictor.output_mode = "binary" if d2h else "logits"
        host = [torch.empty((1, 1, 1024, 1024), dtype=torch.uint8).pin_memory() for _ in range(2)] if d2h else None
        done = [torch.cuda.Event(), torch.cuda.Event()]
        for i in range(2):                      # warm-up: graphs captured, allocator grown
            for _ in session(src, i):
                pass
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = lib.vls_launch_count()
        with sampler as clocks:
            clocks.mark_start()
            a.record()
            k = 0
            for i in mine:
                for _, _, m in session(src, i):
                    if d2h:
                        host[k % 2].copy_(m, non_blocking=True)
                        done[k % 2].record()
                        if k > 0:
                            done[(k - 1) % 2].synchronize()
                        k += 1
            b.record()
            torch.cuda.synchronize()
            clocks.mark_stop()
        predictor.output_mode = "logits"
        if world > 1:
            dist.barrier()
        fps, ms, n = aggregate_throughput(len(mine) * T_clip, a.elapsed_time(b), dev)
        return fps, ms, lib.vls_launch_count() - launches0, clocks.summary(), int(n)

    resident = FeatureClip(lambda t: frames[t], T_clip, resident_device=dev, period=PERIOD)
    value, ms, launches, clocks, n_frames = sweep(resident, False)
    pinned = FeatureClip(lambda t: frames[t], T_clip, pinned=True, period=PERIOD)
    e2e, e2e_ms, _, _, _ = sweep(pinned, True)
    sampler.stop()
    line["config"]["clips_total"] = total
    line["config"]["clips_per_gpu"] = len(mine)
    line["config"]["features"] = f"{PERIOD} distinct synthetic frames per rank, cycled (CPU feature synthesis costs 50 ms per frame)"
    line.update({"value": round(value, 3), "steps": len(mine), "warmup": 2, "ms_per_step": round(ms / max(len(mine), 1), 3),
                 "clips_per_s": round(value / T_clip, 2), "clocks": clocks, "gpu_launches": int(launches),
                 "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": int(pinned.h2d_bytes_per_frame) * T_clip,
                         "d2h_bytes_per_step": 1024 * 1024 * T_clip},
                 "roofline": None, "note": "roofline / cpu_baseline are reported on the configs[1] line (same kernels)"})
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def roofline(predictor, source, prompt_all, lib, torch, B, clocks):
    """Dominant kernel = attn_fwd_kernel on the memory cross-attention (4 launches / frame), timed TOGETHER with its
    combine launch by CUDA events on its stream.  The reference-ALGORITHMIC work per launch is 4 * Nq * Nk * 256 FLOPs per
    object (QK^T + PV over 256-d values, sam/transformer.py:311-360); the kernel EXECUTES 2 * Nq * Nk * (256 + 64) because it
    attends over the 64-d memory and folds the value projection into the output projection.  `achieved` / `frac` are the
    EXECUTED rate (hardware utilisation, never above the peak); `algorithmic` / `frac_algorithmic` the reference-equivalent one."""
    import ctypes

    was = predictor.use_cuda_graph
    predictor.use_cuda_graph = False   # the same kernels, launched eagerly so that each launch can be bracketed by events
    try:
        state = predictor.init_state(source)
        prompt_all(state)
        gen = predictor.propagate_in_video(state)
        for _ in range(RAMP + 2):
            next(gen)
        torch.cuda.synchronize()
        lib.vls_prof_enable(1)
        for _ in range(6):
            next(gen)
        torch.cuda.synchronize()
        lib.vls_prof_enable(0)
        gen.close()
    finally:
        predictor.use_cuda_graph = was
    cnt, tot = ctypes.c_int(0), ctypes.c_double(0.0)
    lib.vls_prof_collect(0, ctypes.byref(cnt), ctypes.byref(tot))
    cnt_s, tot_s = ctypes.c_int(0), ctypes.c_double(0.0)
    lib.vls_prof_collect(1, ctypes.byref(cnt_s), ctypes.byref(tot_s))
    peak, peak_source = pick_tensor_peak(clocks)
    flops_alg = 4.0 * B * NQ * NK_STEADY * D
    flops_exec = 2.0 * B * NQ * NK_STEADY * (D + DV)
    avg_ms = tot.value / max(cnt.value, 1)
    achieved = flops_alg / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
    executed = flops_exec / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "attn_cross_dram_bytes.json")
    if os.path.exists(prof) and B == 1:
        traffic = json.load(open(prof)).get("dram_bytes_per_launch")
    return {"bound": "tensor", "kernel": "attn_x2_kernel + combine (memory cross-attention, two query tiles per CTA, Nq=4096, Nk=28736, qk dim 256, value dim 64)",
            "achieved": round(executed, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(executed / peak, 4),
            "algorithmic": round(achieved, 2), "frac_algorithmic": round(achieved / peak, 4),
            "traffic": traffic, "peak_source": peak_source, "launches_timed": cnt.value,
            "avg_launch_ms": round(avg_ms, 4), "flops_per_launch": flops_alg, "executed_flops_per_launch": flops_exec,
            "self_attn_avg_launch_ms": round(tot_s.value / max(cnt_s.value, 1), 4),
            "note": "achieved / frac = FLOPs the tensor cores actually perform (2*Nq*Nk*(256+64)) / launch time (kernel + combine) "
                    "/ peak; algorithmic = the reference's 4*Nq*Nk*256 per launch / the same time: it exceeds the executed "
                    "rate (and can exceed the peak) because softmax(QK^T) is applied to the 64-d memory, not to its 256-d "
                    "projection"}


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def _oracle_steady_state(num_frames_total, objects=1):
    """A full memory bank for the torch path without tracking 16 frames on it: 7 memories from the oracle's
    own memory encoder on synthetic masks + 16 seeded pointers (same shapes/dtypes the predictor would hold)."""
    import torch

    from oracle import sam2_path as O
    from video_llava_seg_b200 import synth

    sd = synth.init_state_dict(0)
    clip = make_clip(100, min(num_frames_total, PERIOD))
    g = torch.Generator().manual_seed(9)
    out = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
    pos = O.sine_pe_2d(64, 64, 64)[None]
    for t in range(RAMP):
        e = dict(obj_ptr=torch.randn(objects, 256, generator=g) * 0.5, maskmem_features=None, maskmem_pos_enc=[pos])
        if t == 0 or t >= RAMP - 6:
            f = clip.frame(t % PERIOD, objects)
            mask = torch.sigmoid(torch.randn(objects, 1, 1024, 1024, generator=g)) * 20 - 10
            pix = f["vision_feat"].permute(1, 2, 0).reshape(objects, 256, 64, 64)
            e["maskmem_features"] = O.memory_encoder(sd, pix, mask, True)["vision_features"].to(torch.bfloat16)
        (out["cond_frame_outputs"] if t == 0 else out["non_cond_frame_outputs"])[t] = e
    return sd, clip, out


def cpu_steps(steps, warmup=0, objects=1):
    import torch

    from oracle import cc as cc_oracle
    from oracle import sam2_path as O

    torch.set_num_threads(os.cpu_count() or 1)
    T = RAMP + warmup + steps + 1
    sd, clip, bank = _oracle_steady_state(T, objects)
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            t = RAMP + i
            feats = clip.frame(t % PERIOD, objects)
            t0 = time.perf_counter()
            o = O.track_step(sd, O.Cfg, t, False, feats, None, bank, T, run_mem_encoder=True)
            pm = O.fill_holes_in_mask_scores(o["pred_masks"], O.Cfg.fill_hole_area, cc_oracle.cc_label)
            dt = time.perf_counter() - t0
            bank["non_cond_frame_outputs"][t] = dict(
                maskmem_features=o["maskmem_features"].to(torch.bfloat16), maskmem_pos_enc=o["maskmem_pos_enc"],
                pred_masks=pm, obj_ptr=o["obj_ptr"], object_score_logits=o["object_score_logits"])
            if i >= warmup:
                times.append(dt)
    return times, torch.get_num_threads()


def cpu_baseline(steps, objects=1):
    times, cores = cpu_steps(steps, objects=objects)
    return {"value": round(len(times) / sum(times), 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} steady-state frames (Nk=28736, {objects} object(s)) of oracle/sam2_path.py track_step + hole "
                      f"filling, torch CPU fp32, {cores} threads"}


def gpu_eager_baseline(dev, objects=1, steps=8, warmup=3):
    """INFORMATIONAL (r1 verdict): what stock PyTorch gives on this same B200 -- the oracle port's torch operators
    (the restated reference modules: nn.functional linear / SDPA / conv / layer_norm / interpolate) run eagerly on
    cuda:0 under bf16 autocast, as the reference runs its SAM2 (llava/inference: model.to(bfloat16)).  Hole filling is
    left out (the reference's CC kernel has no build recipe; the oracle's is CPU code).  Not a target, not a parity
    claim: it puts the GPU-over-CPU ratio next to a GPU-over-GPU one."""
    import torch

    try:
        from oracle import sam2_path as O

        T = RAMP + warmup + steps + 1
        with torch.inference_mode():
            sd, clip, bank = _oracle_steady_state(T, objects)          # built on the CPU (seeded CPU generators), then moved
        to_dev = lambda x: x.to(dev) if torch.is_tensor(x) else [y.to(dev) for y in x] if isinstance(x, list) else x
        sd = {k: v.to(dev) for k, v in sd.items()}
        for part in bank.values():
            for t, e in part.items():
                part[t] = {k: to_dev(v) for k, v in e.items()}
        feats_all = [{k: v.to(dev) for k, v in clip.frame(t, objects).items()} for t in range(4)]
        with torch.inference_mode(), torch.device(dev), torch.autocast("cuda", dtype=torch.bfloat16):
            evs = []
            for i in range(warmup + steps):
                t = RAMP + i
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                o = O.track_step(sd, O.Cfg, t, False, feats_all[i % 4], None, bank, T, run_mem_encoder=True)
                b.record()
                bank["non_cond_frame_outputs"][t] = dict(
                    maskmem_features=o["maskmem_features"].to(torch.bfloat16), maskmem_pos_enc=o["maskmem_pos_enc"],
                    pred_masks=o["pred_masks"], obj_ptr=o["obj_ptr"].float(), object_score_logits=o["object_score_logits"])
                if i >= warmup:
                    evs.append((a, b))
            torch.cuda.synchronize()
        ms = statistics.median(a.elapsed_time(b) for a, b in evs)
        return {"value": round(1e3 / ms, 2), "unit": UNIT, "ms_per_step": round(ms, 3), "kind": "oracle port, torch eager on cuda, bf16 autocast",
                "sample": f"{len(evs)} steady-state frames, {objects} object(s), no hole filling, no CUDA graph", "informational": True}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {str(e)[:160]}", "informational": True}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    K, W = args.steps, max(args.warmup, 0)
    if wl["kind"] == "clips":       # bounded sample of the clip workloads: steady-state frames of one clip
        K = min(K, 24)
    times, cores = cpu_steps(K, W, objects=wl["objects"])
    total = sum(times)
    value = len(times) / total
    sample = (f"{len(times)} steady-state frames ({wl['objects']} object(s), Nk=28736) of oracle/sam2_path.py track_step + hole "
              f"filling after {W} warm-up frames, torch CPU fp32, {cores} threads"
              + ("; the clip workloads are sampled by steady-state frames of one clip" if wl["kind"] == "clips" else ""))
    line = {
        "impl": "reference", "metric": wl["metric"], "value": round(value, 4), "unit": UNIT, "n_gpus": world,
        "steps": len(times), "warmup": W, "ms_per_step": round(total / len(times) * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic backbone features + seeded random-init weights",
        "config": make_config(args.workload, world),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference path on the host CPU: oracle port of the unmodified PyTorch modules (pinned to the reference by "
                "tests/golden/make_golden.py; the Python reference itself cannot travel to this box)",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--repeats", type=int, default=0, help="timed windows of --steps steps (0: enough for >= 200 steps)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="configs[1]", choices=sorted(WORKLOADS))
    ap.add_argument("--sweep-clips", type=int, default=512, help="configs[4]: clips of the whole job")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pixels", action="store_true", help="skip the informational e2e_from_pixels leg (image encoder + hot path)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
