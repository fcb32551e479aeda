#!/usr/bin/env python
"""bench.py -- propagated frames/s of the SAM 2.1 mask-propagation hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W     # CPU arm: oracle port of the reference path

Workload = BASELINE.json configs[1]: Hiera-B+ propagation at 1024^2 (64x64 tokens), 7-frame memory bank +
16 object pointers (Nk = 28 736), 1 object per GPU, synthetic clip, random-init weights.  The image encoder is
outside the hot path: clips are given as backbone features.  One step = one propagated frame in steady state
(the bank is filled during an untimed 17-frame ramp).  N > 1: one process per GPU (torchrun), one clip per rank,
no collective on the data path ("scaling": "weak"); NCCL only for the barrier / max-over-ranks of the time.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")

METRIC = "propagated frames/sec (SAM2.1 Hiera-B+ hot path, 1024^2, 7-frame memory bank, 1 object per GPU)"
UNIT = "frames/s"
WORKLOAD = "configs[1]: SAM2.1 Hiera-B+ propagation, 64x64 tokens, 7 memories + 16 pointers (Nk=28736), 1 object, synthetic"
RAMP = 17  # prompt frame + 16 propagated frames: full memory bank and 16 pointers afterwards
NQ, NK_STEADY, D = 4096, 7 * 4096 + 64, 256


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(tflops=float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1400.0))), hbm=float(j["hbm_gbs"]),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


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
    from video_llava_seg_b200.shard import aggregate_throughput

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
    sampler = ClockSampler(local).start()
    K, W = args.steps, max(args.warmup, 3)
    T = max(RAMP + W + K + 1, RAMP + 9)  # the roofline pass needs RAMP + 8 frames
    predictor = build_sam.build_sam2_video_predictor(None, synth.init_state_dict(0), dev)
    clip = make_clip(100 + rank, T)
    frames = [clip.frame(t, 1) for t in range(T)]
    prompt = clip.point_prompt(1)["point_coords"][0].tolist()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def timed_pass(source, d2h):
        """Ramp + warm-up untimed, then K steps each bracketed by CUDA events, L2 flushed between steps.
        With d2h the binarised video-resolution mask of EVERY step is read back into pinned host memory inside the
        step's events; the consumer is software-pipelined by one frame (it waits for frame t-1's mask after frame t
        has been enqueued), as a streaming client of propagate_in_video would be."""
        import gc

        host = [torch.empty((1, 1, 1024, 1024), dtype=torch.uint8).pin_memory() for _ in range(2)] if d2h else None
        done = [torch.cuda.Event(), torch.cuda.Event()]
        checksum = 0

        def read_back(m, slot):   # with d2h the predictor runs in output_mode "binary": m is already the uint8 mask
            host[slot].copy_(m, non_blocking=True)
            done[slot].record()

        predictor.output_mode = "binary" if d2h else "logits"

        with sampler as clocks:
            state = predictor.init_state(source)
            predictor.add_new_points_or_box(state, 0, 1, points=prompt, labels=[1])
            gen = predictor.propagate_in_video(state)
            for j in range(RAMP + W):
                _, _, m = next(gen)
                if d2h:
                    read_back(m, j % 2)
                    done[j % 2].synchronize()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
            stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
            launches0 = lib.vls_launch_count()
            out_bytes = 0
            gc.collect()
            gc.disable()          # a generation-2 collection inside a 2 ms step is host noise, not the path
            clocks.mark_start()
            dbg = bool(os.environ.get("VLS_BENCH_DEBUG"))
            host_ms, seg0 = [], torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
            for i in range(K):
                t_host = time.perf_counter()
                flush.zero_()
                starts[i].record()
                _, _, m = next(gen)
                if d2h:
                    read_back(m, i % 2)                      # result of step i -> pinned host memory, inside its events
                    out_bytes = host[i % 2].numel()
                stops[i].record()
                if d2h and i > 0:
                    done[(i - 1) % 2].synchronize()          # consume step i-1's mask on the host
                    checksum += int(host[(i - 1) % 2][0, 0, 0, 0])
                if dbg:
                    host_ms.append(round((time.perf_counter() - t_host) * 1e3, 3))
            torch.cuda.synchronize()
            clocks.mark_stop()
            gc.enable()
            predictor.output_mode = "logits"
        launches = lib.vls_launch_count() - launches0
        per_step = [s.elapsed_time(e) for s, e in zip(starts, stops)]
        ms = sum(per_step)
        if dbg:
            print(f"[bench debug] d2h={d2h} per-step ms: {[round(x, 3) for x in per_step]}", file=sys.stderr, flush=True)
            print(f"[bench debug] d2h={d2h} host ms per iteration: {host_ms}; cudaMalloc segments during the timed loop: "
                  f"{torch.cuda.memory_stats(dev).get('num_device_alloc', 0) - seg0}", file=sys.stderr, flush=True)
        if world > 1:
            dist.barrier()
        fps, ms, _ = aggregate_throughput(K, ms, dev)   # sum of frames over ranks / max-over-ranks device time
        gen.close()
        return fps, ms, launches, clocks.summary(), out_bytes

    # (1) device-resident inputs: kernel + host-orchestration throughput
    resident = FeatureClip(lambda t: frames[t], T, resident_device=dev)
    # one untimed pass over the whole clip first: the predictor keeps every frame's outputs (as the reference
    # does), so a fresh process would otherwise time cudaMalloc growth of the caching allocator, not the path
    warm = predictor.init_state(resident)
    predictor.add_new_points_or_box(warm, 0, 1, points=prompt, labels=[1])
    for _ in predictor.propagate_in_video(warm):
        pass
    del warm
    torch.cuda.synchronize()
    value, ms, launches, clocks, _ = timed_pass(resident, d2h=False)
    # (2) end to end through the public API with host buffers: H2D of each frame's features, D2H of the mask
    pinned = FeatureClip(lambda t: frames[t], T, pinned=True)
    e2e, _, _, _, out_bytes = timed_pass(pinned, d2h=True)
    sampler.stop()
    # (3) informational: whole 64-frame clips (configs[1] as a user runs it: prompt frame, 16 frames with a growing
    # bank, then steady state; captured graph re-used from clip to clip), wall clock around complete sessions
    def whole_clips(n_clips=3, T_clip=64):
        src = FeatureClip(lambda t: frames[t % T], T_clip, resident_device=dev)
        predictor.output_mode = "logits"
        times = []
        for _ in range(n_clips + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = predictor.init_state(src)
            predictor.add_new_points_or_box(st, 0, 1, points=prompt, labels=[1])
            for _ in predictor.propagate_in_video(st):
                pass
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        best = min(times[1:])
        return {"frames_per_clip": T_clip, "value": round(T_clip / best, 1), "unit": UNIT, "ms_per_clip": round(best * 1e3, 2),
                "note": "complete sessions incl. prompt frame and 16-frame ramp, resident features, best of %d" % n_clips}

    clip_info = whole_clips()
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic backbone features + seeded random-init weights",
        "config": {"workload": WORKLOAD, "objects_per_gpu": 1, "clips_per_gpu": 1, "ramp_frames": RAMP,
                   "l2": "flushed between timed steps (256 MiB memset, outside the per-step events)",
                   "parallelism": f"{world} independent replica(s), sharded by clip, no collective on the path",
                   "execution": "steady-state frames replay one CUDA graph (graphed.py); gpu_launches counts the library kernels inside each replay",
                   "e2e_path": "pinned host features -> double-buffered H2D -> propagate_in_video(output_mode='binary': fused resize+threshold) -> uint8 mask D2H into pinned memory every step, consumer pipelined by one frame"},
        "clocks": clocks, "gpu_launches": int(launches), "whole_clip": clip_info,
        "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": int(pinned.h2d_bytes_per_frame),
                "d2h_bytes_per_step": int(out_bytes)},
    }
    if rank == 0:
        line["roofline"] = roofline(predictor, resident, prompt, lib, torch)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(steps=2)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def roofline(predictor, source, prompt, lib, torch):
    """Dominant kernel = attn_fwd_kernel on the memory cross-attention (4 launches / frame).  Algorithmic FLOPs per
    launch = 4 * Nq * Nk * d (QK^T + PV, 2 flops/MAC); duration = CUDA events around each launch on its stream."""
    import ctypes

    predictor.use_cuda_graph = False   # the same kernels, launched eagerly so that each launch can be bracketed by events
    state = predictor.init_state(source)
    predictor.add_new_points_or_box(state, 0, 1, points=prompt, labels=[1])
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
    predictor.use_cuda_graph = True
    cnt, tot = ctypes.c_int(0), ctypes.c_double(0.0)
    lib.vls_prof_collect(0, ctypes.byref(cnt), ctypes.byref(tot))
    cnt_s, tot_s = ctypes.c_int(0), ctypes.c_double(0.0)
    lib.vls_prof_collect(1, ctypes.byref(cnt_s), ctypes.byref(tot_s))
    pk = peaks()
    flops = 4.0 * NQ * NK_STEADY * D
    avg_ms = tot.value / max(cnt.value, 1)
    achieved = flops / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "attn_cross_dram_bytes.json")
    if os.path.exists(prof):
        traffic = json.load(open(prof)).get("dram_bytes_per_launch")
    return {"bound": "tensor", "kernel": "attn_fwd_kernel (memory cross-attention, Nq=4096, Nk=28736, d=256)",
            "achieved": round(achieved, 2), "peak": pk["tflops"], "unit": "TFLOP/s", "frac": round(achieved / pk["tflops"], 4),
            "traffic": traffic, "peak_source": pk["source"], "launches_timed": cnt.value,
            "avg_launch_ms": round(avg_ms, 4), "flops_per_launch": flops,
            "self_attn_avg_launch_ms": round(tot_s.value / max(cnt_s.value, 1), 4)}


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def _oracle_steady_state(num_frames_total):
    """A full memory bank for the CPU path without tracking 16 frames on the CPU: 7 memories from the oracle's
    own memory encoder on synthetic masks + 16 seeded pointers (same shapes/dtypes the predictor would hold)."""
    import torch

    from oracle import sam2_path as O
    from video_llava_seg_b200 import synth

    sd = synth.init_state_dict(0)
    clip = make_clip(100, num_frames_total)
    g = torch.Generator().manual_seed(9)
    out = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
    pos = O.sine_pe_2d(64, 64, 64)[None]
    for t in range(RAMP):
        e = dict(obj_ptr=torch.randn(1, 256, generator=g) * 0.5, maskmem_features=None, maskmem_pos_enc=[pos])
        if t == 0 or t >= RAMP - 6:
            f = clip.frame(t, 1)
            mask = torch.sigmoid(torch.randn(1, 1, 1024, 1024, generator=g)) * 20 - 10
            pix = f["vision_feat"].permute(1, 2, 0).reshape(1, 256, 64, 64)
            e["maskmem_features"] = O.memory_encoder(sd, pix, mask, True)["vision_features"].to(torch.bfloat16)
        (out["cond_frame_outputs"] if t == 0 else out["non_cond_frame_outputs"])[t] = e
    return sd, clip, out


def cpu_steps(steps, warmup=0):
    import torch

    from oracle import cc as cc_oracle
    from oracle import sam2_path as O

    torch.set_num_threads(os.cpu_count() or 1)
    T = RAMP + warmup + steps + 1
    sd, clip, bank = _oracle_steady_state(T)
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            t = RAMP + i
            feats = clip.frame(t, 1)
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


def cpu_baseline(steps):
    times, cores = cpu_steps(steps)
    return {"value": round(len(times) / sum(times), 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} steady-state frames (Nk=28736, 1 object) of oracle/sam2_path.py track_step + hole "
                      f"filling, torch CPU fp32, {cores} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    times, cores = cpu_steps(K, min(W, 1))
    total = sum(times)
    value = len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": len(times), "warmup": min(W, 1), "ms_per_step": round(total / len(times) * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic backbone features + seeded random-init weights",
        "config": {"workload": WORKLOAD, "note": "reference path on the host CPU (oracle port of the unmodified PyTorch "
                   "modules; the Python reference itself cannot travel to this box); each step = one steady-state frame"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} steady-state frames, torch CPU fp32, {cores} threads"},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
