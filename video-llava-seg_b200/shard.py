"""Clip sharding across GPUs: the path is embarrassingly parallel over videos (and objects), exactly
as the reference's inference driver splits its video list per process (llava/inference/main.py:41-49,
scripts/infer.sh:1-7).  No collective is involved; every rank owns a contiguous chunk."""
import math


def shard_clips(clips, world_size, rank):
    """Contiguous chunks of ceil(n / world_size) like the reference's split_list/get_chunk."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    clips = list(clips)
    if not clips:
        return []
    chunk = math.ceil(len(clips) / world_size)
    return clips[rank * chunk:(rank + 1) * chunk]


def aggregate_throughput(units_local, ms_local, device=None):
    """Whole-job throughput of independent replicas: all units processed by all ranks divided by the
    slowest rank's device time (max over ranks).  Uses torch.distributed only for that reduction --
    there is no collective on the data path.  Works with NCCL (GPU tensors) and gloo (CPU)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return units_local / (ms_local / 1e3), ms_local, units_local
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return u.item() / (t.item() / 1e3), t.item(), u.item()
