"""world_size-2 `gloo` test of the N>1 host logic (no GPU): clip sharding is disjoint and complete,
ranks run independently (no data-path collective), and the whole-job throughput is
sum(units) / max-over-ranks(time) -- the same reduction bench.py performs over NCCL."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from video_llava_seg_b200.shard import aggregate_throughput, shard_clips

    clips = list(range(7))
    mine = shard_clips(clips, world, rank)
    frames = sum(10 + c for c in mine)            # pretend every clip c has 10 + c frames
    ms = 100.0 * (rank + 1)                       # rank 1 is the slow one
    dist.barrier()
    value, ms_max, units = aggregate_throughput(frames, ms)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        torch.save(dict(value=value, ms_max=ms_max, units=units, shards=gathered), os.path.join(out_dir, "r0.pt"))
    dist.destroy_process_group()


def test_two_rank_sharding_and_aggregation(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = torch.load(os.path.join(tmp_path, "r0.pt"))
    assert r["shards"] == [[0, 1, 2, 3], [4, 5, 6]]
    assert r["units"] == sum(10 + c for c in range(7))
    assert r["ms_max"] == 200.0
    assert abs(r["value"] - r["units"] / 0.2) < 1e-9
