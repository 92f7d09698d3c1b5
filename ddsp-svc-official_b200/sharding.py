"""Clip sharding across GPUs: clips (batch rows) are independent on this path,
so rank r simply takes clips r, r+world, ... and no collective touches the data.  The only
communication is the benchmark's timing reduction."""
import torch
import torch.distributed as dist


def shard_clips(n_clips, rank, world):
    """Indices of the clips rank `rank` synthesises (round-robin, as SURVEY §8e: clip i -> GPU i mod G)."""
    return range(rank, n_clips, world)


def reduce_timing(elapsed_ms, samples, device):
    """(max elapsed over ranks, total samples over ranks); identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), int(samples)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    n = torch.tensor([int(samples)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())
