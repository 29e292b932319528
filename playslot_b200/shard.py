"""
Multi-GPU plumbing of the slot pipeline: videos are independent, so a batch is split contiguously over the
ranks (one process per GPU) and nothing is exchanged on the data path; timing is reduced with MAX over ranks.
"""
import os

import torch
import torch.distributed as dist


def world():
    """ (rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU) """
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n, rank, world_size):
    """ contiguous [start, stop) of `n` videos owned by `rank`; sizes differ by at most one """
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(videos, rank, world_size):
    """ the slice of a (B, ...) batch this rank processes (a view, no copy) """
    a, b = shard_range(videos.shape[0], rank, world_size)
    return videos[a:b]


def max_over_ranks(value, device=None):
    """ MAX-reduce a python float over all ranks (identity when not distributed) """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_concat(local, device=None):
    """ all-gather per-rank result tensors (possibly of different length along dim 0) and concatenate in rank
    order -- used by tests to check that N ranks reproduce the single-process result; not on the timed path """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    ws = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.zeros_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    return torch.cat([o[:int(s.item())] for o, s in zip(out, sizes)], dim=0)
