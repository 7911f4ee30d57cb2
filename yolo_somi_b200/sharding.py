"""Batch sharding of the DCNv3 hot path across ranks (one process per GPU).

The op is embarrassingly parallel over the batch: forward and backward of image n touch only
image n (reference: the kernels index the batch only to offset pointers,
models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:238,247,315-317), so ranks take disjoint image
ranges and there is NO data-path collective.  The only cross-rank traffic of a training step is
DDP's gradient all-reduce of the surrounding layers' weights (train.py:208,270), which is outside
the op.  These helpers are what bench.py uses for N > 1 and what the gloo tests exercise.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of the images owned by `rank`; remainders go to the lowest ranks
    (the reference drops to batch_size // WORLD_SIZE per rank, train.py:170)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """Slowest rank's time: multi-GPU numbers are the max over ranks, never a mean."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_throughput(units_this_rank: float, seconds_this_rank: float,
                         device: torch.device | str = "cpu") -> float:
    """units processed by all ranks / slowest rank's time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)
