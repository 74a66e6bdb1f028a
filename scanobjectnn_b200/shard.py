"""Batch sharding across ranks (one process per GPU).  The hot path is per-cloud, so ranks take disjoint,
contiguous slices of the global batch and never exchange data; torch.distributed is used only for the
barrier and for reducing timings / counters (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of ``range(n_items)`` owned by ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank seed for synthetic data so that ranks hold different clouds (weak scaling)."""
    return seed + 7919 * rank


def max_over_ranks(value: float, device=None) -> float:
    """MAX-reduce a scalar (device-time in ms) over all ranks; identity without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
