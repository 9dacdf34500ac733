"""Multi-GPU plumbing of the hot path: scenes (tiles) are independent in forward, so the N-GPU path is one
process per GPU with disjoint tile shards and NO data-path collective (`SURVEY.md` §8e; the reference's `pmap`
splits the batch the same way, `snap/trainer.py:452-464`).  torch.distributed is used for the barrier and for
the max-over-ranks timing only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_tiles(tile_ids: Sequence[int], rank: int, world_size: int) -> List[int]:
    """Contiguous, balanced shard of the global tile list for this rank (sizes differ by at most one)."""
    n = len(tile_ids)
    lo = rank * n // world_size
    hi = (rank + 1) * n // world_size
    return list(tile_ids[lo:hi])


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """Job time = slowest rank (all-reduce MAX); identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(units_this_rank: int, seconds_this_rank: float, device: torch.device | str = "cpu") -> float:
    """Whole-job units/s: all units of all ranks divided by the max-over-ranks time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)
