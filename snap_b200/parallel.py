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


# ---- data-parallel gradient mean ----------------------------------------------------------

def _leaves(tree, prefix=()):
    if isinstance(tree, dict):
        for k in sorted(tree):
            yield from _leaves(tree[k], prefix + (k,))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            yield from _leaves(v, prefix + (i,))
    elif tree is not None:
        yield prefix, tree


def pmean_tree(tree, bucket_bytes: int = 256 << 20, group=None):
    """In-place mean over ranks of every tensor of a (nested dict / list) gradient tree — the counterpart of
    `jax.lax.pmean(grad, axis_name='batch')` (`snap/trainer.py:231-234`).

    Leaves are packed, in sorted key order (identical on every rank), into flat fp32 buckets of at most
    `bucket_bytes` and each bucket is reduced with ONE all-reduce: the ≈48 M-parameter tree of the localisation
    model (≈193 MB fp32) goes out in a single NCCL call instead of ≈330 per-leaf calls, so the cost is NVLink
    bandwidth rather than launch latency.  Accumulation is fp32 whatever the leaf dtype; results are written back
    into the leaves.  Identity without a process group.  Returns the number of all-reduce calls issued."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    leaves = [t for _, t in _leaves(tree)]
    calls, i = 0, 0
    while i < len(leaves):
        j, nbytes = i, 0
        while j < len(leaves) and (j == i or nbytes + leaves[j].numel() * 4 <= bucket_bytes):
            nbytes += leaves[j].numel() * 4
            j += 1
        chunk = leaves[i:j]
        flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in chunk])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        o = 0
        for t in chunk:
            n = t.numel()
            t.detach().copy_(flat[o:o + n].view(t.shape))
            o += n
        calls += 1
        i = j
    return calls


class GradBucket:
    """Persistent flat fp32 gradient bucket: the counterpart of `jax.lax.pmean(grad, axis_name='batch')` on the ~48 M
    parameter tree of the localisation model (`snap/trainer.py:231-234`, ~193 MB fp32 per step) without the three extra
    passes of `pmean_tree` (torch.cat -> all-reduce -> divide -> copy back).

    The bucket owns ONE flat fp32 device buffer; `views[i]` is the gradient array of leaf i inside it (16-byte aligned), so
    the backward kernels write gradients straight into the bucket and `allreduce_mean` is a single in-place all-reduce
    (NCCL: ReduceOp.AVG, no separate division; gloo: SUM + one in-place scale).  Nothing is allocated per call, so the
    collective can be captured into a CUDA graph, or enqueued on a side stream (`stream=`) to overlap the tail of the
    backward pass: the caller's stream waits for it with `wait()`.  Identity without a process group."""

    ALIGN = 4   # elements (16 bytes)

    def __init__(self, shapes, device, group=None):
        self.shapes = [tuple(int(d) for d in sh) for sh in shapes]
        self.group = group
        offs, o = [], 0
        for sh in self.shapes:
            n = 1
            for d in sh:
                n *= d
            offs.append((o, n))
            o += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(max(o, self.ALIGN), dtype=torch.float32, device=device)
        self.views = [self.flat[a:a + n].view(sh) for (a, n), sh in zip(offs, self.shapes)]
        self._event = None

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero_(self) -> None:
        self.flat.zero_()

    def allreduce_mean(self, stream=None) -> int:
        """In-place mean over ranks of the whole bucket; returns the number of collectives issued (0 or 1)."""
        if not (dist.is_available() and dist.is_initialized()):
            return 0
        world = dist.get_world_size(self.group)
        if world == 1:
            return 0

        def run():
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.mul_(1.0 / world)
        if stream is None or not self.flat.is_cuda:
            run()
        else:
            stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(stream):
                run()
                self._event = torch.cuda.Event()
                self._event.record(stream)
        return 1

    def wait(self) -> None:
        """Make the current stream wait for a collective enqueued on a side stream."""
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
            self._event = None

    def all_finite(self) -> torch.Tensor:
        """Device-side 0-dim bool: every gradient is finite (`trainer.py:260-276` skips the update otherwise)."""
        return torch.isfinite(self.flat).all()
