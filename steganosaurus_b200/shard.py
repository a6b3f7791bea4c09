"""Sharding of the image batch across ranks (one process per GPU, no data-path collective).

The path shards by image: rank r of R owns a contiguous block of the batch.  The
only collective traffic is control-plane: a barrier around the timed region and a max-reduce of
the per-rank times.  Backend-agnostic (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the items rank `rank` owns; blocks differ by at most one item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(item: int, n_items: int, world: int) -> int:
    base, rem = divmod(n_items, world)
    edge = rem * (base + 1)
    if item < edge:
        return item // (base + 1)
    return rem + (item - edge) // base if base else world - 1


class Timing:
    """Barrier + max-over-ranks helpers around an (optional) torch.distributed process group."""

    def __init__(self, dist=None, device=None):
        self.dist, self.device = dist, device

    @property
    def world(self) -> int:
        return self.dist.get_world_size() if self.dist is not None else 1

    @property
    def rank(self) -> int:
        return self.dist.get_rank() if self.dist is not None else 0

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def aggregate_throughput(units_per_rank: float, seconds_this_rank: float, timing: Timing) -> float:
    """Whole-job throughput = units all ranks processed / max-over-ranks time."""
    total = timing.sum_over_ranks(units_per_rank)
    t = timing.max_over_ranks(seconds_this_rank)
    return total / t if t > 0 else 0.0
