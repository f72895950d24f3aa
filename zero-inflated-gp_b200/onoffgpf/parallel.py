"""Data-parallel plumbing: contiguous row shards, one all-reduce of the packed accumulator per evaluation.
Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests): nothing here touches the kernels."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """rank r owns rows [r*N/P, (r+1)*N/P): contiguous, sizes differ by at most one, every row owned exactly once"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return rank * n_total // world, (rank + 1) * n_total // world


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def global_count(n_local: int, device=None, group=None) -> int:
    if not is_distributed():
        return int(n_local)
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, group=group)
    return int(t.item())


def allreduce_sum_(buf: torch.Tensor, group=None) -> torch.Tensor:
    """in-place sum over ranks of the packed per-shard accumulator (no-op on one rank)"""
    if is_distributed():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def evaluate_sharded(accumulate: Callable[[], torch.Tensor], finish: Callable[[torch.Tensor], object], group=None):
    """accumulate() -> this rank's packed partial sums; they are summed over ranks ONCE; finish(reduced) runs on every rank
    (it holds everything that must be counted exactly once: the KL term and the O(M^3) backward)."""
    buf = accumulate()
    allreduce_sum_(buf, group)
    return finish(buf)
