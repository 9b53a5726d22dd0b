"""Data-parallel plumbing over torch.distributed (one process per GPU, NCCL over NVLink;
gloo in the CPU tests).  Cells shard across ranks as contiguous row ranges; parameters are
replicated; the only per-step exchange is a sum all-reduce of the flat gradient buffer
(SURVEY.md 8e).  The reference has no distributed code at all."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world: int, equal: bool = False) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of rank ``rank``.  equal=True gives every rank exactly
    n // world rows (the last n % world rows are left out) so that all ranks run the same
    number of steps with the same batch sizes; equal=False spreads the remainder."""
    if world <= 1:
        return 0, n
    if equal:
        per = n // world
        return rank * per, (rank + 1) * per
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_sum_host(a: np.ndarray, device) -> np.ndarray:
    """Sum a small host float64 vector over ranks (epoch-level scalars)."""
    if not is_dist():
        return a
    backend = dist.get_backend()
    t = torch.as_tensor(a, dtype=torch.float64)
    if backend == "nccl":
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def all_reduce_max_host(v: float, device) -> float:
    if not is_dist():
        return v
    t = torch.tensor([v], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if is_dist() and t.numel():
        dist.broadcast(t, src=src)
    return t
