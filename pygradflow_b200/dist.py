"""Batch sharding across the GPUs of one box.

Instances are independent -- the reference itself parallelises only over instances
(pygradflow/runners/runner.py:107-153) -- so rank r owns the contiguous block
[r*B/G, (r+1)*B/G) of the batch, runs the whole solve locally with no collective on the hot path, and
the converged iterates / status words / iteration counts are gathered ONCE at the end
(NCCL all-gather over NVLink; gloo in the CPU tests).  One process per GPU (torchrun).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of rank `rank`; block sizes differ by at most one, earlier ranks get the extras."""
    assert 0 <= rank < world and B >= 0
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_rows(local: torch.Tensor, B: int) -> torch.Tensor:
    """All-gather a row-sharded tensor ([b_local, ...] on every rank, shard_range layout) into [B, ...]."""
    rank, world = _world()
    if world == 1:
        return local
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    maxn = max(sizes)
    tail = tuple(local.shape[1:])
    padded = torch.zeros((maxn,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * maxn,) + tail, dtype=local.dtype, device=local.device)
    try:
        dist.all_gather_into_tensor(out, padded.contiguous())
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded.contiguous())
        out = torch.cat(parts, dim=0)
    chunks = [out[r * maxn : r * maxn + sizes[r]] for r in range(world)]
    return torch.cat(chunks, dim=0)


@dataclass
class ShardedResult:
    x: torch.Tensor
    y: torch.Tensor
    status: torch.Tensor
    iterations: torch.Tensor
    accepted_steps: torch.Tensor
    local_range: Tuple[int, int]


def gather_result(local: Dict[str, torch.Tensor], B: int, local_range: Tuple[int, int]) -> ShardedResult:
    """The single collective of a sharded solve: every rank ends up with the full (x, y, status, counts)."""
    g = {k: gather_rows(v, B) for k, v in local.items()}
    return ShardedResult(g["x"], g["y"], g["status"], g["iterations"], g["accepted_steps"], local_range)


def solve_sharded(B: int, problem_factory: Callable[[int, int], "object"], params=None,
                  x0: Optional[torch.Tensor] = None, y0: Optional[torch.Tensor] = None) -> ShardedResult:
    """Solve a batch of B instances over all ranks.

    problem_factory(lo, hi) builds the BatchedProblem of instances [lo, hi) on this rank's GPU; x0 / y0 are
    full-batch starting points (or None)."""
    from .solver import BatchedSolver

    rank, world = _world()
    lo, hi = shard_range(B, rank, world)
    problem = problem_factory(lo, hi)
    solver = BatchedSolver(problem, params)
    res = solver.solve(None if x0 is None else x0[lo:hi], None if y0 is None else y0[lo:hi])
    local = dict(x=res.x, y=res.y, status=res.status, iterations=res.iterations, accepted_steps=res.accepted_steps)
    return gather_result(local, B, (lo, hi))
