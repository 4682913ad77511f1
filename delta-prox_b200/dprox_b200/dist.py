"""Multi-GPU: a batch of independent problems shards over the batch axis, one process per GPU
(SURVEY §8e).  There is NO collective in the iteration; the only exchange is the optional residual
all-reduce of the stopping rule (`algo.ResidualStop`, 4 floats) and the final gather of results.

The helpers work with any initialised torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_range(n: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous, balanced split of `n` problems: the first n % world ranks get one extra."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank=None, world_size=None) -> torch.Tensor:
    lo, hi = shard_range(t.shape[0], rank, world_size)
    return t[lo:hi]


def shard_schedule(s, n: int, rank=None, world_size=None):
    """Per-sample schedules [B,T] follow their problems; shared schedules ([T] / scalars) are replicated."""
    if isinstance(s, dict):
        return {k: shard_schedule(v, n, rank, world_size) for k, v in s.items()}
    if isinstance(s, torch.Tensor) and s.ndim == 2 and s.shape[0] == n:
        return shard_batch(s, rank, world_size)
    return s


def gather_batch(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank result shards back into the full batch (uneven shards are padded)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    w = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, w) for r in range(w)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)


def solve_sharded(make_solver, x0: torch.Tensor, rhos=None, lams=None, gather: bool = True, **solve_kw):
    """Solve a batch of independent problems across the ranks of the default process group.

    make_solver(lo, hi) -> (solver, lams_for_that_solver): builds the solver for problems [lo, hi) on this rank
    (the data term usually carries per-problem measurements, so it has to be built per shard).
    """
    n = x0.shape[0]
    lo, hi = shard_range(n)
    solver, local_lams = make_solver(lo, hi)
    lams = local_lams if local_lams is not None else shard_schedule(lams, n)
    out = solver.solve(x0=x0[lo:hi], rhos=shard_schedule(rhos, n), lams=lams, **solve_kw)
    return gather_batch(out, n) if gather else out


def allreduce_gradients(params, group=None, bucket_bytes: int = 32 << 20, average: bool = True) -> int:
    """Data-parallel training of an unrolled solver (BASELINE config 5): every rank differentiates its own shard of the
    batch; the gradients of the shared trainable parameters (rho / sigma schedules, the DOE height map, ...) are summed
    over ranks in flat buckets — one all-reduce per `bucket_bytes` (NCCL over NVLink on the GPU box, gloo in the CPU
    tests).  The reference has no distributed training at all; semantics follow DDP: parameters without a
    gradient contribute zeros, the result is averaged.  Returns the number of collectives issued."""
    params = [p for p in params if p.requires_grad]
    if not params or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    w = dist.get_world_size(group)
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= w
        off = 0
        for p in bucket:
            n = p.numel()
            gnew = flat[off:off + n].reshape(p.shape).to(p.dtype)
            if p.grad is None:
                p.grad = gnew.clone()
            else:
                p.grad.copy_(gnew)
            off += n
        calls += 1
        bucket, size = [], 0

    for p in params:
        nb = p.numel() * 4
        if bucket and size + nb > bucket_bytes:
            flush()
        bucket.append(p)
        size += nb
    flush()
    return calls
