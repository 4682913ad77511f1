"""N > 1 host logic on CPU: world_size-2 gloo process group.  Shard ranges, per-sample schedule
sharding, the gather of result shards and the residual all-reduce of the stopping rule."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dprox_b200 import dist as ddist
from dprox_b200.algo import ResidualStop


def test_shard_ranges_cover_the_batch():
    for n in (1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [ddist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 5
        x = torch.arange(B * 6, dtype=torch.float32).reshape(B, 1, 2, 3)
        lo, hi = ddist.shard_range(B)
        local = ddist.shard_batch(x)
        assert local.shape[0] == hi - lo
        sched = ddist.shard_schedule({"a": torch.rand(B, 4), "b": torch.rand(4)}, B)
        assert sched["a"].shape == (hi - lo, 4) and sched["b"].shape == (4,)
        back = ddist.gather_batch(local * 2, B)
        ok_gather = torch.equal(back, x * 2)
        # residual stop: rank-local sums are all-reduced; both ranks must take the same decision
        stop = ResidualStop(abstol=0.0, reltol=0.5, every=1)
        sums = torch.tensor([1.0, 1.0, 16.0, 16.0]) if rank == 0 else torch.tensor([3.0, 3.0, 48.0, 48.0])
        dec = stop.converged(sums, n_elems=10)
        r, s = stop.history[-1]
        ok_stop = dec and abs(r - 2.0) < 1e-6 and abs(s - 2.0) < 1e-6          # sqrt(1+3) = 2 <= 0.5*sqrt(64)
        # asynchronous form: decisions are consumed one check late, identically on every rank
        lazy = ResidualStop(abstol=0.0, reltol=0.5, every=1, lag=1)
        lazy.submit(sums * torch.tensor([100.0, 100.0, 1.0, 1.0]), n_elems=10)    # check 0: not converged
        first = lazy.poll()                                                    # nothing old enough yet
        lazy.submit(sums, n_elems=10)                                          # check 1: converged
        second = lazy.poll()                                                   # consumes check 0
        lazy.submit(sums, n_elems=10)
        third = lazy.poll()                                                    # consumes check 1 -> stop
        ok_stop = ok_stop and (first, second, third) == (False, False, True) and len(lazy.history) == 2
        # solve_sharded plumbing with a stand-in solver (no GPU here): each rank handles its shard only
        class Fake:
            def solve(self, x0, rhos, lams, **kw):
                return x0 + float(rhos.sum()) + lams["f"].sum(dim=1).view(-1, 1, 1, 1)
        per_sample = torch.arange(B, dtype=torch.float32).view(B, 1).repeat(1, 2)
        out = ddist.solve_sharded(lambda lo, hi: (Fake(), None), x, rhos=torch.ones(3), lams={"f": per_sample})
        ok_solve = torch.equal(out, x + 3.0 + 2 * torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1))
        # DDP-style gradient all-reduce of shared trainable parameters (unrolled training, config 5)
        p1, p2, p3 = (torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2)),
                      torch.nn.Parameter(torch.zeros(1), requires_grad=False))
        p1.grad = torch.full((3,), float(rank + 1))
        if rank == 0:
            p2.grad = torch.ones(2, 2)                       # rank 1 has no gradient for p2: contributes zeros
        ncalls = ddist.allreduce_gradients([p1, p2, p3], bucket_bytes=12)
        ok_grad = (ncalls == 2 and torch.allclose(p1.grad, torch.full((3,), 1.5)) and torch.allclose(p2.grad, torch.full((2, 2), 0.5))
                   and p3.grad is None)
        q.put((rank, ok_gather, ok_stop, ok_solve, ok_grad))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(all(r[1:]) for r in res), res
