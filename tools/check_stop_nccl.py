#!/usr/bin/env python
"""Multi-GPU check of the only collective on the path: a batch of problems sharded over the ranks of a torchrun
job, iterations with the opt-in residual stopping rule whose sums are all-reduced over NCCL; every rank must stop at the same
iteration, and the gathered result must equal a single-process run of the whole batch.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_stop_nccl.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delta-prox_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dprox_b200 as dp  # noqa: E402
from bench import psf_gaussian  # noqa: E402
from dprox_b200 import dist as ddist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, H, W = 4 * world, 256, 512
g = torch.Generator().manual_seed(7)
b_all = torch.rand(B, 3, H, W, generator=g) - 0.3
psf = psf_gaussian(9, 2.0)


def make(lo, hi):
    x = dp.Variable()
    return dp.compile(dp.sum_squares(dp.conv(x, psf) - b_all[lo:hi].to(dev)) + dp.nonneg(x), method="admm", device=dev), None


stop = dp.ResidualStop(abstol=1e-3, reltol=1e-2, every=5)
lo, hi = ddist.shard_range(B)
solver, _ = make(lo, hi)
out_local = solver.solve(x0=b_all[lo:hi].to(dev), max_iter=200, stop=stop)
its = torch.tensor([solver.iterations_run], device=dev)
all_its = [torch.zeros_like(its) for _ in range(world)]
dist.all_gather(all_its, its)
out = ddist.gather_batch(out_local, B)
ok_same_stop = len({int(t) for t in all_its}) == 1
if rank == 0:
    stop1 = dp.ResidualStop(abstol=1e-3, reltol=1e-2, every=5, group=dist.new_group([0]))   # same rule, this process only
else:
    dist.new_group([0])
if rank == 0:
    x = dp.Variable()
    ref = dp.compile(dp.sum_squares(dp.conv(x, psf) - b_all.to(dev)) + dp.nonneg(x), method="admm", device=dev)
    full = ref.solve(x0=b_all.to(dev), max_iter=int(all_its[0]))
    err = float((out - full).norm() / full.norm())
    print(f"world={world} stopped at iteration {int(all_its[0])} on every rank: {ok_same_stop}; sharded vs single-process rel err {err:.2e}")
    assert ok_same_stop and int(all_its[0]) < 200 and err < 1e-5
dist.destroy_process_group()
