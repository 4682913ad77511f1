#!/usr/bin/env python
"""What the host can feed: every rank copies bench.py's end-to-end payload (H2D of b and D2H of x, 403 MB each for 8 problems
[3,2048,2048]) between pinned host memory and its GPU with NO compute, on two streams (one per direction).  Run alone (N = 1)
and under torchrun on all GPUs of the box: the aggregate rate at N = 8 is the ceiling of bench.py's `e2e` leg.
    python tools/host_copy_ceiling.py            |   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/host_copy_ceiling.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import pin_to_gpu_numa_node  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
pin_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 8 * 3 * 2048 * 2048
h_in, h_out = torch.rand(n).pin_memory(), torch.empty(n).pin_memory()
d_in, d_out = torch.empty(n, device=dev), torch.rand(n, device=dev)
s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
res = {}
for mode in ("h2d", "d2h", "both"):
    for it in range(2 + 6):
        if it == 2:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s_up.wait_event(e0); s_dn.wait_event(e0)
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s_up):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s_dn):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream(dev).wait_stream(s_up)
    torch.cuda.current_stream(dev).wait_stream(s_dn)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    nbytes = 6 * n * 4 * (2 if mode == "both" else 1)
    res[mode] = {"gb_per_s_per_rank": nbytes / (float(ms) * 1e-3) / 1e9, "gb_per_s_aggregate": world * nbytes / (float(ms) * 1e-3) / 1e9}
if rank == 0:
    step_bytes = 2 * n * 4
    res["e2e_ceiling_problem_iters_per_s"] = world * 8 * 50 / (step_bytes / (res["both"]["gb_per_s_per_rank"] * 1e9 / 2 * 2))
    print(json.dumps({"n_gpus": world, "payload_mb_per_direction": n * 4 / 1e6, "cpus": os.cpu_count(), **res}))
if world > 1:
    dist.destroy_process_group()
