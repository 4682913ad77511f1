#!/usr/bin/env python
"""cfg5 of BASELINE.json: unrolled 10-iteration ADMM (conv_doe + deep_prior(ffdnet_color, sqrt=True)) with end-to-end
backward, data-parallel over the ranks of a torchrun job (one process per GPU, NCCL gradient all-reduce of the shared
trainable parameters).  Shapes follow the reference trainer: batch 2 per GPU, 768x768 images, 748x748 PSF
(optic/utils.py:158-166, optic/doe_model.py:165).  The x-update forward/backward are native kernels; the denoiser is a
torch module under bf16 autocast (its tcgen05 forward has no backward yet).  Random (seeded) weights.  Not the headline.

    python tools/bench_unrolled.py [--steps 5] [--size 768] [--psf 748] [--batch 2] [--precision bf16|fp32]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_unrolled.py ...
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delta-prox_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dprox_b200 as dp  # noqa: E402
from dprox_b200 import dist as ddist  # noqa: E402
from dprox_b200.denoisers import FFDNetColorDenoiser  # noqa: E402
from dprox_b200.linop import psf2otf2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--size", type=int, default=768)
    ap.add_argument("--psf", type=int, default=748)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, H, T = args.batch, args.size, args.iters
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    # stand-in for the DOE model: a trainable height-map-sized parameter -> normalised PSF (the optics forward model is §8f-3)
    torch.manual_seed(0)
    hm = torch.nn.Parameter(torch.rand(1, 3, args.psf, args.psf, device=dev))
    r0, s0 = dp.log_descent(49, 7.65, T, sigma=7.65 / 255)
    rhos, sigmas = torch.nn.Parameter(r0.to(dev)), torch.nn.Parameter(s0.to(dev))
    params = [hm, rhos, sigmas]
    opt = torch.optim.Adam(params, lr=1e-4)
    den = FFDNetColorDenoiser(seed=4, precision=args.precision).to(dev)
    x, y, PSF = dp.Variable(), dp.Placeholder(), dp.Placeholder()
    data_term = dp.sum_squares(dp.conv_doe(x, PSF, circular=True), y)
    reg_term = dp.deep_prior(x, denoiser=den, sqrt=True)
    solver = dp.specialize(dp.compile(data_term + reg_term, method="admm", device=dev), method="unroll", max_iter=T)

    def step():
        gt = torch.rand(B, 3, H, H, device=dev, generator=g)
        psf = hm.abs() / hm.abs().sum(dim=(-2, -1), keepdim=True)
        otf = psf2otf2(psf, gt.shape)
        inp = torch.real(torch.fft.ifftn(otf * torch.fft.fftn(gt, dim=[-2, -1]), dim=[-2, -1])).float()
        inp = inp + (7.65 / 255) * torch.randn(gt.shape, device=dev, generator=g)
        y.value, PSF.value = inp, psf.detach()
        out = solver.solve(x0=inp, rhos=rhos, lams={reg_term: sigmas})
        loss = torch.nn.functional.mse_loss(gt, out)
        loss.backward()
        ddist.allreduce_gradients(params)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        sps = args.steps / (float(ms) * 1e-3)
        print(json.dumps({"workload": f"unrolled {T}-iter ADMM conv_doe+deep_prior(ffdnet_color) train step, {B} x [3,{H},{H}] per GPU, "
                                      f"psf {args.psf}, denoiser {args.precision}", "n_gpus": world, "train_steps_per_s": sps,
                          "images_per_s": sps * B * world, "ms_per_step": float(ms) / args.steps, "final_loss": float(loss.detach()),
                          "reference_published": "1.48-1.51 steps/s, bs=2, unstated GPU (BASELINE.md)"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
