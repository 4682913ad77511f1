#!/usr/bin/env python
"""cfg5 of BASELINE.json: unrolled 10-iteration ADMM (conv_doe + deep_prior(ffdnet_color, sqrt=True)) with end-to-end
backward, fed by the DOE optics forward model (height map -> PSF -> data formation), data-parallel over the ranks of a
torchrun job (one process per GPU, NCCL gradient all-reduce of the shared trainable parameters).  Shapes follow the reference trainer: batch 2 per GPU, 768x768 images, 748x748 PSF
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
from dprox_b200.optics import DOEModelConfig, build_doe_model, img_psf_conv  # noqa: E402


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
    # the DOE model of the reference trainer (optic/doe_model.py:156-187): 1496^2 wavefront, Fresnel propagation, 748^2 PSF --
    # native forward and backward kernels (dprox_b200/optics.py); the trainable height map is ~9 MB
    doe = build_doe_model(DOEModelConfig(patch_size=args.psf, wave_resolution=(2 * args.psf, 2 * args.psf))).to(dev)
    r0, s0 = dp.log_descent(49, 7.65, T, sigma=7.65 / 255)
    rhos, sigmas = torch.nn.Parameter(r0.to(dev)), torch.nn.Parameter(s0.to(dev))
    params = [doe.height_map.height_map_sqrt, rhos, sigmas]
    opt = torch.optim.Adam(params, lr=1e-4)
    den = FFDNetColorDenoiser(seed=4, precision=args.precision).to(dev)
    x, y, PSF = dp.Variable(), dp.Placeholder(), dp.Placeholder()
    data_term = dp.sum_squares(dp.conv_doe(x, PSF, circular=True), y)
    reg_term = dp.deep_prior(x, denoiser=den, sqrt=True)
    solver = dp.specialize(dp.compile(data_term + reg_term, method="admm", device=dev), method="unroll", max_iter=T)

    def step():
        gt = torch.rand(B, 3, H, H, device=dev, generator=g)
        psf = doe.get_psf()                                   # height map -> phase -> Fresnel -> intensity -> 2x area pool -> / sum
        inp = img_psf_conv(gt, psf, circular=True)            # data formation (carries d/d height map)
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
