#!/usr/bin/env python
"""cfg2 of BASELINE.json: 2048x2048 circular deconvolution + deep_prior(ffdnet_color) + nonneg, ADMM, one B200.
Reports (i) the FFDNet-color forward alone — native tcgen05 bf16 network vs the framework's cuDNN convolutions —
as TFLOP/s against the measured dense bf16 peak (MEASURED_PEAKS.json), and (ii) plug-and-play ADMM iterations/s.
Random (seeded) weights: the pretrained file needs a download.  Not the headline bench (see bench.py)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delta-prox_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dprox_b200 as dp  # noqa: E402
from dprox_b200.denoisers import FFDNetColorDenoiser  # noqa: E402

FLOP_PER_PIXEL = 2 * 9 * (13 * 96 + 10 * 96 * 96 + 96 * 12) / 4.0      # 0.4255 MFLOP per full-res pixel


def timeit(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--iters", type=int, default=24)
    args = ap.parse_args()
    dev = torch.device("cuda")
    B, H, W = args.batch, args.size, args.size
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("bf16_tflops", 1590.0))
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(B, 3, H, W, device=dev, generator=g)
    sig = torch.full((B,), 0.05, device=dev)
    flop = FLOP_PER_PIXEL * B * H * W
    out = {"workload": f"FFDNet-color forward, {B} x [3,{H},{W}]", "flop_per_call": flop}
    fast = FFDNetColorDenoiser(seed=4, precision="bf16").to(dev)
    ref = FFDNetColorDenoiser(seed=4).to(dev)
    with torch.no_grad():
        ms = timeit(lambda: fast.denoise(x, sig), 10)
        out["native_tcgen05_bf16"] = {"ms": ms, "tflops": flop / ms / 1e9, "frac_of_measured_bf16_peak": flop / ms / 1e9 / peak}
        ms = timeit(lambda: ref.denoise(x, sig), 2, warm=1)
        out["cudnn_fp32"] = {"ms": ms, "tflops": flop / ms / 1e9}
        m16 = ref.model.to(torch.bfloat16).to(memory_format=torch.channels_last)
        xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ms = timeit(lambda: m16(xb, sig.to(torch.bfloat16)), 5)
        out["cudnn_bf16_channels_last"] = {"ms": ms, "tflops": flop / ms / 1e9}
        ref.model.float()
    # plug-and-play ADMM (cfg2) with the native denoiser as the external prox
    psf = np.ones((15, 15, 1), "float32") / 225.0
    xv = dp.Variable()
    b = dp.conv(xv, psf).to(dev).forward(x)
    prior, nn_ = dp.deep_prior(xv, denoiser=fast), dp.nonneg(xv)
    solver = dp.compile(dp.sum_squares(dp.conv(xv, psf) - b) + prior + nn_, method="admm", device=dev)
    rhos, sigmas = dp.log_descent(35, 30, args.iters)
    with torch.no_grad():
        ms = timeit(lambda: solver.solve(x0=b, rhos=rhos, lams={prior: sigmas, nn_: 0.02}, max_iter=args.iters), 2, warm=1)
    out["pnp_admm"] = {"iters": args.iters, "ms_per_solve": ms, "problem_iters_per_s": B * args.iters / (ms * 1e-3),
                       "denoiser_share": out["native_tcgen05_bf16"]["ms"] * args.iters / ms}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
