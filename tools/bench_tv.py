#!/usr/bin/env python
"""Anisotropic-TV deconvolution: sum_squares(conv(x) - b) + norm1(grad_h x) + norm1(grad_w x), ADMM,
B x [3,S,S]; problem-iterations/s and fraction of its 40 B/element HBM roofline (24 B + 16 B for the second prox term).
    python tools/bench_tv.py [--size 2048] [--batch 8] [--iters 30]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delta-prox_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dprox_b200 as dp  # noqa: E402
from bench import load_peaks, psf_gaussian  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
b = torch.rand(a.batch, 3, a.size, a.size, device=dev, generator=g)
x = dp.Variable()
th, tw = dp.norm1(dp.grad(x, dim=0)), dp.norm1(dp.grad(x, dim=1))
s = dp.compile(dp.sum_squares(dp.conv(x, psf_gaussian(15, 5.0)) - b) + th + tw, method="admm", device=dev)
rhos = torch.full((a.iters,), 1.0, device=dev)
lam = torch.full((a.iters,), 0.02, device=dev)
state = s.initialize(b)
for _ in range(2):
    state = s.iters(state, rhos, {th: lam, tw: lam}, a.iters)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    state = s.iters(state, rhos, {th: lam, tw: lam}, a.iters)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (a.steps * a.iters)
N = b.numel()
print(json.dumps({"workload": f"admm anisotropic-TV deconv, {a.batch} x [3,{a.size},{a.size}]", "tier": s.spec.tier,
                  "problem_iters_per_s": a.batch / (ms * 1e-3), "ms_per_iteration": ms,
                  "frac_of_40B_roofline": 40.0 * N / (ms * 1e-3) / 1e9 / load_peaks()[0]}))
