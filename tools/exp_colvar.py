"""Experiment harness (run under gpurun): A/B of the fused-kernel variants selected by environment switches.

For each `--vars` entry (a comma-separated list of `NAME=VALUE;NAME=VALUE` environment settings, `-` = defaults):
  * correctness: 6 ADMM iterations on [4,3,S,S] against the default build of the same engine (max rel-L2 on x, v, u),
  * speed: the headline workload (8 x [3,S,S], 50 iterations per step), CUDA events, best and median of `--reps` steps.
Prints one JSON line per variant.  Not a bench line (bench.py is the bench).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "delta-prox_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vars", default="-")
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--height", type=int, default=0, help="non-square problems: rows (default --size)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--backend", type=int, default=2, help="2 = fused engine, 1 = cuFFT engine")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--method", default="admm")
    args = ap.parse_args()
    import dprox_b200 as dp
    from bench import psf_gaussian, make_measurements

    dev = torch.device("cuda", 0)
    S = args.size
    Hh, Ww = args.height or S, args.width or S
    psf = psf_gaussian(15, 5.0)

    def setenv(spec):
        keys = []
        if spec != "-":
            for kv in spec.split(";"):
                k, v = kv.split("=")
                os.environ[k] = v
                keys.append(k)
        return keys

    def build(B):
        x = dp.Variable()
        y = dp.Placeholder()
        op = dp.conv(x, psf)
        solver = dp.compile(dp.sum_squares(dp.conv(x, psf) - y) + dp.nonneg(x), method=args.method, device=dev, fft_backend=args.backend)
        img, noise = make_measurements(B, 3, Hh, Ww, seed=99, device=dev)
        b = op.to(dev).forward(img) + noise
        y.value = b
        return solver, b

    ref = None
    for spec in args.vars.split(","):
        keys = setenv(spec)
        out = {"var": spec}
        try:
            solver, b = build(4)
            st = solver.solve(x0=b, rhos=1.0, lams=0.02, max_iter=6, return_full_states=True)
            cur = [st[0].clone(), st[1][0].clone()] + ([st[2][0].clone()] if len(st) > 2 and len(st[2]) else [])
            if ref is None:
                ref = cur
            out["rel_vs_first"] = [rel(c, r) for c, r in zip(cur, ref)]
            del solver, b, st
            torch.cuda.empty_cache()
            solver, b = build(args.batch)
            T = args.iters
            rhos = torch.full((T,), 1.0, device=dev)
            lams = torch.full((T,), 0.02, device=dev)
            state = solver.initialize(b)
            for _ in range(3):
                state = solver.iters(state, rhos, lams, T)
            torch.cuda.synchronize()
            times = []
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                state = solver.iters(state, rhos, lams, T)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) / T * 1e3)
            times.sort()
            out["us_per_iteration_best"] = times[0]
            out["us_per_iteration_median"] = times[len(times) // 2]
            nbytes = (16.0 if args.method == "hqs" else 24.0) * args.batch * 3 * Hh * Ww
            out["frac_median"] = nbytes / (out["us_per_iteration_median"] * 1e-6) / 1e9 / 6534.1
            del solver, b, state
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out["error"] = repr(e)[:300]
        for k in keys:
            os.environ.pop(k, None)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
