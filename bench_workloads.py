"""bench_workloads.py — the other BASELINE.json configurations as `bench.py --workload cfgN` lines (same JSON contract as the
headline: value with resident inputs, e2e with host buffers, roofline of the dominant kernel, cpu_baseline, clocks).

  cfg2  2048x2048 circular deconvolution + deep_prior(ffdnet_color) + nonneg, ADMM x 24 (DPIR log_descent schedule), 1 GPU
        dominant kernel: the tcgen05 3x3 convolutions of the denoiser -> TENSOR roofline (measured bf16 peak)
  cfg3  CS-MRI: subsampled-FFT plugin operator + anisotropic TV, ADMM with the fused-kernel PCG inner solve -> HBM roofline
        (28 B / element / CG step + 40 B / element / outer iteration)
  cfg4  64 x [3,1024,1024] deconvolution over 8 GPUs, HQS x 24 (8 problems per GPU): the headline code path with
        --method hqs --size 1024 --batch 8 --iters 24 (handled by bench.py itself)
  cfg5  unrolled 10-iteration ADMM (conv_doe + deep_prior(ffdnet_color, sqrt=True)) train step with end-to-end backward, bf16
        denoiser, DOE optics forward model, data parallel with an NCCL gradient all-reduce -> TENSOR roofline

Random (seeded) denoiser weights: the pretrained file needs a download.  The reference arm (`--impl reference`) times the
UNMODIFIED reference (baseline/_ref) -- or the oracle port when that directory is absent -- on the host cores on a bounded
sample of the same workload.
"""
import json
import os
import sys
import time

import numpy as np
import torch

import bench as B0

FLOP_PER_PIXEL = 2 * 9 * (13 * 96 + 10 * 96 * 96 + 96 * 12) / 4.0        # FFDNet-color: 0.4255 MFLOP per full-resolution pixel


def _peaks():
    p = os.path.join(B0.ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json; bf16 = sustained figure: kernel timed inside a long step)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class _Dist:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.rank = int(os.environ.get("RANK", 0))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        B0.pin_to_gpu_numa_node(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_ms(self, ms):
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def _timed(D, fn, steps, warmup):
    """W warm-up + K timed calls of fn bracketed by barrier + synchronize, CUDA events, max over ranks -> ms total"""
    for _ in range(warmup):
        fn()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    D.barrier()
    return D.max_ms(e0.elapsed_time(e1))


class _TimedDenoiser:
    """collects the CUDA-event pairs NativeFFDNet records around every call (forward, forward with saved activations, backward)
    on the launching stream while `on` is set"""

    def __init__(self, native):
        self.native, self.ev = native, []

    def install(self):
        return self

    @property
    def on(self):
        return self.native.profile is not None

    @on.setter
    def on(self, flag):
        self.native.profile = self.ev if flag else None

    def summary(self):
        tot_ms = {"fwd": 0.0, "bwd": 0.0}
        flop = {"fwd": 0.0, "bwd": 0.0}
        n = {"fwd": 0, "bwd": 0}
        for kind, a, b, pix in self.ev:
            tot_ms[kind] += a.elapsed_time(b)
            flop[kind] += FLOP_PER_PIXEL * pix
            n[kind] += 1
        return tot_ms, flop, n


def _line(args, D, metric, unit, value, ms_total, workload, roof, cpu, e2e, launches, clk, extra=None):
    d = {"metric": metric, "value": value, "unit": unit, "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
         "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": workload.pop("dtype"),
         "data": "synthetic", "config": workload, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
         "clocks": clk}
    if extra:
        d.update(extra)
    print(json.dumps(d))


# =====================================================================================================================
#  cfg2: plug-and-play deconvolution with the deep denoiser
# =====================================================================================================================

def _cfg2_cpu(H, W, T_cpu, threads):
    """reference (or oracle port) on the host cores: 1 problem [3,H,W], T_cpu ADMM iterations with the seeded FFDNet"""
    import dprox_oracle as orc
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    img = torch.rand(1, 3, H, W, generator=g)
    psf = orc.point_spread_function(15, 5)
    b = orc.Conv(psf, orc.Identity()).fwd(img) + 0.01 * torch.randn(1, 3, H, W, generator=g)
    rhos, sigmas = orc.log_descent(35, 30, T_cpu)
    ws = orc.ffdnet_random_weights(4)
    ref = B0._reference_module()
    with torch.no_grad():
        if ref is not None:
            from dprox.proxfn.pnp.denoisers.base import Denoiser
            from dprox.proxfn.pnp.denoisers.models.network_ffdnet import FFDNet

            class Rand(Denoiser):
                def __init__(self):
                    super().__init__()
                    self.model = FFDNet(in_nc=3, out_nc=3, nc=96, nb=12, act_mode="R")
                    self.model.load_state_dict({f"model.{2 * i}.{k}": v for i, (w_, b_) in enumerate(ws) for k, v in (("weight", w_), ("bias", b_))})

                def _denoise(self, x, sigma):
                    return self.model(x, sigma)

            x = ref.Variable()
            prior, nn_ = ref.deep_prior(x, denoiser=Rand()), ref.nonneg(x)
            solver = ref.compile(ref.sum_squares(ref.conv(x, psf) - b) + prior + nn_, method="admm", device="cpu")
            run = lambda: solver.solve(x0=b, rhos=rhos, lams={prior: sigmas, nn_: 0.02}, max_iter=T_cpu)
            kind = "reference"
        else:
            prior, nn_ = orc.Term("deep_prior", denoiser=lambda v, s: orc.ffdnet_forward(ws, v, s)), orc.Term("nonneg")
            solver = orc.Solver([orc.Term("sum_squares", orc.Conv(psf, orc.Identity()), c=b), prior, nn_], "admm")
            run = lambda: solver.solve(b, rhos=rhos, lams={prior: sigmas, nn_: 0.02}, max_iter=T_cpu)
            kind = "port"
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
    return T_cpu / dt, dt, kind


def run_cfg2(args):
    import dprox_b200 as dp
    from dprox_b200 import _cabi as cabi
    from dprox_b200.denoisers import FFDNetColorDenoiser, NativeFFDNet
    D = _Dist()
    dev, lib = D.dev, cabi.lib()
    Bn, H, W, T = (args.batch if args.batch_set else 2), args.size, args.size, (args.iters if args.iters_set else 24)
    psf = B0.psf_gaussian(15, 5.0)
    split = getattr(args, "denoiser", "fp32") == "fp32"       # fp16 operand pairs: three MMAs per k-step, fp32-class accuracy
    den = FFDNetColorDenoiser(seed=4, precision="fp32" if split else "bf16").to(dev).requires_grad_(False)
    den._native = NativeFFDNet(den.model, dev, split=split)
    timer = _TimedDenoiser(den._native).install()
    x, y = dp.Variable(), dp.Placeholder()
    prior, nn_ = dp.deep_prior(x, denoiser=den), dp.nonneg(x)
    solver = dp.compile(dp.sum_squares(dp.conv(x, psf) - y) + prior + nn_, method="admm", device=dev)
    img, noise = B0.make_measurements(Bn, 3, H, W, seed=4321 + D.rank, device=dev)
    b_dev = dp.conv(dp.Variable(), psf).to(dev).forward(img + 0.3) + noise
    del img, noise
    y.value = b_dev
    rhos, sigmas = dp.log_descent(35, 30, T)
    rhos, sigmas, lam_nn = rhos.to(dev), sigmas.to(dev), torch.full((T,), 0.02, device=dev)
    b_host = b_dev.cpu().pin_memory()
    out_host = torch.empty_like(b_host).pin_memory()

    def step():
        with torch.no_grad():
            return solver.solve(x0=b_dev, rhos=rhos, lams={prior: sigmas, nn_: lam_nn}, max_iter=T)

    def e2e_step():
        with torch.no_grad():
            bd = b_host.to(dev, non_blocking=True)
            y.value = bd
            xs = solver.solve(x0=bd, rhos=rhos, lams={prior: sigmas, nn_: lam_nn}, max_iter=T)
            out_host.copy_(xs, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(args.warmup):
        step()
    l0 = lib.dpx_launch_count()
    timer.on = True
    with B0.ClockSampler(D.local) as clk:
        ms = _timed(D, step, args.steps, 0)
        time.sleep(0.15)
    timer.on = False
    launches = lib.dpx_launch_count() - l0
    value = D.world * Bn * T * args.steps / (ms * 1e-3)
    e_steps = max(1, args.steps // 2)
    ems = _timed(D, e2e_step, e_steps, 1)
    e2e_value = D.world * Bn * T * e_steps / (ems * 1e-3)
    if D.rank == 0:
        _, peak_tf, src = _peaks()
        tot_ms, flop, n = timer.summary()
        mmas = 3 if split else 1                                # tensor-core work executed per algorithmic multiply-add
        achieved = mmas * flop["fwd"] / (tot_ms["fwd"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": None,
                "peak_source": src,
                "unit_of_work": f"one FFDNet-color forward of {Bn} x [3,{H},{W}] = {FLOP_PER_PIXEL * Bn * H * W / 1e12:.3f} TFLOP algorithmic"
                                + (" x 3 executed (operands as fp16 pairs hi + 2^-11 lo': a_hi w_hi, a_lo w_hi, a_hi w_lo into two TMEM "
                                   "accumulators; 23 tcgen05 convolution launches" if split else " (12 tcgen05 convolution launches")
                                + " + 2 layout kernels), timed with CUDA events around every call",
                "algorithmic_tflops": flop["fwd"] / (tot_ms["fwd"] * 1e-3) / 1e12,
                "avg_call_ms": tot_ms["fwd"] / max(1, n["fwd"]), "calls": n["fwd"],
                "denoiser_share_of_step": tot_ms["fwd"] / ms}
        cpu = None
        if not args.skip_cpu:
            threads = os.cpu_count() or 1
            r, dt, kind = _cfg2_cpu(H, W, 3, threads)
            cpu = {"value": r, "unit": B0.UNIT, "cores": threads, "kind": kind,
                   "sample": f"1 problem [3,{H},{W}] x 3 ADMM iterations with the fp32 FFDNet ({dt:.1f} s), torch-CPU"}
        dn = "fp16-pair split tcgen05, fp32-class" if split else "bf16 tcgen05"
        wl = {"workload": f"cfg2: admm deconv + deep_prior(ffdnet_color, {dn}) + nonneg, {Bn} problems/GPU [3,{H},{W}], psf gaussian 15/5, "
                          f"log_descent(35,30,{T}), {T} iterations per step", "batch_per_gpu": Bn, "iters_per_step": T,
              "l2_policy": "every activation tensor of the denoiser (403 MB) and every state array exceed the 126 MB L2",
              "parallelism": f"dp{D.world} (problem shards, no collective)",
              "dtype": ("f16x2 split (denoiser, fp32-class) / f32 (iteration)" if split else "bf16 (denoiser) / f32 (iteration)")}
        _line(args, D, "ADMM iters/sec, 2Kx2K PnP deconv (deep_prior ffdnet_color)", B0.UNIT, value, ms, wl, roof, cpu,
              {"value": e2e_value, "unit": B0.UNIT, "h2d_bytes_per_step": int(b_host.numel() * 4), "d2h_bytes_per_step": int(b_host.numel() * 4),
               "steps": e_steps}, launches, clk.summary())
    D.close()


# =====================================================================================================================
#  cfg3: CS-MRI, plugin operator + TV, PCG inner solve
# =====================================================================================================================

def _cfg3_problem(Bn, H, W, seed, dev):
    g = torch.Generator().manual_seed(seed)
    img = torch.zeros(Bn, 1, H, W)
    for b in range(Bn):
        y0, x0 = int(H * (0.2 + 0.1 * torch.rand(1, generator=g))), int(W * (0.25 + 0.1 * torch.rand(1, generator=g)))
        img[b, :, y0:y0 + H // 2, x0:x0 + W // 3] = 1.0
        img[b, :, H // 3:H // 3 + H // 5, W // 8:W - W // 8] += 0.5
    mask = (torch.rand(1, 1, H, W, generator=g) < 0.3).float()
    mask[..., :4, :4] = 1; mask[..., -4:, :4] = 1; mask[..., :4, -4:] = 1; mask[..., -4:, -4:] = 1
    return img.to(dev), mask.to(dev)


def run_cfg3(args):
    import dprox_b200 as dp
    from dprox_b200 import _cabi as cabi
    D = _Dist()
    dev, lib = D.dev, cabi.lib()
    Bn, H, W = (args.batch if args.batch_set else 8), (args.size if args.size_set else 256), (args.size if args.size_set else 256)
    T, CG = (args.iters if args.iters_set else 30), 30
    img, mask = _cfg3_problem(Bn, H, W, 7 + D.rank, dev)
    calls = [0]

    def fwd(x, step=0):
        calls[0] += 1
        return mask * torch.fft.fft2(x, norm="ortho")

    def adj(y, step=0):
        return torch.real(torch.fft.ifft2(mask * y, norm="ortho")).contiguous()

    y0 = fwd(img)
    x0 = adj(y0)
    x = dp.Variable()
    A = dp.LinOpFactory(fwd, adj)
    f1, f2 = dp.norm1(dp.grad(x, dim=0)), dp.norm1(dp.grad(x, dim=1))
    cfg = dp.LinearSolveConfig(rtol=1e-6, max_iters=CG, solver_type="pcg")
    solver = dp.compile(dp.sum_squares(A(x), y0) + f1 + f2, method="admm", device=dev, linear_solve_config=cfg)
    rhos, lam = torch.full((T,), 1.0, device=dev), torch.full((T,), 0.05, device=dev)
    y0_host = torch.view_as_real(y0).cpu().pin_memory()
    out_host = torch.empty(Bn, 1, H, W).pin_memory()

    def step():
        with torch.no_grad():
            return solver.solve(x0=x0, rhos=rhos, lams={f1: lam, f2: lam}, max_iter=T)

    def e2e_step():
        with torch.no_grad():
            yd = torch.view_as_complex(y0_host.to(dev, non_blocking=True))
            xs = solver.solve(x0=adj(yd), rhos=rhos, lams={f1: lam, f2: lam}, max_iter=T)
            out_host.copy_(xs, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(args.warmup):
        out = step()
    err = float((out - img).norm() / img.norm())
    from dprox_b200 import linalg as _linalg
    l0, c0, g0 = lib.dpx_launch_count(), _linalg.steps_enqueued[0], _linalg.replayed_launches[0]
    with B0.ClockSampler(D.local) as clk:
        ms = _timed(D, step, args.steps, 0)
        time.sleep(0.15)
    # CG steps enqueued per solve (replayed graph steps never pass through the Python operator, so `calls` cannot count them;
    # the count includes the few gated no-op steps the host queues before the lagged stop flag reaches it)
    launches, cg_steps = lib.dpx_launch_count() - l0 + _linalg.replayed_launches[0] - g0, (_linalg.steps_enqueued[0] - c0) / args.steps
    value = D.world * Bn * T * args.steps / (ms * 1e-3)
    e_steps = max(1, args.steps // 2)
    ems = _timed(D, e2e_step, e_steps, 1)
    e2e_value = D.world * Bn * T * e_steps / (ems * 1e-3)
    if D.rank == 0:
        hbm, _, src = _peaks()
        N = Bn * H * W
        alg = (28.0 * cg_steps + 40.0 * T) * N                       # 28 B per CG step, 24 + 16 B per outer iteration (2 prox terms)
        achieved = alg / (ms / args.steps * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                "peak_source": src,
                "unit_of_work": f"one solve = {T} ADMM iterations x up to {CG} PCG steps ({cg_steps:.0f} CG steps enqueued) of "
                                f"{Bn} x [1,{H},{W}]: 28 B/element/CG step + 40 B/element/iteration (operator FFTs excluded)",
                "note": "a 256x256 problem is 0.26 MB per array: the solve is launch-latency bound, not bandwidth bound"}
        cpu = None
        if not args.skip_cpu:
            cpu = _cfg3_cpu(H, W, T, CG, os.cpu_count() or 1)
        wl = {"workload": f"cfg3: CS-MRI, sum_squares(mask*fft2(x), y) + norm1(grad_h) + norm1(grad_w), ADMM with PCG (<= {CG} steps, rtol 1e-6), "
                          f"{Bn} problems/GPU [1,{H},{W}], 30 % sampling, {T} iterations per step", "batch_per_gpu": Bn, "iters_per_step": T,
              "l2_policy": "working set fits L2 (the reference's own problem size)", "rel_err_vs_truth": err,
              "parallelism": f"dp{D.world} (problem shards, no collective)", "dtype": "f32"}
        _line(args, D, "ADMM iters/sec, CS-MRI TV (PCG inner solve)", B0.UNIT, value, ms, wl, roof, cpu,
              {"value": e2e_value, "unit": B0.UNIT, "h2d_bytes_per_step": int(y0_host.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 4),
               "steps": e_steps}, launches, clk.summary())
    D.close()


def _cfg3_cpu(H, W, T, CG, threads):
    import dprox_oracle as orc
    torch.set_num_threads(threads)
    img, mask = _cfg3_problem(1, H, W, 7, "cpu")
    fwd = lambda x, step=0: mask * torch.fft.fft2(x, norm="ortho")
    adj = lambda y, step=0: torch.real(torch.fft.ifft2(mask * y, norm="ortho"))
    y0 = fwd(img)
    x0 = adj(y0)
    ref = B0._reference_module()
    with torch.no_grad():
        if ref is not None:
            from dprox.linalg import LinearSolveConfig
            x = ref.Variable()
            A = ref.LinOpFactory(fwd, adj)
            fns = ref.sum_squares(A(x), y0) + ref.norm1(ref.grad(x, dim=0)) + ref.norm1(ref.grad(x, dim=1))
            solver = ref.compile(fns, method="admm", device="cpu", linear_solve_config=LinearSolveConfig(rtol=1e-6, max_iters=CG, solver_type="pcg"))
            run = lambda: solver.solve(x0=x0, rhos=1.0, lams=0.05, max_iter=T)
            kind = "reference"
        else:
            data = orc.Term("sum_squares", orc.BlackBox(fwd, adj, orc.Identity()), b=y0)
            psi = [orc.Term("norm1", orc.Grad(0, orc.Identity())), orc.Term("norm1", orc.Grad(1, orc.Identity()))]
            solver = orc.Solver([data] + psi, "admm", solver_type="pcg", rtol=1e-6, max_iters=CG)
            run = lambda: solver.solve(x0, rhos=1.0, lams=0.05, max_iter=T)
            kind = "port"
        run()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 10.0:
            run()
            n += 1
        dt = time.perf_counter() - t0
    return {"value": n * T / dt, "unit": B0.UNIT, "cores": threads, "kind": kind,
            "sample": f"1 problem [1,{H},{W}] x {T} ADMM iterations, {n} solves in {dt:.1f} s, torch-CPU"}


# =====================================================================================================================
#  cfg5: unrolled training step
# =====================================================================================================================

def run_cfg5(args):
    import dprox_b200 as dp
    from dprox_b200 import _cabi as cabi
    from dprox_b200 import dist as ddist
    from dprox_b200.denoisers import FFDNetColorDenoiser, NativeFFDNet
    from dprox_b200.optics import DOEModelConfig, build_doe_model, img_psf_conv
    D = _Dist()
    dev, lib = D.dev, cabi.lib()
    Bn, H, T = (args.batch if args.batch_set else 2), (args.size if args.size_set else 768), (args.iters if args.iters_set else 10)
    psf_n = H - 20 if H > 64 else H
    g = torch.Generator(device=dev).manual_seed(100 + D.rank)
    doe = build_doe_model(DOEModelConfig(patch_size=psf_n, wave_resolution=(2 * psf_n, 2 * psf_n))).to(dev)
    r0, s0 = dp.log_descent(49, 7.65, T, sigma=7.65 / 255)
    rhos, sigmas = torch.nn.Parameter(r0.to(dev)), torch.nn.Parameter(s0.to(dev))
    train_den = bool(getattr(args, "train_denoiser", False))
    den = FFDNetColorDenoiser(seed=4, precision="bf16").to(dev).requires_grad_(train_den)
    params = [doe.height_map.height_map_sqrt, rhos, sigmas] + (list(den.model.parameters()) if train_den else [])
    opt = torch.optim.Adam(params, lr=1e-4 if not train_den else 1e-6)
    den._native = NativeFFDNet(den.model, dev)
    timer = _TimedDenoiser(den._native).install()
    x, y, PSF = dp.Variable(), dp.Placeholder(), dp.Placeholder()
    data_term = dp.sum_squares(dp.conv_doe(x, PSF, circular=True), y)
    reg_term = dp.deep_prior(x, denoiser=den, sqrt=True, trainable=train_den)
    solver = dp.specialize(dp.compile(data_term + reg_term, method="admm", device=dev), method="unroll", max_iter=T)
    gt_host = torch.rand(Bn, 3, H, H).pin_memory()
    gt_dev = gt_host.to(dev)
    ar_ev = []
    loss_host = torch.zeros(1).pin_memory()

    def train(gt):
        psf = doe.get_psf()
        inp = img_psf_conv(gt, psf, circular=True)
        inp = inp + (7.65 / 255) * torch.randn(gt.shape, device=dev, generator=g)
        y.value, PSF.value = inp, psf.detach()
        out = solver.solve(x0=inp, rhos=rhos, lams={reg_term: sigmas})
        loss = torch.nn.functional.mse_loss(gt, out)
        loss.backward()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ddist.allreduce_gradients(params)
        b.record()
        ar_ev.append((a, b))
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def step():
        return train(gt_dev)

    def e2e_step():
        loss = train(gt_host.to(dev, non_blocking=True))
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(args.warmup):
        step()
    ar_ev.clear()
    l0 = lib.dpx_launch_count()
    timer.on = True
    with B0.ClockSampler(D.local) as clk:
        ms = _timed(D, step, args.steps, 0)
        time.sleep(0.15)
    timer.on = False
    launches = lib.dpx_launch_count() - l0
    value = D.world * Bn * T * args.steps / (ms * 1e-3)
    ar_ms = sum(a.elapsed_time(b) for a, b in ar_ev) / max(1, len(ar_ev))
    e_steps = max(1, args.steps // 2)
    ems = _timed(D, e2e_step, e_steps, 1)
    e2e_value = D.world * Bn * T * e_steps / (ems * 1e-3)
    if D.rank == 0:
        _, peak_tf, src = _peaks()
        tot_ms, flop, n = timer.summary()
        # a backward call = the data gradient through all 12 layers: the same FLOPs as a forward
        achieved = (flop["fwd"] + flop["bwd"]) / ((tot_ms["fwd"] + tot_ms["bwd"]) * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": None,
                "peak_source": src,
                "unit_of_work": f"FFDNet-color forward / data-gradient calls of {Bn} x [3,{H},{H}] ({FLOP_PER_PIXEL * Bn * H * H / 1e12:.3f} TFLOP each; "
                                f"{n['fwd']} forward incl. activation recomputation + {n['bwd']} backward calls in the timed region), CUDA events per call",
                "denoiser_share_of_step": (tot_ms["fwd"] + tot_ms["bwd"]) / ms}
        cpu = None
        if not args.skip_cpu:
            cpu = _cfg5_cpu(H, psf_n, T, os.cpu_count() or 1)
        nbytes = sum(p.numel() for p in params) * 4
        dmode = "TRAINED: data + weight + bias gradients on the native kernels" if train_den else "frozen"
        wl = {"workload": f"cfg5: unrolled {T}-iteration ADMM (conv_doe + deep_prior(ffdnet_color, sqrt), bf16 tcgen05 denoiser, {dmode}) train step "
                          f"with DOE optics model, backward and Adam, {Bn} x [3,{H},{H}] per GPU, psf {psf_n}", "batch_per_gpu": Bn,
              "iters_per_step": T, "l2_policy": "denoiser activations (11 x 28 MB saved per call) and the 2244^2 Fresnel fields exceed L2",
              "parallelism": f"dp{D.world}: NCCL all-reduce of {nbytes / 1e6:.1f} MB of gradients per step ({ar_ms:.3f} ms on the launching stream)",
              "dtype": "bf16 (denoiser) / f32"}
        _line(args, D, "unrolled ADMM iters/sec with backward (train steps/s x batch x 10 iterations)", B0.UNIT, value, ms, wl, roof, cpu,
              {"value": e2e_value, "unit": B0.UNIT, "h2d_bytes_per_step": int(gt_host.numel() * 4), "d2h_bytes_per_step": 4, "steps": e_steps},
              launches, clk.summary(), {"train_steps_per_s": D.world * args.steps / (ms * 1e-3) / D.world, "images_per_s": value / T,
                                        "grad_allreduce_ms": ar_ms, "reference_published": "1.48-1.51 steps/s, bs=2, unstated GPU (BASELINE.md)"})
    D.close()


def _cfg5_cpu(H, psf_n, T, threads):
    """one unrolled train step (forward + backward) of ONE image on the host cores, reference or oracle port, fp32"""
    import dprox_oracle as orc
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(1, 3, H, H, generator=g)
    psf = torch.rand(1, 3, psf_n, psf_n, generator=g)
    psf = (psf / psf.sum(dim=(-2, -1), keepdim=True)).requires_grad_(True)
    ws = orc.ffdnet_random_weights(4)
    r0, s0 = orc.log_descent(49, 7.65, T, sigma=7.65 / 255)
    rhos, sigmas = r0.clone().requires_grad_(True), s0.clone().requires_grad_(True)
    t0 = time.perf_counter()
    inp = orc.ConvDOE(psf, orc.Identity()).fwd(gt) + (7.65 / 255) * torch.randn(gt.shape, generator=g)
    data = orc.Term("sum_squares", orc.ConvDOE(psf.detach(), orc.Identity()), b=inp)
    prior = orc.Term("deep_prior", denoiser=lambda v, s_: orc.ffdnet_forward(ws, v, s_), sqrt=True)
    out = orc.Solver([data, prior], "admm").solve(inp, rhos=rhos, lams={prior: sigmas}, max_iter=T)
    torch.nn.functional.mse_loss(gt, out).backward()
    dt = time.perf_counter() - t0
    return {"value": T / dt, "unit": B0.UNIT, "cores": threads, "kind": "port",
            "sample": f"1 train step of 1 image [3,{H},{H}] (forward + backward through {T} unrolled iterations, fp32), {dt:.1f} s, "
                      f"oracle port of the reference's op sequence under torch autograd, torch-CPU"}


# =====================================================================================================================

def run_reference(args):
    """`bench.py --impl reference --workload cfgN`: rank 0 only, the CPU arm's own JSON line"""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    threads = os.cpu_count() or 1
    size = args.size
    if args.workload == "cfg2":
        vals = [_cfg2_cpu(size, size, 2, threads) for _ in range(max(1, min(args.steps, 3)))]
        value, kind = float(np.mean([v[0] for v in vals])), vals[0][2]
        metric, sample = "ADMM iters/sec, 2Kx2K PnP deconv (deep_prior ffdnet_color)", f"1 problem [3,{size},{size}] x 2 iterations per step, fp32 FFDNet"
    elif args.workload == "cfg3":
        size = size if args.size_set else 256
        cpu = _cfg3_cpu(size, size, args.iters if args.iters_set else 30, 30, threads)
        value, kind, metric, sample = cpu["value"], cpu["kind"], "ADMM iters/sec, CS-MRI TV (PCG inner solve)", cpu["sample"]
    else:
        size = size if args.size_set else 768
        cpu = _cfg5_cpu(size, size - 20, args.iters if args.iters_set else 10, threads)
        value, kind, sample = cpu["value"], cpu["kind"], cpu["sample"]
        metric = "unrolled ADMM iters/sec with backward (train steps/s x batch x 10 iterations)"
    print(json.dumps({"impl": "reference", "metric": metric, "value": value, "unit": B0.UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": {"workload": args.workload, "reference_sample": sample},
                      "cpu_baseline": {"value": value, "unit": B0.UNIT, "cores": threads, "kind": kind, "sample": sample},
                      "e2e": {"value": value, "unit": B0.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run(args):
    if args.impl == "reference":
        return run_reference(args)
    {"cfg2": run_cfg2, "cfg3": run_cfg3, "cfg5": run_cfg5}[args.workload](args)
