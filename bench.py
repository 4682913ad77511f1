#!/usr/bin/env python
"""bench.py — ADMM iterations/sec on a batch of 2K x 2K deconvolution problems (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--batch B_per_gpu] [--iters T] [--size 2048] [--fft-backend 0|1|2]

Workload (SURVEY.md §8d "headline"): per GPU, B problems [3,2048,2048] fp32,
`sum_squares(conv(x, psf) - b) + nonneg(x)`, Gaussian PSF 15x15 sigma 5, x0 = b, rho = 1, lam = 0.02, ADMM.
One *step* = one pass of the hot path over the batch = T ADMM iterations of all B problems (one `dpx_iters`
call).  `value` = problem-iterations / second over all GPUs with every input already resident in HBM;
`e2e` = the same metric through the public API with HOST (pinned) measurements: per step the H2D copy of b,
the constant hoisting (K^T b, its FFT), state init, T iterations and the D2H copy of x are inside the timing.

N > 1: one process per GPU under torchrun, independent problem shards, no data-path collective ("weak").
`--impl reference` times the reference's own CPU implementation on the host cores of this box on a bounded sample
of the same workload: the UNMODIFIED reference package when it is installed under baseline/_ref (`pip install --target
baseline/_ref /root/reference`, git-ignored), else the oracle port of its op sequence (oracle/dprox_oracle.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "delta-prox_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

ALG_BYTES_PER_ELEM = 24.0          # ADMM, one identity prox term: read v,u,F(b); write x,v,u (fp32)
METRIC = "ADMM iters/sec, 2Kx2K deconv batch"
UNIT = "problem-iters/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def psf_gaussian(k=15, sigma=5.0):
    r = (k - 1) / 2.0
    ax = np.arange(-r, r + 1)
    xx, yy = np.meshgrid(ax, ax)
    h = np.exp(-(xx * xx + yy * yy) / (2 * sigma * sigma))
    h[h < np.finfo(float).eps * h.max()] = 0
    return (h / h.sum())[..., None].astype("float32")


def make_measurements(B, C, H, W, seed, device):
    """b = blur(img) + 0.01 noise, synthesised on the device with a seeded generator (img ~ U[-0.3, 0.7))."""
    g = torch.Generator(device=device).manual_seed(seed)
    img = torch.rand(B, C, H, W, device=device, generator=g) - 0.3
    noise = 0.01 * torch.randn(B, C, H, W, device=device, generator=g)
    return img, noise


def workload_config(args, world):
    """Identical for both arms: the reference arm runs a bounded SAMPLE of this workload (see cpu_baseline.sample)."""
    B, H, W, T = args.batch, args.size, args.size, args.iters
    N = B * 3 * H * W
    mib = N * 4 / 2**20
    l2 = (f"state arrays are {mib:.0f} MiB each (> 126 MB L2) and every iteration streams all of them" if mib > 126 else
          f"state arrays are {mib:.1f} MiB each: the working set FITS the 126 MB L2 (the reference's own problem size; no flush between "
          f"iterations -- this configuration is launch-latency bound and its roofline fraction is reported for completeness only)")
    return {"workload": f"{args.method} deconv+nonneg, {B} problems/GPU [3,{H},{W}] fp32, psf gaussian 15/5, rho=1, lam=0.02, "
                        f"{T} iterations per step", "batch_per_gpu": B, "iters_per_step": T,
            "l2_policy": l2,
            "fft_backend": {0: "auto", 1: "cufft", 2: "fused"}[args.fft_backend],
            "parallelism": f"dp{world} (problem shards, no collective)"}


def pin_to_gpu_numa_node(index):
    """Run this rank on the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the end-to-end leg
    are allocated on that NUMA node (eight ranks otherwise cross the socket interconnect for half of their copies)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:                                      # noqa: BLE001 -- affinity is an optimisation, never a requirement
        pass


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"], "samples": 0}
        # "under load" = the upper half of the samples (the sampler also sees the idle edges)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
#  CPU arm: the oracle port timed on the host cores
# ------------------------------------------------------------------------------------------------

REF_DIR = os.path.join(ROOT, "baseline", "_ref")        # `pip install --target baseline/_ref /root/reference` (git-ignored, ships to the box)


def _reference_module():
    """The UNMODIFIED reference package if it was installed next to the repo, else None."""
    if not os.path.isdir(os.path.join(REF_DIR, "dprox")):
        return None
    os.environ["DPROX_REFERENCE_ROOT"] = REF_DIR
    try:
        import refshim                                   # stubs for import-time-only third-party modules (oracle/refshim.py)
        refshim.REFERENCE_ROOT = REF_DIR
        return refshim.import_reference()
    except Exception as e:                               # noqa: BLE001
        print(f"[bench] reference import failed ({type(e).__name__}: {e}); falling back to the oracle port", file=sys.stderr)
        return None


def cpu_rate(H, W, iters, threads, seed=0):
    """problem-iterations/s of the reference path on the host cores, ONE [3,H,W] problem: the reference's own
    `compile(...).solve(...)` (kind 'reference') when baseline/_ref holds it, else the oracle port of its op sequence."""
    import dprox_oracle as orc
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(1, 3, H, W, generator=g) - 0.3
    psf = orc.point_spread_function(15, 5)
    b = orc.Conv(psf, orc.Identity()).fwd(img) + 0.01 * torch.randn(1, 3, H, W, generator=g)
    ref = _reference_module()
    if ref is not None:
        x = ref.Variable()
        solver = ref.compile(ref.sum_squares(ref.conv(x, psf) - b) + ref.nonneg(x), method="admm", device="cpu")
        run = lambda n: solver.solve(x0=b, rhos=1.0, lams=0.02, max_iter=n)
        kind = "reference"
    else:
        solver = orc.Solver([orc.Term("sum_squares", orc.Conv(psf, orc.Identity()), c=b), orc.Term("nonneg")], "admm")
        run = lambda n: solver.solve(b, rhos=1.0, lams=0.02, max_iter=n)
        kind = "port"
    with torch.no_grad():
        run(2)                                           # warm-up (FFT plans, OTF cache)
        t0 = time.perf_counter()
        run(iters)
        dt = time.perf_counter() - t0
    return iters / dt, dt, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    H = W = args.size
    iters = args.ref_iters
    rates = []
    kind = "port"
    for _ in range(args.warmup):
        cpu_rate(H, W, 1, threads)
    t_all = 0.0
    for _ in range(args.steps):
        r, dt, kind = cpu_rate(H, W, iters, threads)
        rates.append(r)
        t_all += dt
    value = float(np.mean(rates))
    what = "the unmodified reference (baseline/_ref)" if kind == "reference" else "oracle port of the reference's op sequence"
    sample = f"1 problem [3,{H},{W}] x {iters} ADMM iterations per step, {args.steps} steps, {what}, torch-CPU {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, args.gpus), reference_sample=sample,
                       note="the CPU arm times a bounded SAMPLE of the workload (one problem, fewer iterations per step); the rate "
                            "is normalised to problem-iterations, and a batch of 1 is the CPU's best case (BASELINE.md §2)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
#  GPU arm
# ------------------------------------------------------------------------------------------------

def run_native(args):
    import torch.distributed as dist
    import dprox_b200 as dp
    from dprox_b200 import _cabi as cabi

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pin_to_gpu_numa_node(local)                            # before the pinned host buffers are allocated
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = cabi.lib()

    B, C, H, W, T = args.batch, 3, args.size, args.size, args.iters
    N = B * C * H * W
    psf = psf_gaussian(15, 5.0)

    # ---- problem batch of this rank: Placeholder-fed measurements so that a new batch re-uses the plan ----
    x = dp.Variable()
    y = dp.Placeholder()
    data_op = dp.conv(x, psf)
    solver = dp.compile(dp.sum_squares(dp.conv(x, psf) - y) + dp.nonneg(x), method=args.method, device=dev,
                        fft_backend=args.fft_backend)
    img, noise = make_measurements(B, C, H, W, seed=1234 + rank, device=dev)
    b_dev = data_op.to(dev).forward(img)
    b_dev.add_(noise)
    del img, noise
    y.value = b_dev
    b_host = b_dev.cpu().pin_memory()
    out_host = torch.empty_like(b_host).pin_memory()

    rhos = torch.full((T,), 1.0, device=dev)               # schedules are solver constants: resident on the device, so that
    lams = torch.full((T,), 0.02, device=dev)              # solve() never blocks the host on a pageable copy

    def resident_step(state):
        return solver.iters(state, rhos, lams, T)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- resident-input throughput (`value`) -------------------------------------------------------
    state = solver.initialize(b_dev)
    for _ in range(args.warmup):
        state = resident_step(state)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.dpx_launch_count()
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            state = resident_step(state)
        ev1.record()
        barrier()
        time.sleep(0.15)
    launches = lib.dpx_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    ms_t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t)
    total_units = world * B * T * args.steps
    value = total_units / (ms_max * 1e-3)

    # ---- end-to-end through the public API with host buffers (`e2e`) ---------------------------------
    # The batch is fed as `n_chunks` sub-batches, each with its own compiled solver (public API) on its own CUDA
    # stream, so the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap the iterations of chunk k.
    del state
    if args.skip_e2e:                                        # experiment mode only: prints the resident-input line and stops
        if rank == 0:
            it_ms = ms_max / (T * args.steps)
            print(json.dumps({"experiment": True, "value": value, "ms_per_iteration": it_ms, "gpu_launches": int(launches),
                              "frac": (16.0 if args.method == "hqs" else ALG_BYTES_PER_ELEM) * N / (it_ms * 1e-3) / 1e9 / load_peaks()[0],
                              "clocks": clk.summary()}))
        if world > 1:
            dist.destroy_process_group()
        return
    n_chunks = max(1, min(args.e2e_chunks, B))
    while B % n_chunks:
        n_chunks -= 1
    Bc = B // n_chunks
    if n_chunks > 1:
        del solver
        torch.cuda.empty_cache()
    chunks = []
    for c in range(n_chunks):
        xc, yc = dp.Variable(), dp.Placeholder()
        sc = solver if n_chunks == 1 else dp.compile(dp.sum_squares(dp.conv(xc, psf) - yc) + dp.nonneg(xc), method=args.method,
                                                     device=dev, fft_backend=args.fft_backend)
        # the sub-batches run concurrently (their kernels fill each other's tails); strictly ordering them, or giving
        # earlier ones a higher stream priority so that their D2H copy starts sooner, measured slower / no different
        chunks.append((sc, y if n_chunks == 1 else yc, b_host[c * Bc:(c + 1) * Bc], out_host[c * Bc:(c + 1) * Bc],
                       torch.cuda.Stream(device=dev, priority=max(-5, c - n_chunks + 1) if args.e2e_priority else 0)))

    def e2e_step():
        # Each sub-batch owns a stream and its work is ordered on it (H2D -> constants -> iterations -> D2H), so consecutive
        # steps pipeline: the H2D copy of step k+1 runs under the iterations / D2H copies of step k.  `e2e_join()` closes
        # the timed region.
        done = []
        for k, (sc, yc, bh, oh, st) in enumerate(chunks):
            with torch.cuda.stream(st):
                bd = bh.to(dev, non_blocking=True)             # H2D of this chunk's measurements (pinned)
                if args.e2e_wave and k >= args.e2e_wave:
                    st.wait_event(done[k - args.e2e_wave])     # at most `e2e_wave` sub-batches iterate concurrently
                yc.value = bd                                  # new measurements -> F(K^T b) re-hoisted on the same plan
                xs = sc.solve(x0=bd, rhos=rhos, lams=lams, max_iter=T)
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
                oh.copy_(xs, non_blocking=True)                # D2H of the result
                bd.record_stream(st)
                xs.record_stream(st)

    def e2e_join():
        main = torch.cuda.current_stream(dev)
        for *_, st in chunks:
            main.wait_stream(st)

    for _ in range(max(2, args.warmup // 2)):
        e2e_step()
    e2e_join()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_steps = max(1, args.steps // 2)
    e0.record()
    for _ in range(e_steps):
        e2e_step()
    e2e_join()
    e1.record()
    barrier()
    ems_t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ems_t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * T * e_steps / (float(ems_t) * 1e-3)
    e2e_ms_per_step = float(ems_t) / e_steps

    # ---- the path's only collective: the opt-in residual stop rule, all-reduced over the ranks -------------
    # per-sample sums reduced on the device, NCCL all-reduce + host copy on a side stream, decision consumed one check later
    stop_leg = None
    if args.stop_leg:
        sc0, yc0 = chunks[0][0], chunks[0][1]
        bd0 = b_host[:Bc].to(dev)
        yc0.value = bd0
        stop = dp.ResidualStop(abstol=1e-4, reltol=1e-3, every=10)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        sc0.solve(x0=bd0, rhos=1.0, lams=0.02, max_iter=200, stop=stop)
        s1.record()
        barrier()
        its = torch.tensor([sc0.iterations_run], device=dev)
        gathered = [torch.zeros_like(its) for _ in range(world)]
        if world > 1:
            dist.all_gather(gathered, its)
        else:
            gathered = [its]
        stop_leg = {"iterations_run": int(its), "same_on_every_rank": len({int(t) for t in gathered}) == 1, "checks": len(stop.history),
                    "every": 10, "ms": s0.elapsed_time(s1), "ranks": world,
                    "how": "dpx_resid_reduce on the device, all-reduce of 5 doubles over NCCL on a side stream, decision consumed one check late"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the fused iteration + CPU baseline (rank 0) ----------------------------------------
    peak, peak_src = load_peaks()
    it_ms = ms_max / (T * args.steps)                      # average duration of one iteration of the whole batch
    alg_bytes = 16.0 if args.method == "hqs" else ALG_BYTES_PER_ELEM      # HQS has no dual variable
    achieved = alg_bytes * N / (it_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "peak_source": peak_src,
            "unit_of_work": f"one {args.method.upper()} iteration of {B} problems = {alg_bytes:.0f} B x {N} elements",
            "avg_iteration_ms": it_ms}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            per_elem = json.load(open(prof)).get("dram_bytes_per_element_per_iteration")
            roof["traffic"] = None if per_elem is None else per_elem * N      # DRAM bytes per iteration of this batch (ncu)
            roof["traffic_source"] = "profiles/traffic.json (ncu --set full, fused engine)"
        except Exception:
            pass
    cpu = None
    if not args.skip_cpu:
        threads = os.cpu_count() or 1
        n_cpu = max(args.ref_iters, 60)                       # ~10-20 s of CPU work
        r, dt, kind = cpu_rate(H, W, n_cpu, threads)
        cpu = {"value": r, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"1 problem [3,{H},{W}] x {n_cpu} ADMM iterations ({dt:.1f} s), "
                         f"{'unmodified reference (baseline/_ref)' if kind == 'reference' else 'oracle port'}, torch-CPU"}
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(N * 4), "d2h_bytes_per_step": int(N * 4),
                "steps": e_steps, "ms_per_step": e2e_ms_per_step,
                "pipeline": f"{n_chunks} sub-batches of {Bc} problems on {n_chunks} CUDA streams "
                            f"(H2D / iterations / D2H overlapped, consecutive steps pipelined)"},
        "stop_leg": stop_leg,
        "gpu_launches": int(launches), "clocks": clk.summary(),
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="problems per GPU")
    ap.add_argument("--iters", type=int, default=50, help="ADMM iterations per step")
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--fft-backend", type=int, default=0)
    ap.add_argument("--ref-iters", type=int, default=12, help="CPU-arm iterations per step (reference arm)")
    ap.add_argument("--method", default="admm", choices=["admm", "hqs"], help="hqs + --size 1024 --iters 24 = BASELINE config 4 per GPU")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="experiments: resident-input number only (not a bench line)")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="sub-batches (streams) of the end-to-end pipeline")
    ap.add_argument("--e2e-wave", type=int, default=0, help="max sub-batches iterating concurrently (0 = all)")
    ap.add_argument("--e2e-priority", type=int, default=0, help="1: earlier sub-batches on higher-priority streams (measured: no gain)")
    ap.add_argument("--stop-leg", type=int, default=1, help="0: skip the residual-stop leg (the path's only collective, untimed extra)")
    ap.add_argument("--denoiser", default="fp32", choices=["fp32", "bf16"],
                    help="cfg2: FFDNet precision on the tensor cores: fp32 = fp16 operand pairs (1e-5 parity bar), bf16 = fast mode")
    ap.add_argument("--train-denoiser", action="store_true",
                    help="cfg5: the FFDNet weights are trained too (native weight-gradient kernel) instead of frozen")
    ap.add_argument("--workload", default="headline", choices=["headline", "cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configs: headline = configs[1]-shape batch (default); cfg1 = single 256x256 ADMM x 50; cfg2 = PnP "
                         "deconv with the deep denoiser; cfg3 = CS-MRI + TV with PCG; cfg4 = HQS 8 x 1024^2 per GPU; cfg5 = unrolled training step")
    args = ap.parse_args()
    given = set(a.split("=")[0] for a in sys.argv[1:] if a.startswith("--"))
    args.batch_set, args.iters_set, args.size_set = "--batch" in given, "--iters" in given, "--size" in given
    if args.workload == "cfg1":                            # BASELINE configs[0]: single [3,256,256] image, ADMM x 50
        args.batch = args.batch if args.batch_set else 1
        args.size = args.size if args.size_set else 256
    elif args.workload == "cfg4":                          # BASELINE configs[3]: 64 x [3,1024,1024] over 8 GPUs, HQS x 24
        args.method = "hqs"
        args.batch = args.batch if args.batch_set else 8
        args.size = args.size if args.size_set else 1024
        args.iters = args.iters if args.iters_set else 24
    if args.workload in ("cfg2", "cfg3", "cfg5"):
        import bench_workloads
        bench_workloads.run(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
