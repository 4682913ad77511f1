#!/usr/bin/env python
"""Per-chunk timeline (CUDA events) of bench.py's end-to-end pipeline: when does each sub-batch finish its H2D copy,
its constants, its iterations and its D2H copy, relative to the start of the step.  Diagnostic only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delta-prox_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dprox_b200 as dp  # noqa: E402
from bench import psf_gaussian  # noqa: E402

dev = torch.device("cuda", 0)
B, T, n_chunks = 8, 50, int(sys.argv[1]) if len(sys.argv) > 1 else 4
Bc = B // n_chunks
psf = psf_gaussian(15, 5.0)
b_host = (torch.rand(B, 3, 2048, 2048) - 0.3).pin_memory()
out_host = torch.empty_like(b_host).pin_memory()
rhos, lams = torch.full((T,), 1.0, device=dev), torch.full((T,), 0.02, device=dev)
chunks = []
for c in range(n_chunks):
    xc, yc = dp.Variable(), dp.Placeholder()
    sc = dp.compile(dp.sum_squares(dp.conv(xc, psf) - yc) + dp.nonneg(xc), method="admm", device=dev)
    chunks.append((sc, yc, b_host[c * Bc:(c + 1) * Bc], out_host[c * Bc:(c + 1) * Bc], torch.cuda.Stream(device=dev)))


def step(trace=None):
    main = torch.cuda.current_stream(dev)
    start = torch.cuda.Event(enable_timing=True)
    start.record(main)
    t0 = time.perf_counter()
    for i, (sc, yc, bh, oh, st) in enumerate(chunks):
        st.wait_event(start)
        with torch.cuda.stream(st):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            bd = bh.to(dev, non_blocking=True)
            ev[0].record(st)
            h0 = time.perf_counter() - t0
            yc.value = bd
            xs = sc.solve(x0=bd, rhos=rhos, lams=lams, max_iter=T)
            ev[2].record(st)
            h1 = time.perf_counter() - t0
            oh.copy_(xs, non_blocking=True)
            ev[3].record(st)
            bd.record_stream(st)
            xs.record_stream(st)
        main.wait_stream(st)
        if trace is not None:
            trace.append((i, start, ev, h0, h1))
    return start


for _ in range(3):
    step()
torch.cuda.synchronize()
tr = []
step(tr)
torch.cuda.synchronize()
for i, start, ev, h0, h1 in tr:
    print(f"chunk {i}: H2D done {start.elapsed_time(ev[0]):7.2f} ms | iterations done {start.elapsed_time(ev[2]):7.2f} | D2H done "
          f"{start.elapsed_time(ev[3]):7.2f} | host enqueued H2D at {1e3 * h0:6.2f} ms, solve returned at {1e3 * h1:6.2f} ms")
