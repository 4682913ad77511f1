"""Where one graph-replayed PCG solve of BASELINE config 3 spends its time (run under gpurun)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "delta-prox_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dprox_b200 as dp
from dprox_b200 import linalg, ops
import bench_workloads as BW

dev = torch.device("cuda", 0)
img, mask = BW._cfg3_problem(8, 256, 256, 7, dev)
fwd = lambda x, step=0: mask * torch.fft.fft2(x, norm="ortho")
adj = lambda y, step=0: torch.real(torch.fft.ifft2(mask * y, norm="ortho")).contiguous()
A = lambda x: ops.axpby(1.0, adj(fwd(x)), 1.0, x)
b = adj(fwd(img))

def t(f, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

for g in ("0", "1"):
    os.environ["DPX_CG_GRAPH"] = g
    print("graph", g, "pcg 30 steps: %.2f ms" % t(lambda: linalg.pcg(A, b, rtol=1e-6, max_iters=30)))
os.environ["DPX_CG_GRAPH"] = "1"
orig = linalg._capture
def timed_capture(body):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    gph = orig(body)
    torch.cuda.synchronize(); print("   capture+instantiate %.2f ms" % ((time.perf_counter() - t0) * 1e3), "ok" if gph is not None else "FAILED")
    if gph is not None:
        print("   replay %.3f ms" % t(gph.replay, 20))
    return gph
linalg._capture = timed_capture
linalg.pcg(A, b, rtol=1e-6, max_iters=30)
print("one eager operator application: %.3f ms" % t(lambda: A(b), 20))
