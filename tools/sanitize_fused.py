#!/usr/bin/env python
"""Small fused-engine runs for compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box:
   compute-sanitizer --tool racecheck python tools/sanitize_fused.py
Covers the plane-pair engine (persistent TMA/cp.async row kernel, k_col), the half-spectrum engine and the TMA column kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delta-prox_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dprox_b200 as dp  # noqa: E402

psf = np.ones((5, 5, 1), "float32") / 25.0
# round 2: (2, 2048, 64) takes the TMA-staged column tile (bulk copies + transaction barrier), every pair case the tensor-map
# staging of the spectrum rows and the dynamic tile counter; (2, 1080, 64): radix-15 / radix-9 column passes
# closing round-2 cases (DPX_SANITIZE=new runs only these): 512-thread one-CTA-per-SM tiles (2560-point rows, 4096-point columns), a size
# added to the tile table (480 = 12*10*4), an odd batch (half-spectrum kernels) -- all launched as programmatic dependents
NEW = ((2, 64, 2560, "admm"), (2, 4096, 64, "admm"), (2, 480, 480, "hqs"), (1, 64, 128, "admm"))
ONLY_NEW = os.environ.get("DPX_SANITIZE") == "new"
OLD = ((2, 64, 128, "admm"), (1, 128, 64, "admm"), (2, 1024, 64, "hqs"), (2, 2048, 64, "admm"), (2, 1080, 64, "admm"))
for B, H, W, method in (NEW if ONLY_NEW else OLD + NEW):
    g = torch.Generator(device="cuda").manual_seed(B + H)
    b = torch.rand(B, 3, H, W, device="cuda", generator=g) - 0.3
    x = dp.Variable()
    s = dp.compile(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), method=method, device="cuda")
    out = s.solve(x0=b, rhos=1.0, lams=0.02, max_iter=4)
    torch.cuda.synchronize()
    print(B, H, W, method, float(out.abs().mean()))

if ONLY_NEW:
    sys.exit(0)

# fp32-class FFDNet: fp16-pair split convolution (tcgen05, two TMEM accumulators), forward + data gradient at a ragged size
from dprox_b200.denoisers import FFDNetColorDenoiser  # noqa: E402

den = FFDNetColorDenoiser(seed=4).cuda().requires_grad_(False)
xx = torch.rand(1, 3, 37, 70, device="cuda").requires_grad_(True)
y = den._denoise(xx, torch.tensor([0.05], device="cuda"))
y.sum().backward()
torch.cuda.synchronize()
print("ffdnet split", float(y.abs().mean()), float(xx.grad.abs().mean()))

# bf16 training path: forward_train, data gradient, MN-major tcgen05 weight gradient (ragged width: partial 128-pixel tile)
dent = FFDNetColorDenoiser(seed=4, precision="bf16").cuda()
xt = torch.rand(1, 3, 40, 150, device="cuda").requires_grad_(True)
yt = dent._denoise(xt, torch.tensor([0.05], device="cuda"))
yt.sum().backward()
torch.cuda.synchronize()
print("ffdnet train", float(yt.abs().mean()), float(sum(p.grad.abs().sum() for p in dent.model.parameters())))

# staged x-update path (external prox between the stages): persistent PM_FIRST / PM_XONLY row kernels
xv = dp.Variable()
bb = torch.rand(2, 3, 64, 128, device="cuda")
prior = dp.deep_prior(xv, denoiser=FFDNetColorDenoiser(seed=4, precision="bf16").cuda().requires_grad_(False))
sv = dp.compile(dp.sum_squares(dp.conv(xv, psf) - bb) + prior, method="admm", device="cuda")
with torch.no_grad():
    ov = sv.solve(x0=bb, rhos=1.0, lams=0.05, max_iter=2)
torch.cuda.synchronize()
print("pnp staged", float(ov.abs().mean()))
