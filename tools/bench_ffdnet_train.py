"""FFDNet-color training step (forward + backward incl. weight gradients) on the native tensor-core kernels vs the framework's
bf16 autocast / channels-last path (cuDNN).  Run under gpurun."""
import json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "delta-prox_b200"))
from dprox_b200.denoisers import FFDNetColorDenoiser

B, S = int(os.environ.get("B", 4)), int(os.environ.get("S", 1024))
x = torch.rand(B, 3, S, S, device="cuda")
sig = torch.full((B,), 0.05, device="cuda")
w = torch.rand(B, 3, S, S, device="cuda")
out = {"shape": [B, 3, S, S]}
for name, prec in (("native", "bf16"), ("torch_bf16_autocast", "torch")):
    den = FFDNetColorDenoiser(seed=4, precision=prec).cuda()
    if prec == "torch":
        den.model.to(memory_format=torch.channels_last)
    def step():
        for p in den.model.parameters():
            p.grad = None
        xa = x.clone().requires_grad_(True)
        if prec == "torch":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = den.model(xa.contiguous(memory_format=torch.channels_last), sig)
            (y.float() * w).sum().backward()
        else:
            (den._denoise(xa, sig) * w).sum().backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flop = 3 * 0.4255e6 * B * S * S          # forward + data gradient + weight gradient
    out[name] = {"ms_per_step": ms, "tflops": flop / (ms * 1e-3) / 1e12}
print(json.dumps(out))
