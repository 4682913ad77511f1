"""GPU check of the hand-written tcgen05 convolution: per-layer parity (forward + data gradient) against torch conv2d on
bf16-rounded operands, whole network vs the fp32 torch network, backward vs autograd, and timing."""
import os, sys, time
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "delta-prox_b200"), os.path.join(ROOT, "oracle")]
from dprox_b200.denoisers import FFDNetColorDenoiser, NativeFFDNet
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

den = FFDNetColorDenoiser(seed=4).cuda()
net = NativeFFDNet(den.model, torch.device("cuda"))
convs = [m for m in den.model.model if isinstance(m, torch.nn.Conv2d)]
bf = lambda t: t.to(torch.bfloat16).float()
g = torch.Generator(device="cuda").manual_seed(0)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "layers"):
    for (B, H, W) in ((1, 16, 24), (2, 37, 300), (1, 64, 512)):
        for layer in (0, 1, 11):
            c = convs[layer]
            x = torch.randn(B, c.in_channels, H, W, device="cuda", generator=g)
            y = net.conv_layer(layer, x, 0, relu=(layer != 11))
            ref = F.conv2d(bf(x), bf(c.weight), c.bias, padding=1)
            if layer != 11:
                ref = ref.relu()
            print(f"fwd  layer {layer:2d} [{B},{c.in_channels},{H},{W}] rel {rel(y, ref):.3e}  max|d| {float((y-ref).abs().max()):.3e}", flush=True)
            gy = torch.randn(B, c.out_channels, H, W, device="cuda", generator=g)
            gx = net.conv_layer(layer, gy, 1)
            refg = F.conv_transpose2d(bf(gy), bf(c.weight), padding=1)
            print(f"dgrad layer {layer:2d} rel {rel(gx, refg):.3e}", flush=True)
if what in ("all", "net"):
    for shape in ((2, 3, 64, 96), (1, 3, 45, 70), (1, 3, 512, 768)):
        x = torch.rand(*shape, device="cuda", generator=g)
        sig = (0.02 + 0.1 * torch.rand(shape[0], device="cuda", generator=g))
        fast = FFDNetColorDenoiser(seed=4, precision="bf16").cuda().requires_grad_(False)
        y = fast.denoise(x, sig)
        yr = den.denoise(x, sig)
        print(f"net {shape} rel vs fp32 torch {rel(y, yr):.3e}", flush=True)
        xg = x.clone().requires_grad_(True)
        sg = sig.clone().requires_grad_(True)
        w = torch.rand(*shape, device="cuda", generator=g)
        (fast.denoise(xg, sg) * w).sum().backward()
        xr = x.clone().requires_grad_(True)
        sr = sig.clone().requires_grad_(True)
        (den.denoise(xr, sr) * w).sum().backward()
        print(f"    backward: g_x rel {rel(xg.grad, xr.grad):.3e}  g_sigma rel {rel(sg.grad, sr.grad):.3e}", flush=True)
        xa = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ya = den.model(xa, sig)
        (ya.float() * w).sum().backward()
        print(f"    (torch bf16 autocast backward vs fp32: g_x rel {rel(xa.grad, xr.grad):.3e})", flush=True)
if what in ("all", "time"):
    fast = FFDNetColorDenoiser(seed=4, precision="bf16").cuda().requires_grad_(False)
    x = torch.rand(2, 3, 2048, 2048, device="cuda", generator=g)
    sig = torch.tensor([0.05, 0.1], device="cuda")
    for _ in range(3):
        fast.denoise(x, sig)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        fast.denoise(x, sig)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flop = 2 * 851040 * (1024 * 1024) * 2
    print(f"FFDNet 2x[3,2048,2048]: {ms:.3f} ms per call = {flop / ms / 1e9:.0f} TFLOP/s", flush=True)
