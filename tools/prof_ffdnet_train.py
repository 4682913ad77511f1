import os, sys, time, torch
sys.path.insert(0, "/root/repo/delta-prox_b200")
from dprox_b200.denoisers import FFDNetColorDenoiser
import torch.nn as nn
den = FFDNetColorDenoiser(seed=4, precision="bf16").cuda()
B, S = 2, 2048
x = torch.rand(B, 3, S, S, device="cuda"); sig = torch.full((B,), 0.05, device="cuda"); g = torch.rand(B, 3, S, S, device="cuda")
net = den._native_net(x.device)
convs = [m for m in den.model.model if isinstance(m, nn.Conv2d)]
shapes = [(tuple(c.weight.shape), tuple(c.bias.shape)) for c in convs]
def t(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("load_weights %.2f ms" % t(lambda: net.load_weights(convs)))
print("forward_train %.2f ms" % t(lambda: net(x, sig, train=True)))
def fb():
    net(x, sig, train=True); net.backward_params(g, 2, shapes)
print("fwd+backward_params %.2f ms" % t(fb))
def fb2():
    net(x, sig, train=True); net.backward(g, 2)
print("fwd+backward(data only) %.2f ms" % t(fb2))
