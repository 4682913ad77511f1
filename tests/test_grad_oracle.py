"""Gradient oracle (`-m "not gpu"`): autograd through the CPU restatement (oracle/dprox_oracle.py, plain torch ops) is
held to the gradients the UNMODIFIED reference produced by autograd through its unrolled loop
(tests/golden/unrolled_grads_*.npz, made by oracle/make_golden.py).  The GPU tests then hold the native backward
kernels to the same vectors."""
import os

import numpy as np
import torch

import dprox_oracle as orc
from conftest import GOLDEN


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def rel(a, b):
    a, b = np.asarray(a.detach() if isinstance(a, torch.Tensor) else a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_oracle_autograd_matches_reference_native_proxes():
    g = load("unrolled_grads_native")
    b = torch.from_numpy(g["b"]).requires_grad_(True)
    x0 = torch.from_numpy(g["x0"]).requires_grad_(True)
    rhos = torch.from_numpy(g["rhos"]).requires_grad_(True)
    lam1 = torch.from_numpy(g["lam1"]).requires_grad_(True)
    f1, f2 = orc.Term("norm1", alpha=0.5), orc.Term("nonneg")
    data = orc.Term("sum_squares", orc.Conv(g["psf"], orc.Identity()), c=b)
    out = orc.Solver([data, f1, f2], "admm").solve(x0, rhos=rhos, lams={f1: lam1, f2: torch.full((4,), 0.02)}, max_iter=int(g["T"]))
    (out * torch.from_numpy(g["wgt"])).sum().backward()
    assert rel(out, g["out"]) < 2e-6
    for name, t in (("g_b", b), ("g_x0", x0), ("g_rhos", rhos), ("g_lam1", lam1)):
        assert rel(t.grad, g[name]) < 2e-5, name


def test_oracle_autograd_matches_reference_unrolled_doe():
    g = load("unrolled_grads_doe")
    gt = torch.from_numpy(g["gt"])
    psf = torch.from_numpy(g["psf"]).requires_grad_(True)
    rhos = torch.from_numpy(g["rhos"]).requires_grad_(True)
    sigmas = torch.from_numpy(g["sigmas"]).requires_grad_(True)
    ws = orc.ffdnet_random_weights(int(g["seed"]))
    inp = orc.ConvDOE(psf, orc.Identity()).fwd(gt) + torch.from_numpy(g["noise"])       # data formation: carries d/d psf
    # inside the solver the PSF is a fresh leaf (conv_doe re-wraps the Placeholder value in nn.Parameter, conv.py:91-96)
    data = orc.Term("sum_squares", orc.ConvDOE(psf.detach(), orc.Identity()), b=inp)
    prior = orc.Term("deep_prior", denoiser=lambda v, s_: orc.ffdnet_forward(ws, v, s_), sqrt=True)
    out = orc.Solver([data, prior], "admm").solve(inp, rhos=rhos, lams={prior: sigmas}, max_iter=int(g["T"]))
    torch.nn.functional.mse_loss(gt, out).backward()
    assert rel(inp, g["inp"]) < 1e-6 and rel(out, g["out"]) < 1e-5
    assert rel(rhos.grad, g["g_rhos"]) < 1e-4 and rel(sigmas.grad, g["g_sigmas"]) < 1e-3 and rel(psf.grad, g["g_psf"]) < 1e-4


def test_xsolve_backward_closed_form_matches_autograd():
    """SURVEY App. D on the half spectrum (what `dpx_xsolve_backward` computes) vs torch autograd in fp64."""
    torch.manual_seed(0)
    dt = torch.float64
    B, C, H, W = 2, 3, 8, 10
    n, eps, m = H * W, 1e-7, 1.0
    t = torch.randn(B, C, H, W, dtype=dt, requires_grad=True)
    ktb = torch.randn(B, C, H, W, dtype=dt, requires_grad=True)
    rho = torch.tensor([0.7, 1.3], dtype=dt, requires_grad=True)
    O = torch.fft.rfft2(torch.rand(1, C, H, W, dtype=dt))
    dq = (O.conj() * O).real
    r = rho.view(-1, 1, 1, 1)
    x = torch.fft.irfft2((torch.fft.rfft2(ktb) + r * torch.fft.rfft2(t) + eps) / (dq + r * m + eps), s=(H, W))
    g = torch.randn_like(x)
    gt, gk, gr = torch.autograd.grad(x, (t, ktb, rho), g)
    with torch.no_grad():
        Dn = dq + r * m + eps
        Wh, Q = torch.fft.rfft2(g), torch.fft.rfft2(x)
        gk_m = torch.fft.irfft2(Wh / Dn, s=(H, W))
        wgt = torch.full((W // 2 + 1,), 2.0, dtype=dt)
        wgt[0] = 1.0
        wgt[-1] = 1.0
        gr_m = (wgt * (Wh.conj() * (Q * (dq + eps) - torch.fft.rfft2(ktb) - eps)).real / Dn / (n * r)).sum((1, 2, 3))
    assert (gk - gk_m).abs().max() < 1e-12 and (gt - r * gk_m).abs().max() < 1e-12 and (gr - gr_m).abs().max() < 1e-10


def test_oracle_implicit_cg_gradients_match_reference():
    """a13: LinearSolve's implicit differentiation through the CG x-update (joint demosaic + deconvolution), including the
    reference's omission of d(matrix)/d(rho) (see oracle._ImplicitSolve)."""
    g = load("unrolled_grads_cg")
    b = torch.from_numpy(g["b"]).requires_grad_(True)
    x0 = torch.from_numpy(g["x0"]).requires_grad_(True)
    rhos = torch.from_numpy(g["rhos"]).requires_grad_(True)
    f = orc.Term("nonneg")
    data = orc.Term("sum_squares", orc.Mosaic(orc.Conv(g["psf"], orc.Identity())), c=b)
    out = orc.Solver([data, f], "admm", solver_type="cg", rtol=float(g["rtol"]), max_iters=int(g["cg_iters"])).solve(
        x0, rhos=rhos, lams={f: torch.full((3,), 0.02)}, max_iter=int(g["T"]))
    (out * torch.from_numpy(g["wgt"])).sum().backward()
    assert rel(out, g["out"]) < 1e-5
    for name, t in (("g_b", b), ("g_x0", x0), ("g_rhos", rhos)):
        assert rel(t.grad, g[name]) < 1e-4, (name, rel(t.grad, g[name]))


def test_oracle_autograd_spatial_diag_unrolled():
    """unrolled training of a demosaicking objective (spatial-diagonal x-update)."""
    g = load("unrolled_grads_mosaic")
    b = torch.from_numpy(g["b"]).requires_grad_(True)
    x0 = torch.from_numpy(g["x0"]).requires_grad_(True)
    rhos = torch.from_numpy(g["rhos"]).requires_grad_(True)
    lam1 = torch.from_numpy(g["lam1"]).requires_grad_(True)
    f1 = orc.Term("norm1")
    data = orc.Term("sum_squares", orc.Mosaic(orc.Identity()), c=b)
    out = orc.Solver([data, f1], "admm").solve(x0, rhos=rhos, lams={f1: lam1}, max_iter=int(g["T"]))
    (out * torch.from_numpy(g["wgt"])).sum().backward()
    assert rel(out, g["out"]) < 2e-6
    for name, t in (("g_b", b), ("g_x0", x0), ("g_rhos", rhos), ("g_lam1", lam1)):
        assert rel(t.grad, g[name]) < 2e-5, name


def test_oracle_unrolled_share_false():
    """UnrolledSolver(share=False) (unroll.py:20-58): learned schedules, and per-iteration denoiser copies."""
    g = load("unrolled_share_false")
    b, wgt = torch.from_numpy(g["b"]), torch.from_numpy(g["wgt"])
    rhos = torch.tensor([0.6, 0.9, 1.3], requires_grad=True)
    lam = torch.tensor([0.05, 0.03, 0.02], requires_grad=True)
    f1 = orc.Term("norm1")
    data = orc.Term("sum_squares", orc.Conv(g["psf"], orc.Identity()), c=b)
    out = orc.Solver([data, f1], "admm").solve(b, rhos=rhos, lams={f1: lam}, max_iter=3)
    (out * wgt).sum().backward()
    assert rel(out, g["lp_out"]) < 2e-6 and rel(rhos.grad, g["lp_g_rhos"]) < 2e-5 and rel(lam.grad, g["lp_g_lam"]) < 2e-5
    copies = [[(w.clone().requires_grad_(True), bb.clone().requires_grad_(True)) for w, bb in orc.ffdnet_random_weights(int(g["seed"]))]
              for _ in range(3)]
    calls = []

    def den(v, s_):
        calls.append(1)
        return orc.ffdnet_forward(copies[len(calls) - 1], v, s_)

    prior = orc.Term("deep_prior", denoiser=den)
    data = orc.Term("sum_squares", orc.Conv(g["psf"], orc.Identity()), c=b)
    out = orc.Solver([data, prior], "admm").solve(b, rhos=torch.tensor([0.6, 0.9, 1.3]), lams={prior: torch.from_numpy(g["sig"])},
                                                  max_iter=3)
    (out * wgt).sum().backward()
    assert rel(out, g["dn_out"]) < 1e-5
    assert rel(copies[0][0][0].grad, g["dn_g_w0"]) < 1e-3 and rel(copies[1][0][0].grad, g["dn_g_w1"]) < 1e-3
