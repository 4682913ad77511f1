"""GPU gradient parity (`pytest -m gpu`): the differentiable engine — native forward AND backward kernels under a
torch tape (dprox_b200/autograd.py, `dpx_xsolve_backward`, `dpx_prox_backward`) — against the gradients the unmodified
reference produced by autograd through its unrolled loop (tests/golden/unrolled_grads_*.npz) and against torch
autograd of the CPU oracle.  Tolerances: 1e-5 on the primal output, 1e-4 relative L2 on gradients (fp32 through
3-4 unrolled iterations; the reference's own fp32 gradients sit ~1e-5 from an fp64 evaluation)."""
import os

import numpy as np
import pytest
import torch

import dprox_oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dp():
    import dprox_b200
    from dprox_b200 import _cabi
    _cabi.lib()
    return dprox_b200


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def T(a, dev="cuda"):
    return torch.from_numpy(np.asarray(a)).to(dev)


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_unrolled_gradients_native_proxes(dp):
    g = load("unrolled_grads_native")
    b = T(g["b"]).requires_grad_(True)
    x0 = T(g["x0"]).requires_grad_(True)
    rhos = T(g["rhos"]).requires_grad_(True)
    lam1 = T(g["lam1"]).requires_grad_(True)
    x = dp.Variable()
    f1, f2 = 0.5 * dp.norm1(x), dp.nonneg(x)
    solver = dp.compile(dp.sum_squares(dp.conv(x, g["psf"]) - b) + f1 + f2, method="admm", device="cuda")
    out = solver.solve(x0=x0, rhos=rhos, lams={f1: lam1, f2: torch.full((4,), 0.02)}, max_iter=int(g["T"]))
    assert out.requires_grad
    (out * T(g["wgt"])).sum().backward()
    assert rel(out, g["out"]) < 1e-5
    for name, t in (("g_b", b), ("g_x0", x0), ("g_rhos", rhos), ("g_lam1", lam1)):
        assert t.grad is not None, name
        assert rel(t.grad, g[name]) < 1e-4, (name, rel(t.grad, g[name]))
    # the same call under no_grad takes the fused forward-only loop and agrees
    with torch.no_grad():
        out2 = solver.solve(x0=x0, rhos=rhos, lams={f1: lam1, f2: torch.full((4,), 0.02)}, max_iter=int(g["T"]))
    assert not out2.requires_grad and rel(out2, out) < 1e-5


def test_unrolled_gradients_doe_deep_prior(dp):
    """BASELINE config 5 in miniature: conv_doe + deep_prior(FFDNet, sqrt=True), specialize('unroll'), mse loss."""
    from dprox_b200.denoisers import FFDNetColorDenoiser
    g = load("unrolled_grads_doe")
    gt = T(g["gt"])
    psf = T(g["psf"]).requires_grad_(True)
    rhos = T(g["rhos"]).requires_grad_(True)
    sigmas = T(g["sigmas"]).requires_grad_(True)
    den = FFDNetColorDenoiser(seed=int(g["seed"])).cuda()
    x, y, PSF = dp.Variable(), dp.Placeholder(), dp.Placeholder()
    data_term = dp.sum_squares(dp.conv_doe(x, PSF, circular=True), y)
    reg_term = dp.deep_prior(x, denoiser=den, sqrt=True)
    solver = dp.specialize(dp.compile(data_term + reg_term, method="admm", device="cuda"), method="unroll", max_iter=int(g["T"]))
    # data formation outside the solver (img_psf_conv in the reference's trainer): plain torch, carries d/d psf
    from dprox_b200.linop import psf2otf2
    otf = psf2otf2(psf, gt.shape)
    inp = torch.real(torch.fft.ifftn(otf * torch.fft.fftn(gt, dim=[-2, -1]), dim=[-2, -1])).float() + T(g["noise"])
    y.value = inp
    PSF.value = psf.detach()
    out = solver.solve(x0=inp, rhos=rhos, lams={reg_term: sigmas})
    loss = torch.nn.functional.mse_loss(gt, out)
    loss.backward()
    assert rel(out, g["out"]) < 2e-5 and abs(float(loss.detach()) - float(g["loss"])) < 1e-5 * float(g["loss"]) + 1e-7
    assert rel(rhos.grad, g["g_rhos"]) < 1e-3, rel(rhos.grad, g["g_rhos"])
    assert rel(sigmas.grad, g["g_sigmas"]) < 5e-3, rel(sigmas.grad, g["g_sigmas"])
    assert rel(psf.grad, g["g_psf"]) < 1e-3, rel(psf.grad, g["g_psf"])
    # second training step with new Placeholder values re-uses the plan and builds a fresh tape
    y.value = inp.detach() * 1.01
    out2 = solver.solve(x0=inp.detach(), rhos=rhos, lams={reg_term: sigmas})
    out2.sum().backward()


@pytest.mark.parametrize("method", ["hqs", "admm"])
def test_gradients_vs_oracle_autograd_tv(dp, method):
    """Non-identity psi linops (anisotropic TV through the stencil gradient) + per-sample schedules, vs oracle autograd."""
    gen = torch.Generator().manual_seed(5)
    B, C, H, W, Tn = 2, 1, 16, 24, 3
    img = torch.rand(B, C, H, W, generator=gen)
    psf = orc.point_spread_function(5, 1.5)
    b0 = orc.Conv(psf, orc.Identity()).fwd(img) + 0.01 * torch.randn(B, C, H, W, generator=gen)
    wgt = torch.rand(B, C, H, W, generator=gen)
    rho0 = 0.5 + torch.rand(B, Tn, generator=gen)
    lam0 = 0.02 + 0.03 * torch.rand(Tn, generator=gen)

    def run(make, dev):
        b = b0.to(dev).requires_grad_(True)
        rho = rho0.to(dev).requires_grad_(True)
        lam = lam0.to(dev).requires_grad_(True)
        out = make(b, rho, lam)
        (out * wgt.to(dev)).sum().backward()
        return out, b.grad, rho.grad, lam.grad

    def ours(b, rho, lam):
        x = dp.Variable()
        th, tw = dp.norm1(dp.grad(x, dim=0)), dp.norm1(dp.grad(x, dim=1))
        s = dp.compile(dp.sum_squares(dp.conv(x, psf) - b) + th + tw, method=method, device="cuda")
        return s.solve(x0=b, rhos=rho, lams={th: lam, tw: lam}, max_iter=Tn)

    def oracle(b, rho, lam):
        th, tw = orc.Term("norm1", orc.Grad(0, orc.Identity())), orc.Term("norm1", orc.Grad(1, orc.Identity()))
        d = orc.Term("sum_squares", orc.Conv(psf, orc.Identity()), c=b)
        return orc.Solver([d, th, tw], method).solve(b, rhos=rho, lams={th: lam, tw: lam}, max_iter=Tn)

    got, want = run(ours, "cuda"), run(oracle, "cpu")
    assert rel(got[0], want[0]) < 1e-5
    for a, w_, name in zip(got[1:], want[1:], ("g_b", "g_rho", "g_lam")):
        assert rel(a, w_) < 2e-4, (name, rel(a, w_))


def test_backward_kernels_directly(dp):
    """dpx_prox_backward for every native body incl. the alpha/beta/offset wrapper chain, vs torch autograd of the oracle."""
    from dprox_b200 import ops, _cabi as cabi
    gen = torch.Generator().manual_seed(3)
    v0 = torch.randn(3, 2, 8, 12, generator=gen)
    off0 = 0.2 * torch.randn(3, 2, 8, 12, generator=gen)
    g0 = torch.randn(3, 2, 8, 12, generator=gen)
    lam0 = torch.tensor([0.1, 0.25, 0.4])
    for kind, name in ((cabi.PROX_NONNEG, "nonneg"), (cabi.PROX_L1, "norm1"), (cabi.PROX_L2SQ, "norm2")):
        for alpha, beta, use_off in ((1.0, 1.0, False), (0.5, 2.0, True)):
            v = v0.cuda().requires_grad_(True)
            lam = lam0.cuda().requires_grad_(True)
            off = off0.cuda() if use_off else None
            out = ops.prox(kind, v, lam, alpha, beta, 0.0, 0.0, off)
            out.backward(g0.cuda())
            vc, lc = v0.clone().requires_grad_(True), lam0.clone().requires_grad_(True)
            w = beta * (vc - (off0 if use_off else 0.0))
            body = {"nonneg": orc.prox_nonneg, "norm1": orc.prox_norm1, "norm2": orc.prox_norm2}[name]
            ref = body(w, beta * beta * lc.view(-1, 1, 1, 1) * alpha) / beta + (off0 if use_off else 0.0)
            ref.backward(g0)
            assert rel(out, ref) < 1e-6 and rel(v.grad, vc.grad) < 1e-6, (name, alpha)
            if name != "nonneg":
                assert rel(lam.grad, lc.grad) < 1e-5, (name, alpha, lam.grad, lc.grad)


def test_implicit_cg_gradients(dp):
    """a13: gradients through the CG x-update (joint demosaic + deconvolution, no closed form) by implicit differentiation:
    backward = a second run of the fused-kernel CG; against the reference's LinearSolve.backward (golden)."""
    g = load("unrolled_grads_cg")
    b = T(g["b"]).requires_grad_(True)
    x0 = T(g["x0"]).requires_grad_(True)
    rhos = T(g["rhos"]).requires_grad_(True)
    x = dp.Variable()
    f = dp.nonneg(x)
    cfg = dp.LinearSolveConfig(rtol=float(g["rtol"]), max_iters=int(g["cg_iters"]), solver_type="cg")
    solver = dp.compile(dp.sum_squares(dp.mosaic(dp.conv(x, g["psf"])) - b) + f, method="admm", device="cuda", linear_solve_config=cfg)
    assert solver.spec.xupdate == "cg"
    out = solver.solve(x0=x0, rhos=rhos, lams={f: torch.full((3,), 0.02)}, max_iter=int(g["T"]))
    (out * T(g["wgt"])).sum().backward()
    assert rel(out, g["out"]) < 2e-5
    for name, t in (("g_b", b), ("g_x0", x0), ("g_rhos", rhos)):
        assert t.grad is not None and rel(t.grad, g[name]) < 2e-4, (name, rel(t.grad, g[name]))
