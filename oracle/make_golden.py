"""TEST INFRASTRUCTURE — golden-vector generator.

Runs the UNMODIFIED reference (`/root/reference/dprox`, imported through `oracle/refshim.py`) on
small seeded inputs on CPU and stores inputs + outputs under `tests/golden/<case>.npz`.
It only runs in the authoring container (the GPU box has no /root/reference); the committed
`.npz` files are what `tests/` reads.

    python oracle/make_golden.py            # regenerate every case
    python oracle/make_golden.py admm_tv    # one case

Every case below names the reference entry points it exercises.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

dp = refshim.import_reference()
from dprox.contrib import blurring, point_spread_function  # noqa: E402
from dprox.linalg import LinearSolveConfig  # noqa: E402
from dprox.linalg.solve import cg as ref_cg, pcg as ref_pcg  # noqa: E402
from dprox.proxfn.pnp.denoisers.base import Denoiser  # noqa: E402
from dprox.proxfn.pnp.denoisers.models.network_ffdnet import FFDNet  # noqa: E402

import dprox_oracle as orc  # noqa: E402  (only for the seeded FFDNet weight generator)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(4)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _deconv_inputs(B, C, H, W, ksize=5, ksigma=2.0, seed=0, lo=-0.3):
    """img ~ U[lo, lo+1) so the nonneg constraint is active; b = blur(img) + 0.01 randn."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, C, H, W, generator=g) + lo
    psf = point_spread_function(ksize, ksigma)
    b = torch.cat([blurring(img[i:i + 1], psf) for i in range(B)], 0).float()
    b = b + 0.01 * torch.randn(B, C, H, W, generator=g)
    return img, psf, b


def _full_state(state):
    out = {}
    names = ["s0", "s1", "s2"]
    for name, s in zip(names, state):
        if isinstance(s, (list, tuple)):
            for i, e in enumerate(s):
                out[f"{name}_{i}"] = _np(e)
        else:
            out[name] = _np(s)
    return out


def _run(fns, method, x0, T, rhos=None, lams=None, **compile_kw):
    solver = dp.compile(fns, method=method, device="cpu", **compile_kw)
    with torch.no_grad():
        state = solver.solve(x0=x0, rhos=rhos, lams=lams, max_iter=T, return_full_states=True)
    return _full_state(state)


# ------------------------------------------------------------------------------------------------
# a4/a6/a5/a7/a8: the five iterations on the headline objective  sum_squares(conv(x)-b)+nonneg(x)
# ------------------------------------------------------------------------------------------------

@case
def admm_conv_nonneg():
    img, psf, b = _deconv_inputs(2, 3, 32, 48)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), "admm", b, 10)
    return dict(psf=psf, b=_np(b), T=10, **out)


@case
def admm_conv_nonneg_50it():
    """cfg1-class: 50 iterations (reference fp32 noise ~8e-6 at 50 it)."""
    img, psf, b = _deconv_inputs(1, 3, 64, 64, ksize=15, ksigma=5.0, seed=1)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), "admm", b, 50, rhos=0.5, lams=0.02)
    return dict(psf=psf, b=_np(b), T=50, rho=0.5, **out)


@case
def hqs_conv_nonneg():
    img, psf, b = _deconv_inputs(2, 3, 32, 48)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), "hqs", b, 10)
    return dict(psf=psf, b=_np(b), T=10, **out)


@case
def ladmm_conv_nonneg_b1():
    img, psf, b = _deconv_inputs(1, 3, 32, 48)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), "ladmm", b, 10)
    return dict(psf=psf, b=_np(b), T=10, **out)


@case
def vxu_conv_nonneg_b1():
    img, psf, b = _deconv_inputs(1, 3, 32, 48)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), "admm_vxu", b, 10)
    return dict(psf=psf, b=_np(b), T=10, **out)


@case
def pc_conv_nonneg():
    """PockChambolle (algo/pc.py:13-36) on the headline objective."""
    img, psf, b = _deconv_inputs(2, 3, 32, 48)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x) + dp.norm1(x), "pc", b, 8, rhos=0.7, lams=0.4)
    return dict(psf=psf, b=_np(b), T=8, rho=0.7, lam=0.4, **out)


@case
def pgd_conv_nonneg():
    img, psf, b = _deconv_inputs(2, 3, 32, 48)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), "pgd", b, 16, rhos=0.8)
    return dict(psf=psf, b=_np(b), T=16, rho=0.8, **out)


@case
def pgd_conv_norm1():
    img, psf, b = _deconv_inputs(2, 3, 32, 48, lo=-0.5)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf), b) + dp.norm1(x), "pgd", b, 12, rhos=0.9, lams=0.05)
    return dict(psf=psf, b=_np(b), T=12, rho=0.9, lam=0.05, **out)


# ------------------------------------------------------------------------------------------------
# a17/a19: prox wrapper chain, alpha scaling, two identity psi terms, per-sample schedules [B,T]
# ------------------------------------------------------------------------------------------------

@case
def admm_two_psi_per_sample():
    img, psf, b = _deconv_inputs(2, 3, 32, 48, lo=-0.5)
    g = torch.Generator().manual_seed(7)
    T = 10
    rhos = 0.5 + torch.rand(2, T, generator=g)
    lam1 = 0.01 + 0.05 * torch.rand(2, T, generator=g)
    lam2 = 0.02 * torch.ones(T)
    x = dp.Variable()
    f1, f2 = 0.5 * dp.norm1(x), dp.nonneg(x)
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + f1 + f2, "admm", b, T, rhos=rhos, lams={f1: lam1, f2: lam2})
    return dict(psf=psf, b=_np(b), T=T, rhos=_np(rhos), lam1=_np(lam1), lam2=_np(lam2), alpha1=0.5, **out)


@case
def hqs_two_psi_norm2():
    img, psf, b = _deconv_inputs(2, 3, 32, 48, lo=-0.5)
    x = dp.Variable()
    f1, f2 = dp.norm2(x), dp.norm1(x)
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + f1 + f2, "hqs", b, 8, rhos=0.7, lams={f1: 0.1, f2: 0.03})
    return dict(psf=psf, b=_np(b), T=8, rho=0.7, lam1=0.1, lam2=0.03, **out)


@case
def admm_psi_offset():
    """psi linop with a constant: norm1(x - c) -> prox_translated path (proxfn/base.py:24-27,43-45)."""
    img, psf, b = _deconv_inputs(2, 3, 32, 48, lo=-0.5)
    g = torch.Generator().manual_seed(11)
    c = 0.2 * torch.randn(2, 3, 32, 48, generator=g)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.norm1(x - c), "admm", b, 8, rhos=0.8, lams=0.05)
    return dict(psf=psf, b=_np(b), c=_np(c), T=8, rho=0.8, lam=0.05, **out)


# ------------------------------------------------------------------------------------------------
# a23/a25: asymmetric (complex-OTF) kernels and grad psi linops (anisotropic TV)
# ------------------------------------------------------------------------------------------------

@case
def admm_even_kernel():
    g = torch.Generator().manual_seed(3)
    img = torch.rand(2, 3, 32, 48, generator=g)
    k = torch.rand(4, 6, generator=g).numpy().astype("float32")
    k = (k / k.sum())[..., None]
    x = dp.Variable()
    op = dp.conv(x, k)
    K = dp.CompGraph(op)
    b = K.forward(img).float() + 0.01 * torch.randn(2, 3, 32, 48, generator=g)
    out = _run(dp.sum_squares(dp.conv(x, k) - b) + dp.nonneg(x), "admm", b, 8, rhos=0.3)
    return dict(kernel=k, b=_np(b), T=8, rho=0.3, **out)


@case
def admm_tv():
    img, psf, b = _deconv_inputs(2, 3, 32, 48, lo=0.0)
    x = dp.Variable()
    f1, f2 = dp.norm1(dp.grad(x, dim=0)), dp.norm1(dp.grad(x, dim=1))
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + f1 + f2, "admm", b, 10, rhos=2.0, lams=0.01)
    return dict(psf=psf, b=_np(b), T=10, rho=2.0, lam=0.01, **out)


@case
def hqs_tv_nonneg():
    img, psf, b = _deconv_inputs(1, 1, 32, 32, lo=0.0)
    x = dp.Variable()
    f1, f2, f3 = dp.norm1(dp.grad(x, dim=0)), dp.norm1(dp.grad(x, dim=1)), dp.nonneg(x)
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + f1 + f2 + f3, "hqs", b, 10, rhos=1.5, lams=0.02)
    return dict(psf=psf, b=_np(b), T=10, rho=1.5, lam=0.02, **out)


@case
def linops():
    """conv / grad forward+adjoint and OTFs on a random tensor (a23, a25)."""
    g = torch.Generator().manual_seed(5)
    t = torch.randn(2, 3, 16, 24, generator=g)
    psf = point_spread_function(5, 2)
    k2 = torch.rand(4, 3, 1, generator=g).numpy().astype("float32")
    x = dp.Variable()
    out = dict(t=_np(t), psf=psf, k2=k2)
    for name, op in [("conv", dp.conv(x, psf)), ("conv2", dp.conv(x, k2)), ("grad0", dp.grad(x, dim=0)),
                     ("grad1", dp.grad(x, dim=1)), ("grad2", dp.grad(x, dim=2))]:
        out[name + "_fwd"] = _np(op.forward(t))
        out[name + "_adj"] = _np(op.adjoint(t))
        fb = op._FB(tuple(t.shape))
        out[name + "_otf"] = _np(fb)
        out[name + "_diag"] = _np(op.get_diag(t, freq=True))
    m = dp.mosaic(x)
    out["mosaic_fwd"] = _np(m.forward(t))
    out["mosaic_diag"] = _np(m.get_diag(t, freq=False))
    return out


# ------------------------------------------------------------------------------------------------
# a11 spatial-diag branch, a26: mosaic / mul_elementwise
# ------------------------------------------------------------------------------------------------

@case
def admm_mosaic_spatial():
    g = torch.Generator().manual_seed(9)
    img = torch.rand(2, 3, 32, 48, generator=g)
    x = dp.Variable()
    b = dp.mosaic(x).forward(img) + 0.01 * torch.randn(2, 3, 32, 48, generator=g)
    s = dp.compile(dp.sum_squares(dp.mosaic(x) - b) + dp.nonneg(x), method="admm", device="cpu")
    flags = (bool(s.least_square.diagonalizable), bool(s.least_square.freq_diagonalizable))
    out = _run(dp.sum_squares(dp.mosaic(x) - b) + dp.nonneg(x), "admm", b, 8, rhos=0.6)
    return dict(b=_np(b), T=8, rho=0.6, flags=np.array(flags), **out)


@case
def hqs_mul_elementwise():
    g = torch.Generator().manual_seed(10)
    img = torch.rand(2, 3, 32, 48, generator=g)
    w = (torch.rand(1, 3, 32, 48, generator=g) > 0.4).float() * (0.5 + torch.rand(1, 3, 32, 48, generator=g))
    x = dp.Variable()
    b = w * img
    out = _run(dp.sum_squares(dp.mul_elementwise(x, w) - b) + dp.norm1(x), "hqs", b, 8, rhos=0.6, lams=0.03)
    return dict(b=_np(b), w=_np(w), T=8, rho=0.6, lam=0.03, **out)


# ------------------------------------------------------------------------------------------------
# a12/a13/a14: CG fallback (joint demosaic + deconv) and the raw solvers
# ------------------------------------------------------------------------------------------------

@case
def admm_cg_mosaic_conv():
    g = torch.Generator().manual_seed(12)
    img = torch.rand(1, 3, 24, 32, generator=g)
    psf = point_spread_function(5, 1.5)
    x = dp.Variable()
    op = dp.mosaic(dp.conv(x, psf))
    b = dp.CompGraph(op).forward(img).float()
    cfg = LinearSolveConfig(rtol=1e-6, max_iters=30, solver_type="cg")
    out = _run(dp.sum_squares(dp.mosaic(dp.conv(x, psf)) - b) + dp.nonneg(x), "admm", b, 4, rhos=0.5,
               linear_solve_config=cfg)
    return dict(psf=psf, b=_np(b), T=4, rho=0.5, cg_iters=30, **out)


@case
def admm_pcg_mosaic_conv():
    g = torch.Generator().manual_seed(13)
    img = torch.rand(2, 3, 24, 32, generator=g)
    psf = point_spread_function(5, 1.5)
    x = dp.Variable()
    b = dp.CompGraph(dp.mosaic(dp.conv(x, psf))).forward(img).float()
    cfg = LinearSolveConfig(rtol=1e-6, max_iters=25, solver_type="pcg")
    out = _run(dp.sum_squares(dp.mosaic(dp.conv(x, psf)) - b) + dp.nonneg(x), "admm", b, 3, rhos=0.5,
               linear_solve_config=cfg)
    return dict(psf=psf, b=_np(b), T=3, rho=0.5, cg_iters=25, **out)


@case
def ladmm_tv_3it():
    """LinearizedADMM with grad psi linops: the reference's self-inconsistent update, 3 iterations."""
    img, psf, b = _deconv_inputs(1, 3, 32, 48, lo=0.0)
    x = dp.Variable()
    f1, f2 = dp.norm1(dp.grad(x, dim=0)), dp.norm1(dp.grad(x, dim=1))
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + f1 + f2, "ladmm", b, 3, rhos=2.0, lams=0.01)
    return dict(psf=psf, b=_np(b), T=3, rho=2.0, lam=0.01, **out)


def _csmri_ops(mask):
    def fwd(x, step=0):
        return mask * torch.fft.fft2(x, norm="ortho")
    def adj(y, step=0):
        return torch.real(torch.fft.ifft2(mask * y, norm="ortho"))
    return fwd, adj


@case
def admm_csmri_blackbox():
    """cfg3: subsampled-FFT BlackBox data term (complex k-space) + anisotropic TV, ADMM, (P)CG inner solve."""
    g = torch.Generator().manual_seed(31)
    H = W = 32
    img = torch.zeros(1, 1, H, W)
    img[..., 8:24, 10:20] = 1.0
    img[..., 12:18, 4:28] += 0.5
    mask = (torch.rand(1, 1, H, W, generator=g) < 0.3).float()
    mask[..., :4, :4] = 1; mask[..., -4:, :4] = 1; mask[..., :4, -4:] = 1; mask[..., -4:, -4:] = 1
    fwd, adj = _csmri_ops(mask)
    y0 = fwd(img)
    x0 = adj(y0)
    out = {}
    for solver in ("cg", "pcg"):
        x = dp.Variable()
        A = dp.LinOpFactory(fwd, adj)
        fns = dp.sum_squares(A(x), y0) + dp.norm1(dp.grad(x, dim=0)) + dp.norm1(dp.grad(x, dim=1))
        cfg = LinearSolveConfig(rtol=1e-6, max_iters=20, solver_type=solver)
        res = _run(fns, "admm", x0, 5, rhos=1.0, lams=0.05, linear_solve_config=cfg)
        for k, v in res.items():
            out[f"{solver}_{k}"] = v
    return dict(mask=_np(mask), y0_re=_np(y0.real), y0_im=_np(y0.imag), x0=_np(x0), img=_np(img), T=5, rho=1.0, lam=0.05,
                cg_iters=20, **out)


@case
def linear_solvers():
    """tests/linalg/test_linear_solver.py:57-111 style: SPD systems, fp64, plus a batched conv system."""
    rs = np.random.RandomState(0)
    M = rs.rand(5, 5)
    A = M @ M.T + 5 * np.eye(5)
    xs = rs.rand(5)
    bvec = A @ xs
    At, bt = torch.from_numpy(A), torch.from_numpy(bvec)
    out = dict(A=A, b=bvec, x_true=xs)
    out["cg"] = _np(ref_cg(lambda v: At @ v, bt, rtol=1e-8, max_iters=100))
    out["pcg"] = _np(ref_pcg(lambda v: At @ v, bt, rtol=1e-8, max_iters=100))
    # batched [B,C,H,W] system (I + 0.5 * K^T K) x = rhs with a conv K
    g = torch.Generator().manual_seed(2)
    psf = point_spread_function(5, 2)
    x = dp.Variable()
    cv = dp.conv(x, psf)
    rhs = torch.rand(2, 3, 16, 24, generator=g)
    Aop = lambda v: v + 0.5 * cv.adjoint(cv.forward(v))
    out["psf"], out["rhs"] = psf, _np(rhs)
    out["cg_conv"] = _np(ref_cg(Aop, rhs, rtol=1e-6, max_iters=12))
    out["pcg_conv"] = _np(ref_pcg(Aop, rhs, rtol=1e-6, max_iters=12))
    return out


# ------------------------------------------------------------------------------------------------
# known answers of the reference's own tests (tests/problem/test_ml_problems.py:5-44)
# ------------------------------------------------------------------------------------------------

@case
def ml_problems():
    out = {}
    rhs = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    x = dp.Variable((3, 3))
    dp.Problem(dp.sum_squares(2 * x - rhs)).solve("admm", device="cpu", x0=np.zeros((3, 3)))
    out["lsq"] = _np(x.value)
    x = dp.Variable((3, 3))
    dp.Problem(dp.sum_squares(2 * x, rhs)).solve("admm", device="cpu", x0=np.zeros((3, 3)))
    out["lsq1"] = _np(x.value)
    x = dp.Variable((3, 3, 1))
    rhs2 = np.array([[[1, 2, 3], [4, 5, 6], [7, 8, 9]]])
    kernel = np.array([[1, 1], [1, 1]]) / 4
    dp.Problem(dp.sum_squares(dp.conv(x, kernel) - rhs2)).solve("admm", device="cpu", x0=np.zeros((3, 3, 1)))
    out["lsq2"] = _np(x.value)
    out["lsq2_resid"] = _np(dp.eval(dp.conv(x, kernel) - rhs2, x.value, zero_out_constant=False))
    x = dp.Variable((3))
    rhs3 = np.array([1, 2, 3])
    dp.Problem(dp.sum_squares(2 * x - rhs3)).solve("admm", device="cpu", x0=np.zeros(3))
    out["lsq3"] = _np(x.value)
    out["rhs"], out["rhs2"], out["rhs3"], out["kernel"] = rhs, rhs2, rhs3, kernel
    return out


# ------------------------------------------------------------------------------------------------
# a24: conv_doe (learnable PSF) incl. the channel-roll and the even-pad off-by-one quirks
# ------------------------------------------------------------------------------------------------

@case
def hqs_conv_doe():
    out = {}
    for tag, h in (("full", 32), ("padded", 24)):
        g = torch.Generator().manual_seed(21)
        img = torch.rand(2, 3, 32, 32, generator=g)
        psf = torch.rand(1, 3, h, h, generator=g)
        psf = psf / psf.sum(dim=(-2, -1), keepdim=True)
        x = dp.Variable()
        op = dp.conv_doe(x, psf, circular=True)
        with torch.no_grad():
            b = op.forward(img) + 0.01 * torch.randn(2, 3, 32, 32, generator=g)
            res = _run(dp.sum_squares(dp.conv_doe(x, psf, circular=True) - b) + dp.nonneg(x), "hqs", b, 6, rhos=0.4)
            out[f"{tag}_otf"] = _np(sys.modules['dprox.linop.conv'].psf2otf2(psf, (2, 3, 32, 32)))
        out[f"{tag}_psf"], out[f"{tag}_b"] = _np(psf), _np(b)
        for k, v in res.items():
            out[f"{tag}_{k}"] = v
    out["T"], out["rho"] = 6, 0.4
    return out


# ------------------------------------------------------------------------------------------------
# a20: deep_prior -> FFDNet-color with seeded random weights (pretrained weights need network)
# ------------------------------------------------------------------------------------------------

class _RandFFDNetColor(Denoiser):
    def __init__(self, seed):
        super().__init__()
        self.model = FFDNet(in_nc=3, out_nc=3, nc=96, nb=12, act_mode="R")
        ws = orc.ffdnet_random_weights(seed)
        sd = {}
        for i, (w, b) in enumerate(ws):
            sd[f"model.{2 * i}.weight"], sd[f"model.{2 * i}.bias"] = w, b
        self.model.load_state_dict(sd, strict=True)

    def _denoise(self, x, sigma):
        return self.model(x, sigma)


@case
def admm_deep_prior_ffdnet():
    """deep_prior + nonneg need a full lams dict (admm.py:56 indexes lam[fn] for every psi fn)."""
    img, psf, b = _deconv_inputs(2, 3, 24, 30, lo=0.0)
    den = _RandFFDNetColor(seed=4)
    x = dp.Variable()
    rhos, sigmas = dp.log_descent(35, 30, 4)
    prior, nn_ = dp.deep_prior(x, denoiser=den), dp.nonneg(x)
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + prior + nn_, "admm", b, 4, rhos=rhos,
               lams={prior: sigmas, nn_: 0.02})
    return dict(psf=psf, b=_np(b), T=4, rhos=_np(rhos), sigmas=_np(sigmas), seed=4, **out)


@case
def admm_deep_prior_wellcond():
    """Same objective with rho = 0.3: well conditioned, so fp32 implementations agree to ~1e-6 (the
    log_descent schedule above has rho ~ 1e-5, where the reference itself is 1e-3 away from fp64)."""
    img, psf, b = _deconv_inputs(2, 3, 24, 30, lo=0.0)
    den = _RandFFDNetColor(seed=4)
    x = dp.Variable()
    _, sigmas = dp.log_descent(35, 30, 4)
    prior, nn_ = dp.deep_prior(x, denoiser=den), dp.nonneg(x)
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + prior + nn_, "admm", b, 4, rhos=0.3,
               lams={prior: sigmas, nn_: 0.02})
    return dict(psf=psf, b=_np(b), T=4, rho=0.3, sigmas=_np(sigmas), seed=4, **out)


@case
def ffdnet_forward():
    den = _RandFFDNetColor(seed=4)
    g = torch.Generator().manual_seed(6)
    xin = torch.rand(2, 3, 15, 18, generator=g)              # odd H: exercises replicate pad + crop
    sig = torch.tensor([0.05, 0.12])
    with torch.no_grad():
        y = den.denoise(xin, sig)
    ws = orc.ffdnet_random_weights(4)
    return dict(x=_np(xin), sigma=_np(sig), y=_np(y), w0_sum=float(ws[0][0].double().sum()),
                wlast_sum=float(ws[-1][0].double().sum()))


@case
def schedules():
    r1, s1 = dp.log_descent(35, 30, 24)
    r2, s2 = dp.log_descent(49, 7.65, 10, sigma=7.65 / 255, sqrt=True)
    return dict(r1=_np(r1), s1=_np(s1), r2=_np(r2), s2=_np(s2))

# ------------------------------------------------------------------------------------------------
# a28 / §3.6: unrolled solver, gradients by the reference's plain autograd through the loop
# ------------------------------------------------------------------------------------------------

@case
def unrolled_grads_native():
    """ADMM with native proxes, 4 unrolled iterations: d loss / d (rhos [B,T], lam of the l1 term, measurements b, x0)."""
    img, psf, b = _deconv_inputs(2, 3, 16, 24)
    g = torch.Generator().manual_seed(77)
    wgt = torch.rand(2, 3, 16, 24, generator=g)
    b = b.clone().requires_grad_(True)
    x0 = torch.rand(2, 3, 16, 24, generator=g).requires_grad_(True)
    rhos = (0.5 + torch.rand(2, 4, generator=g)).requires_grad_(True)
    lam1 = (0.02 + 0.05 * torch.rand(4, generator=g)).requires_grad_(True)
    x = dp.Variable()
    f1, f2 = 0.5 * dp.norm1(x), dp.nonneg(x)
    solver = dp.compile(dp.sum_squares(dp.conv(x, psf) - b) + f1 + f2, method="admm", device="cpu")
    out = solver.solve(x0=x0, rhos=rhos, lams={f1: lam1, f2: torch.full((4,), 0.02)}, max_iter=4)
    loss = (out * wgt).sum()
    loss.backward()
    return dict(psf=psf, b=_np(b), x0=_np(x0), wgt=_np(wgt), rhos=_np(rhos), lam1=_np(lam1), out=_np(out), loss=float(loss),
                g_b=_np(b.grad), g_x0=_np(x0.grad), g_rhos=_np(rhos.grad), g_lam1=_np(lam1.grad), T=4)


@case
def unrolled_grads_doe():
    """BASELINE config 5 in miniature: conv_doe (Placeholder PSF) + deep_prior(FFDNet, sqrt=True), `specialize('unroll')`,
    mse loss; gradients w.r.t. rhos, sigmas and -- through the data formation inp = conv_doe(gt; psf) -- the PSF
    (examples/papers/deltaprox_siggraph_2023/computional_optics/e2e_optics_dprox.py:46-58)."""
    g = torch.Generator().manual_seed(31)
    gt = torch.rand(2, 3, 16, 16, generator=g)
    psf = torch.rand(1, 3, 16, 16, generator=g)
    psf = (psf / psf.sum(dim=(-2, -1), keepdim=True)).requires_grad_(True)
    noise = 0.03 * torch.randn(2, 3, 16, 16, generator=g)
    den = _RandFFDNetColor(seed=9)
    x, y, PSF = dp.Variable(), dp.Placeholder(), dp.Placeholder()
    data_term = dp.sum_squares(dp.conv_doe(x, PSF, circular=True), y)
    reg_term = dp.deep_prior(x, denoiser=den, sqrt=True)
    solver = dp.compile(data_term + reg_term, method="admm", device="cpu")
    solver = dp.specialize(solver, method="unroll", max_iter=3, device="cpu")
    rhos, sigmas = dp.log_descent(49, 7.65, 3, sigma=7.65 / 255)
    rhos, sigmas = rhos.clone().requires_grad_(True), sigmas.clone().requires_grad_(True)
    psf_used = psf * 1.0
    inp = dp.conv_doe(dp.Variable(), psf_used.detach(), circular=True).forward(gt)      # data formation (differentiable below)
    otf = sys.modules['dprox.linop.conv'].psf2otf2(psf_used, gt.shape)
    inp = torch.real(torch.fft.ifftn(otf * torch.fft.fftn(gt, dim=[-2, -1]), dim=[-2, -1])).float() + noise
    y.value = inp
    PSF.value = psf_used
    out = solver.solve(x0=inp, rhos=rhos, lams={reg_term: sigmas})
    loss = torch.nn.functional.mse_loss(gt, out)
    loss.backward()
    return dict(gt=_np(gt), psf=_np(psf), noise=_np(noise), inp=_np(inp), rhos=_np(rhos), sigmas=_np(sigmas), out=_np(out),
                loss=float(loss), g_rhos=_np(rhos.grad), g_sigmas=_np(sigmas.grad), g_psf=_np(psf.grad), seed=9, T=3)

# ------------------------------------------------------------------------------------------------
# §8f-2: csmri closed-form data term (ext_sum_squares hook) on complex state, driven by CustomADMM
# ------------------------------------------------------------------------------------------------

class _RandFFDNetGray(Denoiser):
    def __init__(self, seed):
        super().__init__()
        self.model = FFDNet(in_nc=1, out_nc=1, nc=96, nb=12, act_mode="R")
        ws = orc.ffdnet_random_weights(seed, in_nc=1)
        sd = {}
        for i, (w, b) in enumerate(ws):
            sd[f"model.{2 * i}.weight"], sd[f"model.{2 * i}.bias"] = w, b
        self.model.load_state_dict(sd, strict=True)

    def _denoise(self, x, sigma):
        return self.model(x, sigma)


@case
def csmri_custom_admm():
    """tests/paper/test_csmri.py:29-65: csmri(x, mask, y) + deep_prior, CustomADMM (prox first), complex iterates."""
    from dprox.contrib.csmri import CustomADMM
    from dprox.proxfn.fast.csmri import csmri
    from dprox.utils import fft2, ifft2
    out = {}
    for tag, (H, W) in (("even", (32, 48)), ("odd", (31, 33))):
        g = torch.Generator().manual_seed(41)
        img = torch.zeros(2, 1, H, W)
        img[:, :, H // 4: 3 * H // 4, W // 3: 2 * W // 3] = 1.0
        img = img + 0.1 * torch.rand(2, 1, H, W, generator=g)
        mask = (torch.rand(2, 1, H, W, generator=g) < 0.35).float()
        mask[:, :, H // 2 - 3: H // 2 + 3, W // 2 - 3: W // 2 + 3] = 1.0
        noise = 0.01 * (torch.randn(2, 1, H, W, generator=g) + 1j * torch.randn(2, 1, H, W, generator=g))
        y0 = mask * (fft2(img.to(torch.complex64)) + noise)
        x0 = ifft2(y0)
        den = _RandFFDNetGray(seed=12)
        x, y, m = dp.Variable(), dp.Placeholder(), dp.Placeholder()
        data_term, reg_term = csmri(x, m, y), dp.deep_prior(x, denoiser=den)
        solver = CustomADMM([reg_term], [data_term])
        y.value, m.value = y0, mask
        rhos = torch.tensor([0.5, 0.8, 1.2, 2.0])
        sigmas = torch.tensor([0.08, 0.06, 0.04, 0.03])
        with torch.no_grad():
            st = solver.solve(x0=x0, rhos=rhos, lams={reg_term: sigmas}, max_iter=4, return_full_states=True)
            one = data_term._prox(x0 * (1 + 0.5j), torch.tensor([0.7, 1.3]), 1)
        out.update({f"{tag}_mask": _np(mask), f"{tag}_y0": _np(y0), f"{tag}_x0": _np(x0), f"{tag}_x": _np(st[0]),
                    f"{tag}_z": _np(st[1][0]), f"{tag}_u": _np(st[2][0]), f"{tag}_prox1": _np(one)})
    out.update(rhos=_np(rhos), sigmas=_np(sigmas), seed=12, T=4)
    return out

@case
def drunet_forward():
    """§8f-4: DRUNetDenoiser (UNetRes + the quadrant tiling of wrapper.py:111-146) with seeded random weights: a small
    image (replicate-pad to a multiple of 16) and one just above 256 x 256 (four overlapping quadrants)."""
    import tempfile
    from dprox.proxfn.pnp.denoisers.models.network_unet import UNetRes
    from dprox.proxfn.pnp.denoisers.wrapper import DRUNetDenoiser
    net = UNetRes(in_nc=2, out_nc=1, nc=[64, 128, 256, 512], nb=4, act_mode="R", downsample_mode="strideconv",
                  upsample_mode="convtranspose")
    g = torch.Generator().manual_seed(3)
    sd = net.state_dict()
    with torch.no_grad():
        for k, w in sd.items():
            w.copy_((torch.rand(w.shape, generator=g) * 2 - 1) / (w[0].numel() ** 0.5))
    with tempfile.NamedTemporaryFile(suffix=".pth") as f:
        torch.save(sd, f.name)
        den = DRUNetDenoiser(1, f.name)
    gi = torch.Generator().manual_seed(8)
    out = dict(seed=3, keys=np.array(list(sd.keys())))
    for tag, (h, w) in (("small", (40, 52)), ("tiled", (272, 264))):
        x = torch.rand(1, 1, h, w, generator=gi)
        with torch.no_grad():
            y = den.denoise(x, torch.tensor([0.07]))
        out[f"{tag}_x"], out[f"{tag}_y"] = _np(x), _np(y)
    return out

@case
def unrolled_grads_cg():
    """a13: gradients through the CG x-update by LinearSolve's implicit differentiation (linalg/custom.py:39-62):
    grad_b = solve(A^T, grad_x).  NB: the reference's KtK module multiplies by the closure `rho`, not by its own
    `self.rho` parameter (sum_square.py:160-173), so the rho-dependence of the MATRIX is not differentiated -- only the
    right-hand side's (rho * K_i^T b_i).  The golden vectors record what the reference actually returns."""
    g = torch.Generator().manual_seed(19)
    img = torch.rand(1, 3, 16, 24, generator=g)
    psf = point_spread_function(5, 1.5)
    x = dp.Variable()
    b = dp.CompGraph(dp.mosaic(dp.conv(x, psf))).forward(img).float().detach()
    wgt = torch.rand(1, 3, 16, 24, generator=g)
    b = b.clone().requires_grad_(True)
    x0 = torch.rand(1, 3, 16, 24, generator=g).requires_grad_(True)
    rhos = torch.tensor([0.5, 0.8, 1.1]).requires_grad_(True)
    cfg = LinearSolveConfig(rtol=1e-7, max_iters=60, solver_type="cg")
    f = dp.nonneg(x)
    solver = dp.compile(dp.sum_squares(dp.mosaic(dp.conv(x, psf)) - b) + f, method="admm", device="cpu", linear_solve_config=cfg)
    out = solver.solve(x0=x0, rhos=rhos, lams={f: torch.full((3,), 0.02)}, max_iter=3)
    (out * wgt).sum().backward()
    return dict(psf=psf, b=_np(b), x0=_np(x0), wgt=_np(wgt), rhos=_np(rhos), out=_np(out), g_b=_np(b.grad), g_x0=_np(x0.grad),
                g_rhos=_np(rhos.grad), T=3, cg_iters=60, rtol=1e-7)

# ------------------------------------------------------------------------------------------------
# §8f-3: DOE optics forward model (contrib/optic): get_psf pipeline + img_psf_conv, values and autograd gradients
# ------------------------------------------------------------------------------------------------

@case
def doe_forward_model():
    from dprox.contrib.optic.doe_model import RGBCollimator
    from dprox.contrib.optic.common import img_psf_conv
    N, n = 64, 32
    kw = dict(sensor_distance=15e-3, refractive_idcs=torch.tensor([1.4648, 1.4599, 1.4568]),
              wave_lengths=torch.tensor([460, 550, 640]) * 1e-9, patch_size=n, sample_interval=2e-6 * (1496 / N) * 0.25,
              wave_resolution=(N, N))
    m = RGBCollimator(**kw)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():                                   # a non-trivial operating point: perturb the Fresnel-lens init
        m.height_map.height_map_sqrt.mul_(1.0 + 0.05 * torch.rand(1, 1, N, N, generator=g))
    h0 = m.height_map.height_map_sqrt.detach().clone()
    wgt = torch.rand(1, 3, n, n, generator=g)
    psf = m.get_psf()
    (psf * wgt).sum().backward()
    out = dict(N=N, n=n, sample_interval=kw["sample_interval"], h0=_np(h0), wgt=_np(wgt), psf=_np(psf),
               g_h=_np(m.height_map.height_map_sqrt.grad), aperture=_np(m.aperture), H=_np(m.propagator.H))
    # data formation with a smaller PSF than the image (even pad: the off-by-one quirk) and gradients to both inputs
    img = torch.rand(2, 3, 40, 40, generator=g).requires_grad_(True)
    p2 = psf.detach().clone().requires_grad_(True)
    w2 = torch.rand(2, 3, 40, 40, generator=g)
    y = img_psf_conv(img, p2, circular=True)
    (y * w2).sum().backward()
    out.update(img=_np(img), w2=_np(w2), y=_np(y), g_img=_np(img.grad), g_psf=_np(p2.grad))
    return out


# ------------------------------------------------------------------------------------------------
# round 2: reference vectors at sizes the fused sm_100a FFT engine takes (>= 64 points per side), every first-pass radix
# ------------------------------------------------------------------------------------------------

@case
def cfg1_admm_256_50it():
    """BASELINE configs[0]: single [1,3,256,256] deconv, sum_squares(conv)+nonneg, ADMM 50 iterations, PSF 15/5."""
    img, psf, b = _deconv_inputs(1, 3, 256, 256, ksize=15, ksigma=5.0, seed=2)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), "admm", b, 50, rhos=1.0, lams=0.02)
    return dict(psf=psf, b=_np(b), T=50, rho=1.0, s0=out["s0"])            # x only: keeps the fixture at ~1.5 MB


@case
def admm_fused_128x192():
    """plane-pair engine, radix-12 row pass (192 = 3 * 64), radix-8/16 columns."""
    img, psf, b = _deconv_inputs(2, 1, 128, 192, ksize=9, ksigma=2.0, seed=3)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), "admm", b, 10, rhos=0.7, lams=0.02)
    return dict(psf=psf, b=_np(b), T=10, rho=0.7, **out)


@case
def hqs_fused_64x320():
    """plane-pair engine, radix-10 row pass (320 = 5 * 64), 64-point columns; HQS."""
    img, psf, b = _deconv_inputs(2, 1, 64, 320, ksize=9, ksigma=2.0, seed=4)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.nonneg(x), "hqs", b, 10, rhos=0.7, lams=0.02)
    return dict(psf=psf, b=_np(b), T=10, rho=0.7, **out)


@case
def ladmm_csmri_blackbox_3it():
    """BASELINE configs[2] as stated: subsampled-FFT BlackBox + TV under LADMM with a PCG inner solve, 3 iterations
    (the reference's LADMM diverges on this operator from iteration 4 on)."""
    g = torch.Generator().manual_seed(32)
    H = W = 32
    img = torch.zeros(1, 1, H, W)
    img[..., 6:22, 9:21] = 1.0
    img[..., 12:18, 4:28] += 0.5
    mask = (torch.rand(1, 1, H, W, generator=g) < 0.3).float()
    mask[..., :4, :4] = 1; mask[..., -4:, :4] = 1; mask[..., :4, -4:] = 1; mask[..., -4:, -4:] = 1
    fwd, adj = _csmri_ops(mask)
    y0 = fwd(img)
    x0 = adj(y0)
    x = dp.Variable()
    A = dp.LinOpFactory(fwd, adj)
    fns = dp.sum_squares(A(x), y0) + dp.norm1(dp.grad(x, dim=0)) + dp.norm1(dp.grad(x, dim=1))
    cfg = LinearSolveConfig(rtol=1e-6, max_iters=30, solver_type="pcg")
    res = _run(fns, "ladmm", x0, 3, rhos=1.0, lams=0.05, linear_solve_config=cfg)
    return dict(mask=_np(mask), y0_re=_np(y0.real), y0_im=_np(y0.imag), x0=_np(x0), T=3, rho=1.0, lam=0.05, cg_iters=30, **res)


# ------------------------------------------------------------------------------------------------
# round 2: rows that were partial (a7 external prox, a20 x8, a24 circular=False, a25 dim=2, a26 spatial diag in the
# generic / differentiable engine, a28 share=False / learned_params) and the advisor's cases
# ------------------------------------------------------------------------------------------------

@case
def conv_doe_linear():
    """conv_doe(circular=False) (linop/conv.py:100-153): zero-pad to 2H, FFT conv, crop; forward, adjoint and an HQS solve
    (whose closed-form diagonal is -- as in the reference -- the CIRCULAR |OTF|^2 at the image size, conv.py:143-152)."""
    out = {}
    for tag, (H, h) in (("even", (32, 32)), ("small", (32, 20)), ("odd", (27, 27))):
        g = torch.Generator().manual_seed(23)
        img = torch.rand(2, 3, H, H, generator=g)
        psf = torch.rand(1, 3, h, h, generator=g)
        psf = psf / psf.sum(dim=(-2, -1), keepdim=True)
        t = torch.randn(2, 3, H, H, generator=g)
        x = dp.Variable()
        op = dp.conv_doe(x, psf, circular=False)
        with torch.no_grad():
            out[f"{tag}_fwd"], out[f"{tag}_adj"] = _np(op.forward(t)), _np(op.adjoint(t))
            b = op.forward(img) + 0.01 * torch.randn(2, 3, H, H, generator=g)
            res = _run(dp.sum_squares(dp.conv_doe(x, psf, circular=False) - b) + dp.nonneg(x), "hqs", b, 5, rhos=0.4)
        out[f"{tag}_psf"], out[f"{tag}_t"], out[f"{tag}_b"] = _np(psf), _np(t), _np(b)
        for k, v in res.items():
            out[f"{tag}_{k}"] = v
    out["T"], out["rho"] = 5, 0.4
    return out


@case
def img_psf_conv_linear():
    """contrib/optic/common.py:85-118 with circular=False: values and gradients w.r.t. image and PSF."""
    from dprox.contrib.optic.common import img_psf_conv
    out = {}
    for tag, (H, h) in (("same", (24, 24)), ("small", (24, 16))):
        g = torch.Generator().manual_seed(29)
        img = torch.rand(2, 3, H, H, generator=g).requires_grad_(True)
        psf = torch.rand(1, 3, h, h, generator=g)
        psf = (psf / psf.sum(dim=(-2, -1), keepdim=True)).requires_grad_(True)
        w = torch.rand(2, 3, H, H, generator=g)
        y = img_psf_conv(img, psf, circular=False)
        (y * w).sum().backward()
        out.update({f"{tag}_img": _np(img), f"{tag}_psf": _np(psf), f"{tag}_w": _np(w), f"{tag}_y": _np(y),
                    f"{tag}_g_img": _np(img.grad), f"{tag}_g_psf": _np(psf.grad)})
    return out


@case
def admm_grad_dim2():
    """grad(x, dim=2) (channel axis, linop/grad.py:14-23) as a psi linop next to the H and W gradients (all three, so that
    the closed-form denominator has no near-null frequencies: the reference solves this in float64 -- int64 kernel -- and an
    ill-conditioned case would measure precision, not the operator)."""
    img, psf, b = _deconv_inputs(2, 3, 32, 48, lo=0.0)
    x = dp.Variable()
    f1, f2, f3 = dp.norm1(dp.grad(x, dim=2)), dp.norm1(dp.grad(x, dim=1)), dp.norm1(dp.grad(x, dim=0))
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + f1 + f2 + f3, "admm", b, 6, rhos=1.5, lams=0.02)
    return dict(psf=psf, b=_np(b), T=6, rho=1.5, lam=0.02, **out)


@case
def vxu_deep_prior():
    """ADMM_vxu (prox first) with an external deep prior + nonneg."""
    img, psf, b = _deconv_inputs(2, 3, 24, 30, lo=0.0)
    den = _RandFFDNetColor(seed=4)
    x = dp.Variable()
    _, sigmas = dp.log_descent(35, 30, 4)
    prior, nn_ = dp.deep_prior(x, denoiser=den), dp.nonneg(x)
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + prior + nn_, "admm_vxu", b, 4, rhos=0.3,
               lams={prior: sigmas, nn_: 0.02})
    return dict(psf=psf, b=_np(b), T=4, rho=0.3, sigmas=_np(sigmas), seed=4, **out)


@case
def deep_prior_x8():
    """deep_prior(x8=True): the Augment wrapper cycles the 8 flips/rotations, one per call (composite.py:6-47)."""
    img, psf, b = _deconv_inputs(1, 3, 24, 30, lo=0.0)
    den = _RandFFDNetColor(seed=4)
    x = dp.Variable()
    T = 9
    sig = torch.linspace(0.12, 0.04, T)
    prior, nn_ = dp.deep_prior(x, denoiser=den, x8=True), dp.nonneg(x)
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + prior + nn_, "admm", b, T, rhos=0.3, lams={prior: sig, nn_: 0.02})
    return dict(psf=psf, b=_np(b), T=T, rho=0.3, sigmas=_np(sig), seed=4, **out)


@case
def pc_identity():
    """PockChambolle on an identity-only objective (scalar closed form in the node-by-node engine)."""
    g = torch.Generator().manual_seed(51)
    b = torch.rand(2, 3, 16, 24, generator=g) - 0.4
    x = dp.Variable()
    out = _run(dp.sum_squares(x - b) + dp.norm1(x), "pc", b, 8, rhos=0.7, lams=0.4)
    return dict(b=_np(b), T=8, rho=0.7, lam=0.4, **out)


@case
def admm_mask_psi():
    """a mask as a PSI linop: norm1(mosaic(x)); spatial-diagonal closed form with dq + rho * dpsi."""
    g = torch.Generator().manual_seed(52)
    b = torch.rand(2, 3, 16, 24, generator=g) - 0.4
    x = dp.Variable()
    out = _run(dp.sum_squares(x - b) + dp.norm1(dp.mosaic(x)), "admm", b, 6, rhos=0.8, lams=0.1)
    return dict(b=_np(b), T=6, rho=0.8, lam=0.1, **out)


@case
def ladmm_scaled_identity():
    """LADMM with a SCALED identity psi linop: b_i = x - K^T(Kx - v + u) differs from v - u unless the scale is 1."""
    img, psf, b = _deconv_inputs(1, 3, 32, 48, lo=-0.5)      # B = 1: a single psi fn with B > 1 hits the batch-index slip (App. A-17)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf) - b) + dp.norm1(2 * x), "ladmm", b, 5, rhos=0.8, lams=0.05)
    return dict(psf=psf, b=_np(b), T=5, rho=0.8, lam=0.05, **out)


@case
def pgd_psi_linop():
    """PGD ignores the psi linop and applies the prox to x directly (pgd.py:39-43)."""
    img, psf, b = _deconv_inputs(2, 3, 32, 48, lo=-0.5)
    x = dp.Variable()
    out = _run(dp.sum_squares(dp.conv(x, psf), b) + dp.norm1(dp.grad(x, dim=1)), "pgd", b, 6, rhos=0.9, lams=0.05)
    out2 = _run(dp.sum_squares(dp.conv(x, psf), b) + dp.norm1(2 * x), "pgd", b, 6, rhos=0.9, lams=0.05)
    return dict(psf=psf, b=_np(b), T=6, rho=0.9, lam=0.05, scaled_s0=out2["s0"], **out)


@case
def unrolled_grads_mosaic():
    """unrolled training of a demosaicking objective: spatial-diagonal x-update under autograd."""
    g = torch.Generator().manual_seed(53)
    img = torch.rand(2, 3, 16, 24, generator=g)
    x = dp.Variable()
    b = (dp.mosaic(x).forward(img) + 0.01 * torch.randn(2, 3, 16, 24, generator=g)).detach()
    wgt = torch.rand(2, 3, 16, 24, generator=g)
    b = b.clone().requires_grad_(True)
    x0 = torch.rand(2, 3, 16, 24, generator=g).requires_grad_(True)
    rhos = (0.5 + torch.rand(2, 3, generator=g)).requires_grad_(True)
    lam1 = (0.02 + 0.05 * torch.rand(3, generator=g)).requires_grad_(True)
    f1 = dp.norm1(x)
    solver = dp.compile(dp.sum_squares(dp.mosaic(x) - b) + f1, method="admm", device="cpu")
    out = solver.solve(x0=x0, rhos=rhos, lams={f1: lam1}, max_iter=3)
    (out * wgt).sum().backward()
    return dict(b=_np(b), x0=_np(x0), wgt=_np(wgt), rhos=_np(rhos), lam1=_np(lam1), out=_np(out), g_b=_np(b.grad),
                g_x0=_np(x0.grad), g_rhos=_np(rhos.grad), g_lam1=_np(lam1.grad), T=3)


@case
def unrolled_share_false():
    """UnrolledSolver(share=False) (unroll.py:20-58): per-iteration deep copies of the solver; (i) learned_params=True:
    rhos / lams are nn.Parameters initialised to ones, gradients recorded; (ii) given schedules, trainable per-iteration
    denoisers: gradient w.r.t. the first conv weight of iteration 0's and iteration 1's copy."""
    from dprox.algo.specialization import build_unrolled_solver
    img, psf, b = _deconv_inputs(2, 3, 16, 24)
    g = torch.Generator().manual_seed(61)
    wgt = torch.rand(2, 3, 16, 24, generator=g)
    out = {}
    x = dp.Variable()
    f1 = dp.norm1(x)
    solver = dp.compile(dp.sum_squares(dp.conv(x, psf) - b) + f1, method="admm", device="cpu")
    us = build_unrolled_solver(solver, share=False, max_iter=3, learned_params=True)
    with torch.no_grad():
        us.rhos.copy_(torch.tensor([0.6, 0.9, 1.3]))
        us.lams[f1].copy_(torch.tensor([0.05, 0.03, 0.02]))
    y = us.solve(x0=b, rhos=1.0, lams={f1: 1.0})
    (y * wgt).sum().backward()
    out.update(lp_out=_np(y), lp_g_rhos=_np(us.rhos.grad), lp_g_lam=_np(us.lams[f1].grad))
    # (ii) trainable denoiser copies
    den = _RandFFDNetColor(seed=5)
    x = dp.Variable()
    prior = dp.deep_prior(x, denoiser=den, trainable=True)
    solver = dp.compile(dp.sum_squares(dp.conv(x, psf) - b) + prior, method="admm", device="cpu")
    us = build_unrolled_solver(solver, share=False, max_iter=3)
    sig = torch.tensor([0.1, 0.07, 0.05])
    y = us.solve(x0=b, rhos=torch.tensor([0.6, 0.9, 1.3]), lams={prior: sig})
    (y * wgt).sum().backward()
    w0 = us.solvers[0].psi_fns[0].denoiser.model.model[0].weight
    w2 = us.solvers[1].psi_fns[0].denoiser.model.model[0].weight      # (iteration 2's prox does not reach x)
    out.update(dn_out=_np(y), dn_g_w0=_np(w0.grad), dn_g_w1=_np(w2.grad), sig=_np(sig))
    return dict(psf=psf, b=_np(b), wgt=_np(wgt), seed=5, T=3, **out)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        data = CASES[n]()
        path = os.path.join(OUT, n + ".npz")
        np.savez_compressed(path, **data)
        print(f"{n:28s} -> {os.path.getsize(path) / 1024:8.1f} KiB  keys={sorted(data)[:6]}...")
