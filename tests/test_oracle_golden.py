"""Pins the oracle (oracle/dprox_oracle.py) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only; this is the "oracle is trustworthy" gate."""
import os

import numpy as np
import pytest
import torch

import dprox_oracle as orc
from conftest import GOLDEN

torch.set_num_threads(4)


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def T(a):
    return torch.from_numpy(np.asarray(a))


def rel(a, b):
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a, b = np.asarray(a, dtype=dt), np.asarray(b, dtype=dt)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def check_state(state, g, tol):
    names = ["s0", "s1", "s2"]
    for name, s in zip(names, state):
        if isinstance(s, (list, tuple)):
            for i, e in enumerate(s):
                assert rel(e.numpy(), g[f"{name}_{i}"]) < tol, (name, i, rel(e.numpy(), g[f"{name}_{i}"]))
        else:
            assert rel(s.numpy(), g[name]) < tol, (name, rel(s.numpy(), g[name]))


def deconv_terms(g, psi):
    data = orc.Term("sum_squares", orc.Conv(g["psf"] if "psf" in g else g["kernel"], orc.Identity()), c=T(g["b"]))
    return [data] + psi


@pytest.mark.parametrize("case,method", [("admm_conv_nonneg", "admm"), ("hqs_conv_nonneg", "hqs"),
                                         ("ladmm_conv_nonneg_b1", "ladmm"), ("vxu_conv_nonneg_b1", "admm_vxu"),
                                         ("admm_even_kernel", "admm")])
def test_headline_objective(case, method):
    g = load(case)
    s = orc.Solver(deconv_terms(g, [orc.Term("nonneg")]), method)
    st = s.solve(T(g["b"]), rhos=float(g["rho"]) if "rho" in g else None, max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 2e-6)


def test_pock_chambolle():
    g = load("pc_conv_nonneg")
    s = orc.Solver(deconv_terms(g, [orc.Term("nonneg"), orc.Term("norm1")]), "pc")
    st = s.solve(T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 2e-6)


def test_admm_50_iterations():
    g = load("admm_conv_nonneg_50it")
    s = orc.Solver(deconv_terms(g, [orc.Term("nonneg")]), "admm")
    st = s.solve(T(g["b"]), rhos=0.5, lams=0.02, max_iter=50, return_full_states=True)
    check_state(st, g, 1e-5)
    # and the fp64 arbiter agrees with the reference to the reference's own fp32 noise
    s64 = orc.Solver(deconv_terms(g, [orc.Term("nonneg")]), "admm", dtype=torch.float64)
    x64 = s64.solve(T(g["b"]).double(), rhos=0.5, lams=0.02, max_iter=50)
    assert rel(g["s0"], x64.numpy()) < 2e-5


@pytest.mark.parametrize("case,kind", [("pgd_conv_nonneg", "nonneg"), ("pgd_conv_norm1", "norm1")])
def test_pgd(case, kind):
    g = load(case)
    data = orc.Term("sum_squares", orc.Conv(g["psf"], orc.Identity()), c=T(g["b"]))
    s = orc.Solver([data, orc.Term(kind)], "pgd")
    st = s.solve(T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]) if "lam" in g else None,
                 max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 2e-6)


def test_two_psi_per_sample_schedules():
    g = load("admm_two_psi_per_sample")
    f1, f2 = orc.Term("norm1", alpha=float(g["alpha1"])), orc.Term("nonneg")
    s = orc.Solver(deconv_terms(g, [f1, f2]), "admm")
    st = s.solve(T(g["b"]), rhos=T(g["rhos"]), lams={f1: T(g["lam1"]), f2: T(g["lam2"])}, max_iter=int(g["T"]),
                 return_full_states=True)
    check_state(st, g, 2e-6)


def test_hqs_two_psi():
    g = load("hqs_two_psi_norm2")
    f1, f2 = orc.Term("norm2"), orc.Term("norm1")
    s = orc.Solver(deconv_terms(g, [f1, f2]), "hqs")
    st = s.solve(T(g["b"]), rhos=float(g["rho"]), lams={f1: float(g["lam1"]), f2: float(g["lam2"])},
                 max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 2e-6)


def test_psi_offset():
    g = load("admm_psi_offset")
    s = orc.Solver(deconv_terms(g, [orc.Term("norm1", c=T(g["c"]))]), "admm")
    st = s.solve(T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 2e-6)


@pytest.mark.parametrize("case,method,extra", [("admm_tv", "admm", []), ("hqs_tv_nonneg", "hqs", ["nonneg"])])
def test_tv(case, method, extra):
    g = load(case)
    psi = [orc.Term("norm1", orc.Grad(0, orc.Identity())), orc.Term("norm1", orc.Grad(1, orc.Identity()))]
    psi += [orc.Term(k) for k in extra]
    s = orc.Solver(deconv_terms(g, psi), method)
    st = s.solve(T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 3e-6)


def test_linops():
    g = load("linops")
    t = T(g["t"])
    ident = orc.Identity()
    ops = {"conv": orc.Conv(g["psf"], ident), "conv2": orc.Conv(g["k2"], ident), "grad0": orc.Grad(0, ident),
           "grad1": orc.Grad(1, ident), "grad2": orc.Grad(2, ident)}
    for name, op in ops.items():
        assert rel(op.fwd(t).numpy(), g[name + "_fwd"]) < 1e-6, name
        assert rel(op.adj(t).numpy(), g[name + "_adj"]) < 1e-6, name
        assert rel(op.FB(t.shape).numpy(), g[name + "_otf"]) < 1e-6, name
        assert op.FB(t.shape).numpy().dtype == g[name + "_otf"].dtype, name
        assert rel(op.diag(t, True).numpy(), g[name + "_diag"]) < 1e-6, name
    assert rel(orc.Mosaic(ident).fwd(t).numpy(), g["mosaic_fwd"]) == 0.0
    # grad semantics in pixel space: forward difference with circular wrap
    g0 = orc.Grad(0, ident).fwd(t)
    assert torch.allclose(g0, torch.roll(t, -1, dims=-2) - t, atol=2e-6)
    g1 = orc.Grad(1, ident).fwd(t)
    assert torch.allclose(g1, torch.roll(t, -1, dims=-1) - t, atol=2e-6)


def test_spatial_diag_paths():
    g = load("admm_mosaic_spatial")
    data = orc.Term("sum_squares", orc.Mosaic(orc.Identity()), c=T(g["b"]))
    s = orc.Solver([data, orc.Term("nonneg")], "admm")
    assert (s.ls.diagonalizable, s.ls.freq_diagonalizable) == tuple(bool(v) for v in g["flags"])
    st = s.solve(T(g["b"]), rhos=float(g["rho"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 2e-6)

    g = load("hqs_mul_elementwise")
    data = orc.Term("sum_squares", orc.Mul(T(g["w"]), orc.Identity()), c=T(g["b"]))
    s = orc.Solver([data, orc.Term("norm1")], "hqs")
    st = s.solve(T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 2e-6)


@pytest.mark.parametrize("case,solver", [("admm_cg_mosaic_conv", "cg"), ("admm_pcg_mosaic_conv", "pcg")])
def test_cg_fallback(case, solver):
    g = load(case)
    data = orc.Term("sum_squares", orc.Mosaic(orc.Conv(g["psf"], orc.Identity())), c=T(g["b"]))
    s = orc.Solver([data, orc.Term("nonneg")], "admm", solver_type=solver, rtol=1e-6, max_iters=int(g["cg_iters"]))
    assert not s.ls.diagonalizable and not s.ls.freq_diagonalizable
    st = s.solve(T(g["b"]), rhos=float(g["rho"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 1e-5)


def test_ladmm_with_grad_terms_reference_semantics():
    g = load("ladmm_tv_3it")
    psi = [orc.Term("norm1", orc.Grad(0, orc.Identity())), orc.Term("norm1", orc.Grad(1, orc.Identity()))]
    s = orc.Solver(deconv_terms(g, psi), "ladmm")
    st = s.solve(T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 3e-6)


def csmri_ops(mask):
    fwd = lambda x, step=0: mask * torch.fft.fft2(x, norm="ortho")
    adj = lambda y, step=0: torch.real(torch.fft.ifft2(mask * y, norm="ortho"))
    return fwd, adj


@pytest.mark.parametrize("solver", ["cg", "pcg"])
def test_csmri_blackbox_cg(solver):
    g = load("admm_csmri_blackbox")
    fwd, adj = csmri_ops(T(g["mask"]))
    y0 = torch.complex(T(g["y0_re"]), T(g["y0_im"]))
    data = orc.Term("sum_squares", orc.BlackBox(fwd, adj, orc.Identity()), b=y0)
    psi = [orc.Term("norm1", orc.Grad(0, orc.Identity())), orc.Term("norm1", orc.Grad(1, orc.Identity()))]
    s = orc.Solver([data] + psi, "admm", solver_type=solver, rtol=1e-6, max_iters=int(g["cg_iters"]))
    st = s.solve(T(g["x0"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, {k[len(solver) + 1:]: v for k, v in g.items() if k.startswith(solver + "_s")}, 1e-5)


def test_linear_solvers_known_answers():
    g = load("linear_solvers")
    A, b = T(g["A"]), T(g["b"])
    for name, fn in (("cg", orc.cg), ("pcg", orc.pcg)):
        x = fn(lambda v: A @ v, b, rtol=1e-8, max_iters=100)
        assert np.allclose(x.numpy(), g[name], rtol=1e-10, atol=1e-12)
        assert np.allclose(x.numpy(), g["x_true"], rtol=1e-6)
    cv = orc.Conv(g["psf"], orc.Identity())
    Aop = lambda v: v + 0.5 * cv.adj(cv.fwd(v))
    rhs = T(g["rhs"])
    assert rel(orc.cg(Aop, rhs, rtol=1e-6, max_iters=12).numpy(), g["cg_conv"]) < 2e-6
    assert rel(orc.pcg(Aop, rhs, rtol=1e-6, max_iters=12).numpy(), g["pcg_conv"]) < 2e-6


def test_ml_problems_known_answers():
    """tests/problem/test_ml_problems.py:5-44 — x == rhs/2 exactly; conv residual < 1e-5."""
    g = load("ml_problems")
    rhs = T(g["rhs"]).float().reshape(1, 1, 3, 3)
    for key, term in (("lsq", orc.Term("sum_squares", orc.Scale(2, orc.Identity()), c=rhs)),
                      ("lsq1", orc.Term("sum_squares", orc.Scale(2, orc.Identity()), b=rhs))):
        x = orc.Solver([term], "admm").solve(torch.zeros(1, 1, 3, 3))
        assert (x.numpy().reshape(3, 3) == g["rhs"] / 2).all()
        assert (x.numpy().reshape(g[key].shape) == g[key]).all()
    rhs2 = T(g["rhs2"]).float().reshape(1, 1, 3, 3)           # (3,3,1) HWC -> [1,1,3,3]; rhs2 is (1,3,3)
    term = orc.Term("sum_squares", orc.Conv(g["kernel"].astype("float32"), orc.Identity()), c=rhs2)
    x = orc.Solver([term], "admm").solve(torch.zeros(1, 1, 3, 3))
    assert rel(x.numpy().ravel(), g["lsq2"].ravel()) < 1e-5
    assert (np.abs(term.K(x).numpy()) < 1e-4).all()


def test_conv_doe():
    g = load("hqs_conv_doe")
    for tag in ("full", "padded"):
        psf = T(g[f"{tag}_psf"])
        assert rel(orc.psf2otf2(psf, (2, 3, 32, 32)).numpy(), g[f"{tag}_otf"]) < 1e-6
        data = orc.Term("sum_squares", orc.ConvDOE(psf, orc.Identity()), c=T(g[f"{tag}_b"]))
        s = orc.Solver([data, orc.Term("nonneg")], "hqs")
        st = s.solve(T(g[f"{tag}_b"]), rhos=float(g["rho"]), max_iter=int(g["T"]), return_full_states=True)
        check_state(st, {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + "_s")}, 3e-6)


def test_ffdnet_and_deep_prior():
    g = load("ffdnet_forward")
    ws = orc.ffdnet_random_weights(4)
    assert abs(float(ws[0][0].double().sum()) - float(g["w0_sum"])) < 1e-9      # RNG reproducibility
    assert abs(float(ws[-1][0].double().sum()) - float(g["wlast_sum"])) < 1e-9
    y = orc.ffdnet_forward(ws, T(g["x"]), T(g["sigma"]))
    assert rel(y.numpy(), g["y"]) < 1e-5

    for case in ("admm_deep_prior_ffdnet", "admm_deep_prior_wellcond"):
        g = load(case)
        prior = orc.Term("deep_prior", denoiser=lambda v, s: orc.ffdnet_forward(ws, v, s))
        nn_ = orc.Term("nonneg")
        s = orc.Solver(deconv_terms(g, [prior, nn_]), "admm")
        rhos = T(g["rhos"]) if "rhos" in g else float(g["rho"])
        st = s.solve(T(g["b"]), rhos=rhos, lams={prior: T(g["sigmas"]), nn_: 0.02}, max_iter=int(g["T"]),
                     return_full_states=True)
        check_state(st, g, 1e-5)


def test_schedules():
    g = load("schedules")
    r1, s1 = orc.log_descent(35, 30, 24)
    r2, s2 = orc.log_descent(49, 7.65, 10, sigma=7.65 / 255, sqrt=True)
    for a, b in ((r1, g["r1"]), (s1, g["s1"]), (r2, g["r2"]), (s2, g["s2"])):
        assert np.allclose(a.numpy(), b, rtol=1e-6)


def test_csmri_closed_form_and_custom_admm():
    """§8f-2: csmri._prox and the CustomADMM loop on complex iterates vs the reference (even and odd sizes: the centred
    transforms shift by n//2, which differs between fftshift and ifftshift for odd n)."""
    g = dict(np.load(os.path.join(GOLDEN, "csmri_custom_admm.npz"), allow_pickle=False))
    ws = orc.ffdnet_random_weights(int(g["seed"]), in_nc=1)
    den = lambda v, s_: orc.ffdnet_forward(ws, v, s_)
    for tag in ("even", "odd"):
        mask, y0, x0 = (torch.from_numpy(g[f"{tag}_{k}"]) for k in ("mask", "y0", "x0"))
        one = orc.csmri_prox(x0 * (1 + 0.5j), torch.tensor([0.7, 1.3]), 1, mask, y0)
        assert np.abs(one.numpy() - g[f"{tag}_prox1"]).max() < 2e-6
        with torch.no_grad():
            x, z, u = orc.custom_admm_csmri(den, mask, y0, x0, torch.from_numpy(g["rhos"]), torch.from_numpy(g["sigmas"]))
        for got, name in ((x, "x"), (z, "z"), (u, "u")):
            want = g[f"{tag}_{name}"]
            assert np.linalg.norm(got.numpy() - want) / np.linalg.norm(want) < 1e-5, (tag, name)


# ---- round 2: reference vectors at fused-engine sizes, BASELINE cfg1 / cfg3 as stated, formerly partial rows --------

def test_cfg1_256_50_iterations():
    """BASELINE configs[0] ([1,3,256,256], 50 it): the reference's own fp32 round-off is ~1e-5 at 50 iterations, so the
    oracle is held to it at 2e-5 and both are compared with the fp64 run."""
    g = load("cfg1_admm_256_50it")
    x = orc.Solver(deconv_terms(g, [orc.Term("nonneg")]), "admm").solve(T(g["b"]), rhos=1.0, lams=0.02, max_iter=50)
    x64 = orc.Solver([orc.Term("sum_squares", orc.Conv(g["psf"], orc.Identity()), c=T(g["b"]).double()), orc.Term("nonneg")],
                     "admm", dtype=torch.float64).solve(T(g["b"]).double(), rhos=1.0, lams=0.02, max_iter=50)
    assert rel(x.numpy(), g["s0"]) < 2e-5
    assert rel(x.numpy(), x64.numpy()) < 1e-5 and rel(g["s0"], x64.numpy()) < 2e-5


@pytest.mark.parametrize("case,method", [("admm_fused_128x192", "admm"), ("hqs_fused_64x320", "hqs")])
def test_fused_engine_size_goldens(case, method):
    g = load(case)
    st = orc.Solver(deconv_terms(g, [orc.Term("nonneg")]), method).solve(T(g["b"]), rhos=float(g["rho"]), lams=0.02,
                                                                         max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 3e-6)


def test_cfg3_blackbox_tv_ladmm_pcg():
    g = load("ladmm_csmri_blackbox_3it")
    fwd, adj = csmri_ops(T(g["mask"]))
    y0 = torch.complex(T(g["y0_re"]), T(g["y0_im"]))
    data = orc.Term("sum_squares", orc.BlackBox(fwd, adj, orc.Identity()), b=y0)
    psi = [orc.Term("norm1", orc.Grad(0, orc.Identity())), orc.Term("norm1", orc.Grad(1, orc.Identity()))]
    s = orc.Solver([data] + psi, "ladmm", solver_type="pcg", rtol=1e-6, max_iters=int(g["cg_iters"]))
    st = s.solve(T(g["x0"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 1e-5)


def test_conv_doe_linear():
    g = load("conv_doe_linear")
    for tag in ("even", "small", "odd"):
        psf, t = T(g[f"{tag}_psf"]), T(g[f"{tag}_t"])
        op = orc.ConvDOE(psf, orc.Identity(), circular=False)
        assert rel(op.fwd(t).numpy(), g[f"{tag}_fwd"]) < 1e-6 and rel(op.adj(t).numpy(), g[f"{tag}_adj"]) < 1e-6
        data = orc.Term("sum_squares", op, c=T(g[f"{tag}_b"]))
        st = orc.Solver([data, orc.Term("nonneg")], "hqs").solve(T(g[f"{tag}_b"]), rhos=float(g["rho"]), max_iter=int(g["T"]),
                                                               return_full_states=True)
        check_state(st, {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + "_s")}, 3e-6)


def test_img_psf_conv_linear():
    g = load("img_psf_conv_linear")
    for tag in ("same", "small"):
        img, psf = T(g[f"{tag}_img"]).requires_grad_(True), T(g[f"{tag}_psf"]).requires_grad_(True)
        y = orc.img_psf_conv(img, psf, circular=False)
        (y * T(g[f"{tag}_w"])).sum().backward()
        assert rel(y.detach().numpy(), g[f"{tag}_y"]) < 1e-6
        assert rel(img.grad.numpy(), g[f"{tag}_g_img"]) < 1e-6 and rel(psf.grad.numpy(), g[f"{tag}_g_psf"]) < 1e-6


def test_grad_channel_axis():
    g = load("admm_grad_dim2")
    psi = [orc.Term("norm1", orc.Grad(d, orc.Identity())) for d in (2, 1, 0)]
    st = orc.Solver(deconv_terms(g, psi), "admm").solve(T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]),
                                                        max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 3e-6)


def test_vxu_with_deep_prior_and_x8():
    ws = orc.ffdnet_random_weights(4)
    den = lambda v, s: orc.ffdnet_forward(ws, v, s)
    g = load("vxu_deep_prior")
    prior, nn_ = orc.Term("deep_prior", denoiser=den), orc.Term("nonneg")
    st = orc.Solver(deconv_terms(g, [prior, nn_]), "admm_vxu").solve(T(g["b"]), rhos=float(g["rho"]),
                                                                     lams={prior: T(g["sigmas"]), nn_: 0.02}, max_iter=int(g["T"]),
                                                                     return_full_states=True)
    check_state(st, g, 1e-5)
    g = load("deep_prior_x8")
    prior, nn_ = orc.Term("deep_prior", denoiser=orc.Augment(den)), orc.Term("nonneg")
    st = orc.Solver(deconv_terms(g, [prior, nn_]), "admm").solve(T(g["b"]), rhos=float(g["rho"]),
                                                                 lams={prior: T(g["sigmas"]), nn_: 0.02}, max_iter=int(g["T"]),
                                                                 return_full_states=True)
    check_state(st, g, 1e-5)


def test_advisor_cases():
    g = load("pc_identity")
    st = orc.Solver([orc.Term("sum_squares", c=T(g["b"])), orc.Term("norm1")], "pc").solve(
        T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 2e-6)
    g = load("admm_mask_psi")
    st = orc.Solver([orc.Term("sum_squares", c=T(g["b"])), orc.Term("norm1", orc.Mosaic(orc.Identity()))], "admm").solve(
        T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 2e-6)
    g = load("ladmm_scaled_identity")
    st = orc.Solver(deconv_terms(g, [orc.Term("norm1", orc.Scale(2.0, orc.Identity()))]), "ladmm").solve(
        T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]), return_full_states=True)
    check_state(st, g, 3e-6)
    g = load("pgd_psi_linop")
    data = lambda: orc.Term("sum_squares", orc.Conv(g["psf"], orc.Identity()), b=T(g["b"]))
    x = orc.Solver([data(), orc.Term("norm1", orc.Grad(1, orc.Identity()))], "pgd").solve(
        T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]))
    assert rel(x.numpy(), g["s0"]) < 3e-6
    x = orc.Solver([data(), orc.Term("norm1", orc.Scale(2.0, orc.Identity()))], "pgd").solve(
        T(g["b"]), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=int(g["T"]))
    assert rel(x.numpy(), g["scaled_s0"]) < 3e-6
