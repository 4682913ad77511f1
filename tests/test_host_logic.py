"""Host-side logic of the dprox-compatible surface, runnable without a GPU: objective partitioning, the
compile()-time lowering decisions (which must reproduce the reference's diagonalisability analysis,
proxfn/sum_square.py:104-107), argument defaults, error behaviour and the no-CPU-fallback rule."""
import numpy as np
import pytest
import torch

import dprox_b200 as dp
from dprox_b200 import _cabi as cabi
from dprox_b200.linop import kernel_otf
import dprox_oracle as orc


def objective(x, kind="conv", psi=("nonneg",)):
    b = torch.rand(2, 3, 32, 48)
    psf = orc.point_spread_function(5, 2)
    data = {"conv": lambda: dp.sum_squares(dp.conv(x, psf) - b), "ident": lambda: dp.sum_squares(2 * x - b),
            "mosaic": lambda: dp.sum_squares(dp.mosaic(x) - b), "joint": lambda: dp.sum_squares(dp.mosaic(dp.conv(x, psf)) - b)}[kind]()
    regs = {"nonneg": lambda: dp.nonneg(x), "l1": lambda: dp.norm1(x), "tvh": lambda: dp.norm1(dp.grad(x, dim=0)),
            "tvw": lambda: dp.norm1(dp.grad(x, dim=1)), "convpsi": lambda: dp.norm1(dp.conv(x, psf))}
    fns = data
    for p in psi:
        fns = fns + regs[p]()
    return fns


@pytest.mark.parametrize("method,kind,psi,tier,xupdate", [
    ("admm", "conv", ("nonneg",), "native", "freq"), ("hqs", "conv", ("nonneg", "l1"), "native", "freq"),
    ("ladmm", "conv", ("nonneg",), "native", "freq"), ("admm_vxu", "conv", ("nonneg",), "native", "freq"),
    ("pgd", "conv", ("nonneg",), "native", "freq"), ("admm", "conv", ("tvh", "tvw"), "native", "freq"),
    ("ladmm", "conv", ("tvh", "tvw"), "generic", "freq"), ("admm", "conv", ("convpsi",), "generic", "freq"),
    ("admm", "mosaic", ("nonneg",), "native", "spatial"), ("admm", "ident", ("nonneg",), "native", "scalar"),
    ("admm", "joint", ("nonneg",), "generic", "cg")])
def test_lowering_decisions(method, kind, psi, tier, xupdate):
    x = dp.Variable()
    s = dp.compile(objective(x, kind, psi), method=method, device="cpu")
    assert (s.spec.tier, s.spec.xupdate) == (tier, xupdate), s.spec.reason


def test_diagonalisability_flags_match_reference_rules():
    x = dp.Variable()
    s = dp.compile(objective(x, "conv"), device="cpu")
    assert (s.least_square.diagonalizable, s.least_square.freq_diagonalizable) == (False, True)
    s = dp.compile(objective(x, "mosaic"), device="cpu")
    assert (s.least_square.diagonalizable, s.least_square.freq_diagonalizable) == (True, False)
    s = dp.compile(objective(x, "ident"), device="cpu")
    assert (s.least_square.diagonalizable, s.least_square.freq_diagonalizable) == (True, True)
    s = dp.compile(objective(x, "conv"), device="cpu", try_freq_diagonalize=False)
    assert s.spec.xupdate == "cg" and not s.least_square.freq_diagonalizable
    s = dp.compile(objective(x, "joint"), device="cpu")
    assert (s.least_square.diagonalizable, s.least_square.freq_diagonalizable) == (False, False)


def test_partition_rules():
    x = dp.Variable()
    f = objective(x, "conv", ("nonneg", "l1"))
    psi, omega = dp.ADMM.partition(f)
    assert [type(t).__name__ for t in psi] == ["nonneg", "norm1"] and [type(t).__name__ for t in omega] == ["sum_squares"]
    with pytest.raises(ValueError):
        dp.ProximalGradientDescent.partition(f)                       # needs exactly two fns (pgd.py:11-13)
    with pytest.raises(ValueError):
        dp.ProximalGradientDescent.partition([dp.nonneg(x), dp.norm1(x)])     # none differentiable (pgd.py:21-24)
    with pytest.raises(ValueError):
        dp.compile(f, method="nope")


def test_defaults_and_schedules():
    x = dp.Variable()
    f = objective(x, "conv", ("nonneg", "l1"))
    s = dp.compile(f, device="cpu")
    _, rhos, lams, T = s.defaults(None, None, None, 24)
    assert T == 24 and rhos.shape == (24,) and float(rhos[0]) == 1.0
    assert set(lams) == set(s.psi_fns) and all(float(v[3]) == pytest.approx(0.02) for v in lams.values())
    with pytest.raises(KeyError):
        s.defaults(None, None, {s.psi_fns[0]: 0.1}, 5)               # dict must cover every psi fn (admm.py:56)
    r, sg = dp.log_descent(35, 30, 24)
    ro, so = orc.log_descent(35, 30, 24)
    assert torch.allclose(r, ro) and torch.allclose(sg, so)
    assert s.state_split == [1, [2], [2]] and s.state_dim == 5 and s.nparams == 3


def test_prox_scaling_sets_alpha():
    x = dp.Variable()
    f = 0.5 * dp.norm1(x)
    assert f.alpha == 0.5 and f.is_native() and f.native_kind == cabi.PROX_L1
    with pytest.raises(TypeError):
        -1 * dp.norm1(x)


def test_no_cpu_fallback():
    x = dp.Variable()
    s = dp.compile(objective(x, "conv"), device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        s.solve(x0=torch.rand(2, 3, 32, 48))
    with pytest.raises(RuntimeError, match="CUDA"):
        dp.conv(x, np.ones((3, 3, 1), "float32")).forward(torch.rand(1, 3, 8, 8))


def test_autograd_contract_routes_to_the_differentiable_engine():
    """Inputs that require grad are never silently detached: they select the differentiable engine (which,
    like everything else, needs a CUDA device), also for the CG x-update (implicit differentiation, linalg.ImplicitSolve)."""
    x = dp.Variable()
    s = dp.compile(objective(x, "conv"), device="cpu")
    rhos = torch.ones(4, requires_grad=True)
    assert s._wants_grad(rhos) and not s._wants_grad(torch.ones(4))
    with torch.no_grad():
        assert not s._wants_grad(rhos)
    with pytest.raises(RuntimeError, match="CUDA"):
        s.solve(x0=torch.rand(2, 3, 32, 48), rhos=rhos, max_iter=4)
    s2 = dp.compile(dp.sum_squares(dp.mosaic(dp.conv(x, np.ones((3, 3, 1), "float32"))) - torch.rand(1, 3, 8, 8)) + dp.nonneg(x),
                    device="cpu")
    assert s2.spec.xupdate == "cg"
    with pytest.raises(RuntimeError, match="CUDA"):
        s2.solve(x0=torch.rand(1, 3, 8, 8), rhos=rhos, max_iter=4)


def test_kernel_otf_matches_reference_construction():
    g = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "linops.npz"))
    for name, k in (("conv", g["psf"]), ("conv2", g["k2"])):
        otf = kernel_otf(k.astype("float32"), 16, 24, 3)
        assert np.abs(otf - g[name + "_otf"][0]).max() < 1e-6
    x = dp.Variable()
    gr = dp.grad(x, dim=0)
    assert np.abs(gr._FB((2, 3, 16, 24)).numpy() - g["grad0_otf"]).max() < 1e-12


def test_layout_conventions():
    from dprox_b200.tensors import to_torch_tensor, as_bchw
    assert to_torch_tensor(np.zeros((5, 7, 3)), batch=True).shape == (1, 3, 5, 7)        # HWC -> BCHW
    assert to_torch_tensor(np.zeros((5, 7)), batch=True).shape == (1, 5, 7)
    t = dp.tensor(np.zeros((2, 3, 4, 4)))
    assert to_torch_tensor(t, batch=True) is t                                           # dp.tensor is never re-batched
    assert as_bchw(torch.zeros(2, 5, 7)).shape == (2, 1, 5, 7) and as_bchw(torch.zeros(1, 3)).shape == (1, 1, 1, 3)
