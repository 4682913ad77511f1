"""dprox_b200 — Blackwell-native proximal-iteration backend with the Delta-Prox (`dprox`) API.

`from dprox_b200 import *` mirrors `from dprox import *` for the hot path (dprox/__init__.py:1-9): like
the reference it deliberately shadows the builtins `sum`, `eval`, `compile` and `copy`.
"""
from . import linalg  # noqa: F401
from .algo import (ADMM, ADMM_vxu, HQS, Algorithm, LinearizedADMM, PockChambolle, Problem, ProximalGradientDescent, ResidualStop,
                   SOLVERS, UnrolledSolver, build_unrolled_solver, compile, log_descent, specialize)
from .linalg import LinearSolveConfig, linear_solve
from .linop import (BlackBox, CompGraph, Constant, LinOp, LinOpFactory, Placeholder, Variable, adjoint, conv, conv_doe,
                    copy, eval, grad, grad2d, gram, mosaic, mul_elementwise, scale, split, sum, validate, vstack)
from . import contrib  # noqa: F401
from .proxfn import (Denoiser, ProxFn, box, csmri, deep_prior, ext_sum_squares, iso_tv, nonneg, norm1, norm2, sum_squares)
from .tensors import array, tensor

__version__ = "0.1.0"

__all__ = [
    "ADMM", "ADMM_vxu", "HQS", "Algorithm", "LinearizedADMM", "PockChambolle", "Problem", "ProximalGradientDescent", "ResidualStop",
    "SOLVERS", "UnrolledSolver", "build_unrolled_solver", "compile", "log_descent", "specialize", "LinearSolveConfig", "linear_solve", "linalg", "contrib",
    "BlackBox", "CompGraph", "Constant", "LinOp", "LinOpFactory", "Placeholder", "Variable", "adjoint", "conv", "conv_doe",
    "copy", "eval", "grad", "grad2d", "gram", "mosaic", "mul_elementwise", "scale", "split", "sum", "validate", "vstack",
    "Denoiser", "ProxFn", "box", "csmri", "deep_prior", "ext_sum_squares", "iso_tv", "nonneg", "norm1", "norm2", "sum_squares",
    "array", "tensor",
]
