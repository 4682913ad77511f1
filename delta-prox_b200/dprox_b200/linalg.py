"""Matrix-free linear solvers (mirrors dprox.linalg).

The operator `A` is any Python callable on CUDA tensors (the plugin surface: LinOp trees, user
BlackBoxes); everything *around* it — dot products, the x/r update, the direction update — runs in
three fused sm_100a kernels (`dpx_cg_dot`, `dpx_cg_update`, `dpx_cg_direction`), so one CG step costs
28 B/element of vector traffic instead of the reference's ~10 full-tensor passes, and there is no
host synchronisation per step: the stop test runs on the device (`dpx_cg_gate`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import partial
from typing import Callable, Optional

import numpy as np
import torch

from . import ops


@dataclass
class LinearSolveConfig:
    """dprox/linalg/custom.py:9-26."""
    rtol: float = 1e-6
    max_iters: int = 100
    verbose: bool = False
    solver_type: str = "cg"
    solver_kwargs: dict = field(default_factory=dict)
    use_analytic_grad: bool = True


class _StopPoll:
    """Lagged, non-blocking read of the device-side `done` flag: every step queues a 4-byte copy into pinned memory behind
    the step's kernels; the host only looks at copies whose event has already fired, so it never waits for the device and
    simply stops enqueueing a couple of steps after the solve froze itself (extra steps are no-ops on x and r)."""

    def __init__(self, device, depth=4):
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self.slots = torch.zeros(depth, dtype=torch.int32).pin_memory()
        self.events = [None] * depth
        self.n = 0

    def push(self):
        k = self.n % len(self.events)
        self.slots[k:k + 1].copy_(self.done, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev
        self.n += 1

    def stopped(self) -> bool:
        for k, ev in enumerate(self.events):
            if ev is not None and ev.query() and int(self.slots[k]) != 0:
                return True
        return False


def cg(A: Callable, b: torch.Tensor, x0: Optional[torch.Tensor] = None, rtol: float = 1e-6, max_iters: int = 100,
       verbose: bool = False, check_every: Optional[int] = None):
    """Conjugate gradients, batched over dim 0 (linalg/solve/solver_cg.py:56-136).

    Stop test: per-sample ||r_b|| <= rtol * ||b_b|| for all b.  (The reference compares the *spectral*
    norm of the [B,n] residual matrix with the per-sample tolerances, solver_cg.py:103-104; for B == 1
    the two coincide — SURVEY App. A-7.)  The test runs ON THE DEVICE (`dpx_cg_gate`): once it holds the
    iterate is frozen at exactly the iteration where the reference breaks, and the host -- which never
    blocks on the device -- stops enqueueing steps as soon as a lagged copy of the flag tells it so.
    `check_every=k > 0` restores a blocking host test every k steps; `check_every=0` never tests.
    """
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    r = b.clone() if x0 is None else ops.axpby(1.0, b, -1.0, A(x))
    n_it = int(min(max_iters, int(np.prod(b.shape))))
    tol2 = None
    poll = None
    if check_every is None or check_every:
        bn2 = ops.dot(b, b)
        tol2 = (rtol * rtol) * bn2                               # B scalars (cold)
        if check_every is None:
            poll = _StopPoll(b.device)
    gamma = ops.dot(r, r)
    p = None
    for it in range(n_it):
        if check_every and it % check_every == 0 and bool(torch.all(gamma <= tol2)):
            if verbose:
                print("Converged at CG Iter %03d" % it)
            break
        if poll is not None and poll.stopped():
            break
        if it == 0:
            p = r.clone()
        q = A(p)
        pq = ops.dot(p, q)
        if poll is not None:
            ops.cg_gate(gamma, tol2, pq, poll.done)             # converged -> pq = inf -> the update below is a no-op
            poll.push()
        gamma_new = ops.cg_update(x, r, p, q, gamma, pq)        # x += a p ; r -= a q ; <r,r>
        ops.cg_direction(p, r, gamma_new, gamma)                # p = r + (g'/g) p
        gamma = gamma_new
    return x


def pcg(A: Callable, b: torch.Tensor, x0: Optional[torch.Tensor] = None, rtol: float = 1e-6, max_iters: int = 100,
        verbose: bool = False, Minv: Optional[Callable] = None, check_every: Optional[int] = None):
    """Preconditioned CG with the reference's conventions (solver_cg.py:172-233): starts from ones,
    whole-tensor dot products, absolute inf-norm stop `max|r| < rtol` -- tested on the device like `cg`'s.

    With `Minv=None` it is algebraically CG started at ones with global dots: the same three fused
    kernels are used with batch=1 (the residual tracked here is b - A x = -r_ref)."""
    x = torch.ones_like(b) if x0 is None else x0.clone()
    r = ops.axpby(1.0, b, -1.0, A(x))
    poll = _StopPoll(b.device) if check_every is None else None
    tol = torch.full((1,), rtol, device=b.device) if poll is not None else None
    if Minv is None:
        gamma = ops.dot(r, r, per_sample=False)
        p = r.clone()
        rmax = None
        for it in range(max_iters):
            if poll is not None and poll.stopped():
                break
            q = A(p)
            pq = ops.dot(p, q, per_sample=False)
            if poll is not None and rmax is not None:
                ops.cg_gate(rmax, tol, pq, poll.done, strict=True)      # the reference breaks right after the update that met the test
                poll.push()
            gamma_new = ops.cg_update(x, r, p, q, gamma, pq, per_sample=False)
            ops.cg_direction(p, r, gamma_new, gamma, per_sample=False)
            gamma = gamma_new
            if poll is not None:
                rmax = ops.absmax(r)
            elif check_every and (it + 1) % check_every == 0 and float(ops.absmax(r)) < rtol:
                break
        return x
    # general preconditioner: y = Minv(r) is a user callable; dots/axpys stay native (host test: Minv is user code anyway)
    check_every = 1 if check_every is None else check_every
    y = Minv(r)
    p = y.clone()
    ry = ops.dot(r, y, per_sample=False)
    for it in range(max_iters):
        q = A(p)
        pq = ops.dot(p, q, per_sample=False)
        alpha = ry / pq
        x = ops.lincomb(x, None, p, alpha)
        r = ops.lincomb(r, None, q, -alpha)
        y = Minv(r)
        ry_new = ops.dot(r, y, per_sample=False)
        p = ops.lincomb(y, None, p, ry_new / ry)
        ry = ry_new
        if check_every and (it + 1) % check_every == 0 and float(ops.absmax(r)) < rtol:
            break
    return x


SOLVERS = {"cg": cg, "pcg": pcg}


def _build_solver(config: LinearSolveConfig):
    if config.solver_type not in SOLVERS:
        raise NotImplementedError(f"solver_type={config.solver_type!r}: only {sorted(SOLVERS)} are lowered "
                                  f"(minres/plss are out of scope, SURVEY §2 row 22)")
    return partial(SOLVERS[config.solver_type], rtol=config.rtol, max_iters=config.max_iters, verbose=config.verbose,
                   **config.solver_kwargs)


class ImplicitSolve(torch.autograd.Function):
    """LinearSolve (linalg/custom.py:39-62): x = solve(A, b) with the implicit-differentiation backward
    grad_b = solve(A^T, grad_x) -- a second run of the same fused-kernel CG (A is the symmetric normal-equation
    operator, so A^T = A, sum_square.py:175-177).  Like the reference, whose KtK module multiplies by the closure `rho`
    rather than by its own parameter (sum_square.py:160-173), nothing is propagated to quantities inside A: rho and the
    measurements receive their gradients through the right-hand side only."""

    @staticmethod
    def forward(ctx, A, b, config):
        ctx.A, ctx.config = A, config
        return _build_solver(config)(A, b)

    @staticmethod
    def backward(ctx, g):
        return None, _build_solver(ctx.config)(ctx.A, g.contiguous()), None


def linear_solve(A: Callable, b: torch.Tensor, config: LinearSolveConfig = LinearSolveConfig()):
    """Solve A x = b matrix-free (linalg/custom.py:65-82); differentiable w.r.t. b by implicit differentiation."""
    if config.use_analytic_grad and torch.is_grad_enabled() and b.requires_grad:
        return ImplicitSolve.apply(A, b, config)
    return _build_solver(config)(A, b)
