"""Matrix-free linear solvers (mirrors dprox.linalg).

The operator `A` is any Python callable on CUDA tensors (the plugin surface: LinOp trees, user
BlackBoxes); everything *around* it — dot products, the x/r update, the direction update — runs in
three fused sm_100a kernels (`dpx_cg_dot`, `dpx_cg_update`, `dpx_cg_direction`), so one CG step costs
28 B/element of vector traffic instead of the reference's ~10 full-tensor passes, and there is no
host synchronisation per step: the stop test runs on the device (`dpx_cg_gate`).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from functools import partial
from typing import Callable, Optional

import numpy as np
import torch

from . import _cabi as cabi
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


# CG / PCG steps enqueued by this process (eager or replayed; a few steps past convergence are gated no-ops): what bench.py's
# cfg3 line counts its algorithmic bytes from -- a replayed step does not pass through the Python operator tree
steps_enqueued = [0]
# kernels of this library launched from replayed graphs (dpx_launch_count only sees launches issued through the C-ABI by the host)
replayed_launches = [0]


class _Replay:
    """A captured CG step plus the number of this library's launches it holds (counted while it was recorded)."""

    def __init__(self, graph, n_launches):
        self.graph, self.n_launches = graph, n_launches

    def replay(self):
        self.graph.replay()
        replayed_launches[0] += self.n_launches


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
        if self.events[k] is not None:
            # the slot's previous copy must have landed before it is overwritten: with graph-replayed steps the host runs ahead
            # of the device, and this wait also keeps it at most `depth` steps ahead (steps queued past convergence are no-ops,
            # but not free)
            self.events[k].synchronize()
            if int(self.slots[k]) != 0:
                self._stop = True
        self.slots[k:k + 1].copy_(self.done, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev
        self.n += 1

    def reset(self):
        """new solve on the same (graph-captured) flag: drain the copies still in flight, clear the flag"""
        for ev in self.events:
            if ev is not None:
                ev.synchronize()
        self.events = [None] * len(self.events)
        self.n, self._stop = 0, False
        self.done.zero_()

    def stopped(self) -> bool:
        if getattr(self, "_stop", False):
            return True
        for k, ev in enumerate(self.events):
            if ev is not None and ev.query() and int(self.slots[k]) != 0:
                return True
        return False


def _capture(body: Callable):
    """One CG step as a CUDA graph.  A step is ~10 small launches (operator nodes, three fused vector kernels, the gate) issued
    from Python: at the reference's own problem sizes (256 x 256: 0.26 MB per vector) it is launch- and interpreter-bound, not
    bandwidth-bound.  `body` has already run once eagerly (plans, lazy buffers and constants exist), so it is recorded once
    and REPLAYED for the remaining steps: one graph launch per step instead of the Python walk over the operator tree.
    Returns None -- and the caller keeps stepping eagerly -- when the user's operator is not capturable (host
    synchronisation, `.item()`, pageable copies ...) or `DPX_CG_GRAPH=0`."""
    if os.environ.get("DPX_CG_GRAPH", "1") == "0" or torch.cuda.is_current_stream_capturing():
        return None
    try:
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = cabi.lib().dpx_launch_count()
        with torch.cuda.graph(g):
            body()
        return _Replay(g, int(cabi.lib().dpx_launch_count() - n0))
    except Exception:                                            # noqa: BLE001 -- any failure means "run eagerly"
        try:
            torch.cuda.synchronize()
        except Exception:                                        # noqa: BLE001
            pass
        return None


def cg(A: Callable, b: torch.Tensor, x0: Optional[torch.Tensor] = None, rtol: float = 1e-6, max_iters: int = 100,
       verbose: bool = False, check_every: Optional[int] = None, cache: Optional[dict] = None, cache_key=None):
    """Conjugate gradients, batched over dim 0 (linalg/solve/solver_cg.py:56-136).

    Stop test: per-sample ||r_b|| <= rtol * ||b_b|| for all b.  (The reference compares the *spectral*
    norm of the [B,n] residual matrix with the per-sample tolerances, solver_cg.py:103-104; for B == 1
    the two coincide — SURVEY App. A-7.)  The test runs ON THE DEVICE (`dpx_cg_gate`): once it holds the
    iterate is frozen at exactly the iteration where the reference breaks, and the host -- which never
    blocks on the device -- stops enqueueing steps as soon as a lagged copy of the flag tells it so.
    `check_every=k > 0` restores a blocking host test every k steps; `check_every=0` never tests.

    `cache` (a dict owned by the caller, e.g. the compiled solver) turns the step into a CUDA graph that is captured on the
    first solve with this `cache_key` and replayed by every later one (see `_capture`); the operator must then read only
    tensors whose storage persists between solves.
    """
    n_it = int(min(max_iters, int(np.prod(b.shape))))
    key = ("cg", cache_key, tuple(b.shape), b.dtype, b.device, rtol, n_it)
    st = cache.get(key) if (cache is not None and check_every is None) else None
    if st is not None:                                           # replay: static state re-initialised in place
        x, r, p, gamma, tol2, poll, graph = st
        if x0 is None:
            x.zero_(); r.copy_(b)
        else:
            x.copy_(x0); r.copy_(ops.axpby(1.0, b, -1.0, A(x)))
        tol2.copy_((rtol * rtol) * ops.dot(b, b))
        gamma.copy_(ops.dot(r, r))
        p.copy_(r)
        poll.reset()
        for it in range(n_it):
            if poll.stopped():
                break
            graph.replay()
            poll.push()
            steps_enqueued[0] += 1
        return x.clone()
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    r = b.clone() if x0 is None else ops.axpby(1.0, b, -1.0, A(x))
    tol2 = None
    poll = None
    if check_every is None or check_every:
        bn2 = ops.dot(b, b)
        tol2 = (rtol * rtol) * bn2                               # B scalars (cold)
        if check_every is None:
            poll = _StopPoll(b.device)
    gamma = ops.dot(r, r)
    p = None
    graph = None

    def body():                                                  # one step of the device-gated loop, on static tensors
        q = A(p)
        pq = ops.dot(p, q)
        ops.cg_gate(gamma, tol2, pq, poll.done)
        gamma_new = ops.cg_update(x, r, p, q, gamma, pq)
        ops.cg_direction(p, r, gamma_new, gamma)
        gamma.copy_(gamma_new)

    for it in range(n_it):
        if check_every and it % check_every == 0 and bool(torch.all(gamma <= tol2)):
            if verbose:
                print("Converged at CG Iter %03d" % it)
            break
        if poll is not None and poll.stopped():
            break
        if graph is not None:
            graph.replay()
            poll.push()
            steps_enqueued[0] += 1
            continue
        if it == 0:
            p = r.clone()
        steps_enqueued[0] += 1
        q = A(p)
        pq = ops.dot(p, q)
        if poll is not None:
            ops.cg_gate(gamma, tol2, pq, poll.done)             # converged -> pq = inf -> the update below is a no-op
            poll.push()
        gamma_new = ops.cg_update(x, r, p, q, gamma, pq)        # x += a p ; r -= a q ; <r,r>
        ops.cg_direction(p, r, gamma_new, gamma)                # p = r + (g'/g) p
        gamma = gamma_new
        if it == 0 and cache is not None and poll is not None and n_it > 3:
            graph = _capture(body)
    if graph is not None:
        cache[key] = (x, r, p, gamma, tol2, poll, graph)
        return x.clone()
    return x


def pcg(A: Callable, b: torch.Tensor, x0: Optional[torch.Tensor] = None, rtol: float = 1e-6, max_iters: int = 100,
        verbose: bool = False, Minv: Optional[Callable] = None, check_every: Optional[int] = None, cache: Optional[dict] = None,
        cache_key=None):
    """Preconditioned CG with the reference's conventions (solver_cg.py:172-233): starts from ones,
    whole-tensor dot products, absolute inf-norm stop `max|r| < rtol` -- tested on the device like `cg`'s.

    With `Minv=None` it is algebraically CG started at ones with global dots: the same three fused
    kernels are used with batch=1 (the residual tracked here is b - A x = -r_ref).  `cache`: see `cg`."""
    key = ("pcg", cache_key, tuple(b.shape), b.dtype, b.device, rtol, max_iters)
    st = cache.get(key) if (cache is not None and check_every is None and Minv is None) else None
    if st is not None:
        x, r, p, gamma, rmax, poll, graph, _tol = st              # _tol: the gate's threshold, read by the captured step
        if x0 is None:
            x.fill_(1.0)
        else:
            x.copy_(x0)
        r.copy_(ops.axpby(1.0, b, -1.0, A(x)))
        gamma.copy_(ops.dot(r, r, per_sample=False))
        p.copy_(r)
        rmax.fill_(float("inf"))                                 # the reference tests from the second step on
        poll.reset()
        for it in range(max_iters):
            if poll.stopped():
                break
            graph.replay()
            poll.push()
            steps_enqueued[0] += 1
        return x.clone()
    x = torch.ones_like(b) if x0 is None else x0.clone()
    r = ops.axpby(1.0, b, -1.0, A(x))
    poll = _StopPoll(b.device) if check_every is None else None
    tol = torch.full((1,), rtol, device=b.device) if poll is not None else None
    if Minv is None:
        gamma = ops.dot(r, r, per_sample=False)
        p = r.clone()
        rmax = None
        graph = None

        def body():                                              # one step of the device-gated loop, on static tensors
            q = A(p)
            pq = ops.dot(p, q, per_sample=False)
            ops.cg_gate(rmax, tol, pq, poll.done, strict=True)
            gamma_new = ops.cg_update(x, r, p, q, gamma, pq, per_sample=False)
            ops.cg_direction(p, r, gamma_new, gamma, per_sample=False)
            gamma.copy_(gamma_new)
            rmax.copy_(ops.absmax(r))

        for it in range(max_iters):
            if poll is not None and poll.stopped():
                break
            if graph is not None:
                graph.replay()
                poll.push()
                steps_enqueued[0] += 1
                continue
            steps_enqueued[0] += 1
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
                if it == 0 and cache is not None and max_iters > 3:
                    graph = _capture(body)
            elif check_every and (it + 1) % check_every == 0 and float(ops.absmax(r)) < rtol:
                break
        if graph is not None:
            # every tensor the captured step reads must outlive this call -- the 4-byte threshold included: freed, its block
            # is handed to the next small allocation and the replayed gate compares max|r| with whatever lands there
            cache[key] = (x, r, p, gamma, rmax, poll, graph, tol)
            return x.clone()
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


def _build_solver(config: LinearSolveConfig, cache: Optional[dict] = None, cache_key=None):
    if config.solver_type not in SOLVERS:
        raise NotImplementedError(f"solver_type={config.solver_type!r}: only {sorted(SOLVERS)} are lowered "
                                  f"(minres/plss are out of scope, SURVEY §2 row 22)")
    kw = dict(config.solver_kwargs)
    if cache is not None:
        kw.update(cache=cache, cache_key=cache_key)
    return partial(SOLVERS[config.solver_type], rtol=config.rtol, max_iters=config.max_iters, verbose=config.verbose, **kw)


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


def linear_solve(A: Callable, b: torch.Tensor, config: LinearSolveConfig = LinearSolveConfig(), cache: Optional[dict] = None,
                 cache_key=None):
    """Solve A x = b matrix-free (linalg/custom.py:65-82); differentiable w.r.t. b by implicit differentiation.
    `cache` / `cache_key`: CUDA-graph replay of the CG step across solves (see `cg`); not used under autograd."""
    if config.use_analytic_grad and torch.is_grad_enabled() and b.requires_grad:
        return ImplicitSolve.apply(A, b, config)
    if torch.is_grad_enabled() and b.requires_grad:
        cache = None
    return _build_solver(config, cache, cache_key)(A, b)
