"""Problem / compile() / Algorithm.solve — the reference's L2/L3 surface over the
B200 engines of `dprox_b200.lowering`.

What stays the same for callers: class names, constructor and `solve/iters/iter/initialize/pack/unpack`
signatures, the partition rules, defaults (rho=1.0, lam=0.02 for every psi fn, max_iter=24), the
`callback(iter=, state=, rho=, lam=)` hook, `Variable.value` holding the result after a solve.

What is different underneath: the loop body is not Python.  When no callback/progress bar is requested
and every prox is native, `solve()` makes ONE C-ABI call that runs all iterations on the GPU with the
schedules resident on the device.  State tensors are updated in place.
"""
from __future__ import annotations

import copy as _copy
import warnings
from functools import partial
from typing import Callable, Dict, Iterable, List, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from . import _cabi as cabi
from .linalg import LinearSolveConfig
from .linop import CompGraph, Variable, vstack
from .lowering import GenericEngine, NativeEngine, PlanSpec, _placeholder_versions, analyze
from .proxfn import ProxFn, ext_sum_squares, sum_squares
from .tensors import to_torch_tensor


def _isscalar(x):
    return np.isscalar(x) or (isinstance(x, torch.Tensor) and x.ndim == 0)


def _to_tensor(x, batch=False):
    if isinstance(x, dict):
        return {k: _to_tensor(v, batch) for k, v in x.items()}
    return to_torch_tensor(x, batch)


class ResidualStop:
    """Opt-in residual stopping rule (absent from the reference's imaging loop; semantics follow
    lp/solvers.py:324-336): stop when  |r| <= abstol*sqrt(n) + reltol*max(|Kx|,|v|)  and  |s| <= abstol*sqrt(n) +
    reltol*|v|, with r = Kx - v and s = rho*(v - v_prev), summed over all psi terms, all samples and -- through
    `group` -- all ranks.

    The check is ASYNCHRONOUS: every `every` iterations the per-sample sums are reduced on the device, the
    5-float vector is all-reduced on a SIDE stream (NCCL) and copied to pinned host memory behind an event; the loop keeps
    enqueueing iterations and consumes the decision `lag` checks later (default: one check late), by which time the event
    has long fired, so neither the device nor the host ever waits for the collective.  The solve therefore runs
    `lag * every` iterations past the first satisfied check.  `lag=0` gives the blocking variant."""

    def __init__(self, abstol=1e-4, reltol=1e-3, every=4, group=None, lag=1):
        self.abstol, self.reltol, self.every, self.group, self.lag = abstol, reltol, max(1, int(every)), group, max(0, int(lag))
        self.history = []
        self._side = None
        self._pending = []

    def _decide(self, tot) -> bool:
        r, s, kx, v = [max(float(t), 0.0) ** 0.5 for t in tot[:4]]
        n_elems = float(tot[4])
        self.history.append((r, s))
        eps_abs = self.abstol * (n_elems ** 0.5)
        return r <= eps_abs + self.reltol * max(kx, v) and s <= eps_abs + self.reltol * v

    def converged(self, sums: torch.Tensor, n_elems: int) -> bool:
        """Blocking form.  sums = [sum|r|^2, sum|s|^2, sum|Kx|^2, sum|v|^2] (device or CPU tensor), already local-summed."""
        import torch.distributed as dist
        tot = torch.cat([sums.double(), torch.tensor([float(n_elems)], dtype=torch.float64, device=sums.device)])
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
        return self._decide(tot.cpu())

    # -- asynchronous form ---------------------------------------------------------------------------------------------
    def submit(self, sums: torch.Tensor, n_elems: int):
        """Queue the all-reduce + host copy of one check behind the work already enqueued on the current stream."""
        import torch.distributed as dist
        dev = sums.device
        tot = torch.cat([sums.double(), torch.tensor([float(n_elems)], dtype=torch.float64, device=dev)])
        host = torch.empty(5, dtype=torch.float64).pin_memory() if dev.type == "cuda" else None
        if dev.type != "cuda":                                        # CPU tensors (gloo tests): nothing to overlap with
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
            self._pending.append((None, tot, tot))
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)     # NCCL enqueues on the side stream
            host.copy_(tot, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        tot.record_stream(self._side)
        self._pending.append((done, host, tot))

    def poll(self, flush: bool = False) -> bool:
        """Consume the checks that are at least `lag` submissions old (all of them with flush=True)."""
        hit = False
        while self._pending and (flush or len(self._pending) > self.lag):
            done, host, _keep = self._pending.pop(0)
            if done is not None:
                done.synchronize()                                    # fired long ago unless lag == 0
            hit = self._decide(host) or hit
        return hit


def _flat_tensors(*objs):
    for o in objs:
        if isinstance(o, torch.Tensor):
            yield o
        elif isinstance(o, dict):
            yield from _flat_tensors(*o.values())
        elif isinstance(o, (list, tuple)):
            yield from _flat_tensors(*o)


class Algorithm(nn.Module):
    """Base class of the proximal solvers (dprox/algo/base.py:58-275)."""

    method = None

    @classmethod
    def partition(cls, prox_fns: List[ProxFn]):
        raise NotImplementedError

    @classmethod
    def create(cls, *args, **kwargs):
        return cls(*args, **kwargs)

    def __init__(self, psi_fns: List[ProxFn], omega_fns: List[ProxFn], try_diagonalize=True, try_freq_diagonalize=True,
                 linear_solve_config: LinearSolveConfig = LinearSolveConfig(), fft_backend: int = cabi.FFT_AUTO):
        super().__init__()
        self.psi_fns = nn.ModuleList(psi_fns)
        self.omega_fns = nn.ModuleList(omega_fns)
        self.K = CompGraph(vstack([fn.linop for fn in psi_fns]))
        self.Kall = CompGraph(vstack([fn.linop for fn in list(psi_fns) + list(omega_fns)]))
        self.try_diagonalize, self.try_freq_diagonalize = try_diagonalize, try_freq_diagonalize
        self.linear_solve_config = linear_solve_config
        self.fft_backend = fft_backend
        self._dev_anchor = nn.Parameter(torch.tensor(0.0), requires_grad=False)
        self.spec: PlanSpec = analyze(list(psi_fns), list(omega_fns), self.method, try_diagonalize, try_freq_diagonalize)
        self._engine = None

    def __deepcopy__(self, memo):
        """Per-iteration copies of a solver (UnrolledSolver(share=False), unroll.py:10-27): everything but the native plans,
        which every copy rebuilds on first use."""
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_engine", "_engine_d") else _copy.deepcopy(v, memo)
        return new

    # reference attribute: `solver.least_square.{diagonalizable,freq_diagonalizable}`
    @property
    def least_square(self):
        return self.spec

    @property
    def device(self):
        return self._dev_anchor.device

    def _wants_grad(self, *objs) -> bool:
        """Autograd contract: `solve/iters/iter` are differentiable w.r.t. the state, the measurements, `rhos`,
        `lams` and trainable parameters whenever grad mode is on and one of them requires grad (unrolled training,
        specialization/unroll.py:42-58).  Such calls run on the differentiable engine, whose forward AND backward are
        native kernels (dprox_b200.autograd); everything else takes the fused forward-only loop."""
        if not torch.is_grad_enabled():
            return False
        if any(t.requires_grad for t in _flat_tensors(*objs)) or any(p.requires_grad for p in self.parameters()):
            return True
        for fn in list(self.psi_fns) + list(self.omega_fns):            # Placeholder-fed measurements carrying a tape
            stack = [fn.linop] if fn.linop is not None else []
            while stack:
                n = stack.pop()
                val = getattr(n, "_value", None)
                if isinstance(val, torch.Tensor) and val.requires_grad:
                    return True
                stack += list(n.input_nodes)
            b = getattr(fn, "_b", None)
            val = getattr(b, "_value", None) if b is not None else None
            if isinstance(val, torch.Tensor) and val.requires_grad:
                return True
        return False

    # -- engine management ---------------------------------------------------------------------------
    def engine(self, x0: torch.Tensor, diff: bool = False):
        if diff:
            return self._diff_engine(x0)
        key = (tuple(x0.shape), str(x0.device))
        ckey = _placeholder_versions(list(self.psi_fns) + list(self.omega_fns))
        if self._engine is not None and self._engine.key == key and self._engine.const_key != ckey:
            # only Placeholder values changed: keep the plan.  A new batch of measurements (additive constants) only
            # moves K^T b; a new operator parameter (PSF / weight) also moves the diagonals.
            if isinstance(self._engine, NativeEngine):
                if self._engine.const_key[1] == ckey[1]:
                    self._engine.update_rhs(x0)
                else:
                    self._engine._set_constants(x0)
            else:
                with torch.no_grad():
                    self._engine.set_constants()
            self._engine.const_key = ckey
        if self._engine is None or self._engine.key != key or self._engine.const_key != ckey:
            if not x0.is_cuda:
                raise RuntimeError(f"dprox_b200 computes on CUDA devices only; the solver lives on {x0.device}. "
                                   f"Use compile(..., device='cuda'). There is no CPU fallback.")
            if self.spec.tier == "native":
                eng = NativeEngine(self.spec, x0, fft_backend=self.fft_backend)
            else:
                with torch.no_grad():
                    eng = GenericEngine(self.spec, x0, self.linear_solve_config)
            eng.key, eng.const_key = key, ckey
            self._engine = eng
        return self._engine

    def _diff_engine(self, x0: torch.Tensor):
        """The differentiable engine: the algorithm composed kernel by kernel, each with a native backward."""
        if not x0.is_cuda:
            raise RuntimeError(f"dprox_b200 computes on CUDA devices only; the solver lives on {x0.device}.")
        key = (tuple(x0.shape), str(x0.device))
        ckey = _placeholder_versions(list(self.psi_fns) + list(self.omega_fns))
        eng = getattr(self, "_engine_d", None)
        if eng is None or eng.key != key:
            eng = GenericEngine(self.spec, x0, self.linear_solve_config)
            eng.key = key
            self._engine_d = eng
        elif eng.const_key != ckey:
            eng.set_constants()
        eng.const_key = ckey
        return eng

    # -- argument handling (base.py:20-45, 205-218) ---------------------------------------------------
    def defaults(self, x0=None, rhos=None, lams=None, max_iter=24):
        if rhos is None:
            rhos = 1.0
        if lams is None:
            lams = 0.02
        if _isscalar(rhos):
            rhos = _to_tensor([float(rhos)] * max_iter)
        if _isscalar(lams):
            lams = {fn: _to_tensor([float(lams)] * max_iter) for fn in self.psi_fns}
        elif not isinstance(lams, dict):                               # one schedule shared by every psi fn
            lams = {fn: lams for fn in self.psi_fns}
        lams = {k: _to_tensor([float(v)] * max_iter) if _isscalar(v) else _to_tensor(v) for k, v in lams.items()}
        missing = [str(fn) for fn in self.psi_fns if fn not in lams]
        if missing:
            raise KeyError(f"`lams` given as a dict must contain every psi fn (missing: {missing}); "
                           f"the reference indexes lam[fn] for all of them (algo/admm.py:56)")
        return x0, _to_tensor(rhos), lams, max_iter

    def solve(self, x0=None, rhos=None, lams=None, max_iter: int = 24, pbar: bool = False, callback: Callable = None,
              return_full_states: bool = False, stop: Optional[ResidualStop] = None, **kwargs) -> torch.Tensor:
        """Algorithm.solve (base.py:85-126): fixed `max_iter` iterations unless an opt-in `stop` rule is given."""
        x0 = _to_tensor(x0, batch=True)
        x0, rhos, lams, max_iter = self.defaults(x0, rhos, lams, max_iter)
        x0 = x0.to(self.device, torch.complex64 if x0.is_complex() else torch.float32)
        diff = self._wants_grad(x0, rhos, lams)
        if diff:
            self._diff_engine(x0).set_constants()                    # fresh tape from the measurements to K^T b
        state = self.initialize(x0, _diff=diff, **kwargs)
        state = self.iters(state, rhos, lams, max_iter, pbar, callback=callback, stop=stop, _diff=diff)
        return state if return_full_states else state[0]

    def initialize(self, x0, _diff=None, **kwargs):
        x0 = torch.as_tensor(x0)                                     # already batched by solve() (base.py:20-33)
        x0 = x0.to(self.device, torch.complex64 if x0.is_complex() else torch.float32)
        diff = self._wants_grad(x0) if _diff is None else _diff
        return self.engine(x0, diff).initialize(x0)

    def iters(self, state, rhos, lams, max_iter, pbar=False, callback=None, stop: Optional[ResidualStop] = None, _diff=None):
        """Algorithm.iters (base.py:128-156).

        State ownership differs from the reference on the native engine: x, v, u are updated IN PLACE (the reference returns
        fresh tensors every iteration, admm.py:49-59), and x / v are only materialised by the last iteration of a native call.
        The `state` a `callback` receives therefore aliases buffers that the next iteration overwrites -- clone what must
        outlive the call."""
        diff = self._wants_grad(state, rhos, lams) if _diff is None else _diff
        eng = self.engine(state[0], diff)
        if _isscalar(lams) or not isinstance(lams, dict):
            lams = {fn: lams for fn in self.psi_fns}
        dev = state[0].device
        rhos = torch.as_tensor(rhos, dtype=torch.float32).to(dev)         # schedules live on the device
        lams = {k: torch.as_tensor(v, dtype=torch.float32).to(dev) for k, v in lams.items()}
        per_iter = callback is not None or pbar or isinstance(eng, GenericEngine)
        self.iterations_run = max_iter
        if stop is not None and isinstance(eng, NativeEngine) and not eng.spec.has_external and not per_iter:
            state = self._iters_with_stop(eng, state, rhos, lams, max_iter, stop)
        elif not per_iter:
            state = eng.run(state, rhos, lams, 0, max_iter)                 # ONE native call for the whole loop
        else:
            if stop is not None:
                warnings.warn("dprox_b200: the residual stop rule only applies to the fused native loop (no callback / progress "
                              "bar / external prox / node-by-node engine); running the fixed max_iter iterations", stacklevel=2)
            it_range = range(max_iter)
            if pbar:
                from tqdm import tqdm
                it_range = tqdm(it_range)
            for it in it_range:
                rho = rhos[..., it]
                lam = {k: v[..., it] for k, v in lams.items()}
                self._notify_all_op_current_step(it)
                if isinstance(eng, NativeEngine):
                    state = eng.run(state, rhos, lams, it, 1)
                else:
                    state = eng.step(state, rho.to(self.device), {k: v.to(self.device) for k, v in lam.items()}, it)
                if callback is not None:
                    callback(iter=it, state=state, rho=rho, lam=lam)
        self.Kall.update_vars([state[0]])
        return state

    def _iters_with_stop(self, eng, state, rhos, lams, max_iter, stop: ResidualStop):
        """Residual stopping rule: the first `every - 1` iterations of each block run in the fused loop (which keeps v only
        implicitly, so it has no v - v_prev to measure); the block's last iteration runs through the kernel that also
        accumulates {|r|^2, |s|^2, |Kx|^2, |v|^2} per sample with warp-shuffle reductions.  Those sums are reduced over the
        samples on the device (`dpx_resid_reduce`) and handed to the rule's asynchronous check (side-stream all-reduce,
        decision consumed one check later): the host never waits for the check of the block it has just enqueued."""
        B = eng.shape4[0]
        n_elems = state[0].numel() * max(1, len(self.psi_fns))
        dev = state[0].device
        it = 0
        stop._pending.clear()
        while it < max_iter:
            n = min(stop.every, max_iter - it)
            if n > 1:
                state = eng.run(state, rhos, lams, it, n - 1)
            resid = torch.empty(1, B, 4, device=dev, dtype=torch.float32)
            state = eng.run(state, rhos, lams, it + n - 1, 1, resid=resid)
            sums = torch.empty(1, 4, device=dev, dtype=torch.float32)
            with torch.cuda.device(dev):
                cabi.check(cabi.lib().dpx_resid_reduce(cabi.ptr(resid), cabi.ptr(sums), 1, B, cabi.stream_ptr(dev)), "dpx_resid_reduce")
            it += n
            stop.submit(sums[0], n_elems)
            if stop.poll():
                break
        else:
            stop.poll(flush=True)
        stop._pending.clear()
        self.iterations_run = it
        return state

    def iter(self, state, rho, lam):
        """One iteration with explicit (rho, lam) values (base.py:174-178); used by unrolled / DEQ callers."""
        eng = self.engine(state[0], self._wants_grad(state, rho, lam))
        rho = torch.as_tensor(rho, dtype=torch.float32)
        lam = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in lam.items()}
        if isinstance(eng, NativeEngine):
            rhos = rho.reshape(-1, 1) if rho.ndim == 1 else rho.reshape(1)
            lams = {k: (v.reshape(-1, 1) if v.ndim == 1 else v.reshape(1)) for k, v in lam.items()}
            state = eng.run(state, rhos, lams, 0, 1)
        else:
            state = eng.step(state, rho.to(self.device), {k: v.to(self.device) for k, v in lam.items()}, self._step_hint)
        self.Kall.update_vars([state[0]])
        return state

    _step_hint = 0

    def _iter(self, state, rho, lam):
        return self.iter(state, rho, lam)

    def _notify_all_op_current_step(self, step):
        self._step_hint = step
        for fn in list(self.psi_fns) + list(self.omega_fns):
            fn.step = step
            stack = [fn.linop]
            while stack:
                n = stack.pop()
                n.step = step
                stack += list(n.input_nodes)

    # -- helpers used by RL / DEQ wrappers (base.py:224-275) -------------------------------------------
    def pack(self, state):
        flat = []
        for s in state:
            flat += s if isinstance(s, list) else [s]
        return torch.cat(flat, dim=1)

    def unpack(self, tensor):
        vars_ = list(torch.split(tensor, tensor.shape[1] // self.state_dim, dim=1))
        out, start = [], 0
        for d in self.state_split:
            if d == 1:
                out.append(vars_[start].contiguous())
                start += 1
            else:
                out.append([t.contiguous() for t in vars_[start:start + d[0]]])
                start += d[0]
        return out

    @property
    def state_dim(self):
        return sum(s[0] if isinstance(s, list) else s for s in self.state_split)

    @property
    def nparams(self):
        return len(self.psi_fns) + 1

    @property
    def state_split(self):
        raise NotImplementedError


class ADMM(Algorithm):
    """algo/admm.py:25-76."""
    method = "admm"

    @classmethod
    def partition(cls, prox_fns: List[ProxFn]):
        omega, took_ext = [], False
        for fn in prox_fns:
            if not took_ext and isinstance(fn, ext_sum_squares):
                omega.append(fn)
                took_ext = True
            elif type(fn) == sum_squares:
                omega.append(fn)
        psi = [fn for fn in prox_fns if not any(fn is o for o in omega)]
        return psi, omega

    @property
    def state_split(self):
        return [1, [len(self.psi_fns)], [len(self.psi_fns)]]


class LinearizedADMM(ADMM):
    """algo/admm.py:79-100.  Identity psi linops: bit-for-bit the ADMM kernels (the reference's b_i collapses to
    v_i - u_i there); other linops: the reference's exact (self-inconsistent, App. A-6) update via the generic engine."""
    method = "ladmm"


class ADMM_vxu(ADMM):
    """algo/admm.py:103-120 (per-sample semantics; the reference's batch-index slip for a single psi fn,
    App. A-17, is not reproduced)."""
    method = "admm_vxu"


class HQS(ADMM):
    """algo/hqs.py."""
    method = "hqs"

    @property
    def state_split(self):
        return [1, [len(self.psi_fns)]]


class PockChambolle(ADMM):
    """algo/pc.py — primal-dual iteration; state (x, [z_i], xbar)."""
    method = "pc"

    @property
    def state_split(self):
        return [1, [len(self.psi_fns)], 1]


class ProximalGradientDescent(Algorithm):
    """algo/pgd.py."""
    method = "pgd"

    @classmethod
    def partition(cls, prox_fns: List[ProxFn]):
        if len(prox_fns) != 2:
            raise ValueError("Proximal gradient descent only supports two proximal functions for now.")
        omega = [fn for fn in prox_fns if hasattr(fn, "grad")]
        psi = [fn for fn in prox_fns if not any(fn is o for o in omega)]
        if len(omega) == 0:
            raise ValueError("Proximal gradient descent requires at least one proximal function is differentiable.")
        return psi, omega

    @property
    def state_split(self):
        return [1]


SOLVERS = {"admm": ADMM, "admm_vxu": ADMM_vxu, "ladmm": LinearizedADMM, "hqs": HQS, "pc": PockChambolle,
           "pgd": ProximalGradientDescent}


def compile(prox_fns: List[ProxFn], method: str = "admm", device: Union[str, torch.device] = "cuda", **kwargs):  # noqa: A001
    """Compile an objective into a solver (algo/primitives.py:40-67).  `device` must be a CUDA device for `solve()`."""
    if method not in SOLVERS:
        raise ValueError(f"unknown or unsupported method {method!r} (supported: {sorted(SOLVERS)}; "
                         f"'pc' is listed as next in SURVEY §8f)")
    if isinstance(prox_fns, ProxFn):
        prox_fns = [prox_fns]
    algorithm = SOLVERS[method]
    device = torch.device(device) if isinstance(device, str) else device
    psi_fns, omega_fns = algorithm.partition(prox_fns)
    solver = algorithm.create(psi_fns, omega_fns, **kwargs)
    return solver.to(device)


class UnrolledSolver(nn.Module):
    """algo/specialization/unroll.py:20-58: one solver copy per unrolled iteration (`share=False`: deep copies, so trainable
    denoisers / operator parameters are per-iteration) and, with `learned_params=True`, rho / lam schedules that are
    nn.Parameters initialised to ones.  Like the reference it keys the per-iteration lam by `psi_fns[0]` of the copy, i.e. it
    serves objectives with a single prox term."""

    def __init__(self, solver: Algorithm, max_iter: int, share: bool = False, learned_params: bool = False):
        super().__init__()
        if share is False:
            self.solvers = nn.ModuleList([_copy.deepcopy(solver) for _ in range(max_iter)])
        else:
            self.solver = solver
            self.solvers = [solver for _ in range(max_iter)]
        self.max_iter, self.share, self.learned_params = max_iter, share, learned_params
        if learned_params:
            self.rhos = nn.Parameter(torch.ones(max_iter))
            self.lams = {}
            for fn in solver.psi_fns:
                lam = nn.Parameter(torch.ones(max_iter))
                setattr(self, str(fn), lam)
                self.lams[fn] = lam

    def solve(self, x0=None, rhos=None, lams=None, max_iter=None):
        first = self.solvers[0]
        x0 = _to_tensor(x0, batch=True)
        dev = first.device
        if self.learned_params:
            rhos, lams = self.rhos, self.lams
        else:
            rhos = _to_tensor(rhos).to(dev)
            if not isinstance(lams, dict):
                lams = {fn: lams for fn in first.psi_fns}
            lams = {k: _to_tensor(v).to(dev) for k, v in lams.items()}
        x0 = x0.to(dev, torch.complex64 if x0.is_complex() else torch.float32)
        max_iter = self.max_iter if max_iter is None else max_iter
        diff = first._wants_grad(x0, rhos, lams) or any(s._wants_grad() for s in self.solvers[1:max_iter])
        if diff:
            for s in {id(s): s for s in self.solvers[:max_iter]}.values():
                s._diff_engine(x0).set_constants()                   # fresh tape from the measurements to K^T b
        state = first.initialize(x0, _diff=diff)
        for i in range(max_iter):
            rho = rhos[..., i:i + 1]
            lam = {self.solvers[i].psi_fns[0]: v[..., i:i + 1] for v in lams.values()}
            state = self.solvers[i].iters(state, rho, lam, 1, False, _diff=diff)
        return state[0]


def build_unrolled_solver(solver: Algorithm, share: bool = True, **kwargs):
    """unroll.py:14-18."""
    if share is True:
        solver.solve = partial(solver.solve, **kwargs)
        return solver
    return UnrolledSolver(solver, share=share, **kwargs)


def specialize(solver: Algorithm, method: str = "unroll", device="cuda", **kwargs):
    """algo/primitives.py:70-95; only the 'unroll' specialisation is part of this backend (DEQ / RL wrap training loops)."""
    if method != "unroll":
        raise NotImplementedError("only specialize(..., method='unroll') is available in this backend")
    solver = build_unrolled_solver(solver, **kwargs)
    device = torch.device(device) if isinstance(device, str) else device
    return solver.to(device)


class Problem:
    """algo/problem.py:13-58 (the LP branch is out of scope)."""

    def __init__(self, prox_fns, constraints=[], absorb=True, merge=True, try_diagonalize=True, try_freq_diagonalize=True,
                 linear_solve_config=LinearSolveConfig()):
        if isinstance(prox_fns, ProxFn):
            prox_fns = [prox_fns]
        self.prox_fns = prox_fns
        self.solver_args = dict(try_diagonalize=try_diagonalize, try_freq_diagonalize=try_freq_diagonalize,
                                linear_solve_config=linear_solve_config)

    @property
    def objective(self):
        return self.prox_fns

    def solve(self, method="admm", device="cuda", **kwargs):
        solver = compile(self.prox_fns, method=method, device=device, **self.solver_args)
        return solver.solve(**kwargs)


def log_descent(upper, lower, iter=24, sigma=0.255 / 255, w=1.0, lam=0.23, sqrt=False):
    """rho/sigma schedules of DPIR (algo/tune/dpir.py:13-39)."""
    s_log = np.logspace(np.log10(upper), np.log10(lower), iter).astype(np.float32)
    s_lin = np.linspace(upper, lower, iter).astype(np.float32)
    sigmas = (s_log * w + s_lin * (1 - w)) / 255.0
    rhos = [lam * (sigma ** 2) / (s ** 2) for s in sigmas]
    if not sqrt:
        sigmas = sigmas ** 2
    return torch.tensor(np.asarray(rhos)).float(), torch.tensor(np.asarray(sigmas)).float()
