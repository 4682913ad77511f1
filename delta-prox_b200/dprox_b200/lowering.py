"""compile()-time lowering: objective (psi_fns, omega_fns) -> execution engine.

`analyze()` is pure host logic (no tensors, testable without a GPU).  It classifies every term with the
reference's diagonalisability rules (least_squares.__init__, proxfn/sum_square.py:104-107; is_gram_diag
recursion) and picks one of

  tier 'native'  : every term folds into the plan's normal form -> one `dpx_iters` call runs all
                   iterations (or the staged form when a prox is an external Python callable);
  tier 'generic' : something does not fold (BlackBox, mosaic(conv(x)), conv as a psi linop, LADMM with
                   grad terms, ext_sum_squares) -> the algorithm is composed node by node from the
                   stand-alone kernels, with the closed-form `dpx_xsolve` when the Gram matrix is still
                   diagonal(isable) and fused-kernel CG otherwise.

Engines hold the native plan and the hoisted constants; they are rebuilt when the input shape, the
device or a Placeholder value changes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import _cabi as cabi
from . import ops
from .linalg import LinearSolveConfig, linear_solve
from .linop import Lowered, Placeholder, Variable, evaluate, evaluate_adjoint
from .tensors import as_bchw

ALGO_IDS = {"admm": cabi.ALGO_ADMM, "ladmm": cabi.ALGO_LADMM, "admm_vxu": cabi.ALGO_ADMM_VXU, "hqs": cabi.ALGO_HQS,
            "pgd": cabi.ALGO_PGD, "pc": -1,       # PockChambolle has no fused plan: always composed by the generic engine
            "custom_admm": -1}                    # contrib.CustomADMM (CS-MRI, prox first, complex iterates): generic engine


@dataclass
class TermSpec:
    fn: object
    low: Optional[Lowered]
    kind: str                       # lowered kind or 'generic'
    prox_kind: int = cabi.PROX_EXTERNAL

    @property
    def scale(self):
        return self.low.scale if self.low is not None else 1.0


@dataclass
class PlanSpec:
    method: str
    tier: str                       # 'native' | 'generic'
    xupdate: str                    # 'freq' | 'spatial' | 'scalar' | 'cg' | 'ext' | 'none'
    psi: List[TermSpec] = field(default_factory=list)
    quad: List[TermSpec] = field(default_factory=list)
    has_external: bool = False
    reason: str = ""
    diagonalizable: bool = False
    freq_diagonalizable: bool = False


def _term(fn) -> TermSpec:
    low = fn.linop.lower() if fn.linop is not None else None
    if low is not None and low.kind == "const":
        low = None
    kind = low.kind if low is not None else "generic"
    prox_kind = fn.native_kind if fn.is_native() else cabi.PROX_EXTERNAL
    return TermSpec(fn, low, kind, prox_kind)


def analyze(psi_fns, omega_fns, method: str, try_diagonalize=True, try_freq_diagonalize=True) -> PlanSpec:
    from .proxfn import ext_sum_squares
    if method not in ALGO_IDS:
        raise ValueError(f"method {method!r} is not lowered by this backend (supported: {sorted(ALGO_IDS)})")
    psi = [_term(f) for f in psi_fns]
    quad = [_term(f) for f in omega_fns]
    terms = psi + quad
    vars_ = {v.uuid for t in terms if t.fn.linop is not None for v in t.fn.linop.variables}
    if len(vars_) > 1:
        raise NotImplementedError("objectives over more than one Variable are not supported "
                                  "(neither are they by the reference's iteration, SURVEY App. A-18)")
    kinds = [t.kind for t in terms]
    freq_ok = all(k in ("identity", "spectral", "grad", "grad2d") for k in kinds)
    for t in psi:
        if t.prox_kind == cabi.PROX_ISO_TV and t.kind != "grad2d":
            raise ValueError("iso_tv needs the stacked gradient operator grad2d(x) as its linop")
    spat_ok = all(k in ("identity", "mask") for k in kinds)
    # generic nodes may still be diagonalisable through the plugin protocol (BlackBox with diag=...)
    diagonalizable = (spat_ok or all(t.fn.linop.is_gram_diag(False) for t in terms)) and try_diagonalize
    freq_diagonalizable = (freq_ok or all(t.fn.linop.is_gram_diag(True) for t in terms)) and try_diagonalize \
        and try_freq_diagonalize
    spec = PlanSpec(method, "generic", "cg", psi, quad, any(t.prox_kind == cabi.PROX_EXTERNAL for t in psi), "",
                    diagonalizable, freq_diagonalizable)

    if any(isinstance(t.fn, ext_sum_squares) for t in quad):
        spec.xupdate, spec.reason = "ext", "ext_sum_squares data term brings its own x-update"
        return spec
    if method == "pgd":
        q = quad[0]
        if q.kind == "generic" or not try_diagonalize:
            spec.xupdate, spec.reason = "none", "PGD gradient through a generic operator"
            return spec
        spec.xupdate = "freq" if q.kind in ("identity", "spectral", "grad") else "spatial"
        if spec.xupdate == "freq" and not try_freq_diagonalize:
            spec.xupdate, spec.reason = "none", "frequency diagonalisation disabled"
            return spec
        # The reference applies the prox to x itself and ignores the psi LINOP except for its constant (pgd.py:39-43): only
        # an unscaled identity linop has that meaning in the fused plan; anything else takes the reference's literal step.
        if psi[0].kind != "identity" or float(psi[0].scale) != 1.0 or (q.kind == "spectral" and q.low.otf_fn is None):
            spec.xupdate, spec.reason = "none", "PGD with a non-identity psi linop: prox applied to x as in the reference"
            return spec
        spec.tier = "native"
        return spec

    # x-update closed form (reference: freq branch wins whenever freq_diagonalizable, sum_square.py:118-121,150)
    if freq_ok and spat_ok and try_diagonalize and (try_freq_diagonalize or True):
        spec.xupdate = "scalar"
    elif freq_ok and try_diagonalize and try_freq_diagonalize:
        spec.xupdate = "freq"
    elif spat_ok and try_diagonalize:
        spec.xupdate = "spatial"
    else:
        spec.xupdate, spec.reason = "cg", "Gram matrix is not diagonal(isable): CG fallback"
        return spec

    if method in ("pc", "custom_admm"):
        spec.reason = f"{method} is composed node by node (closed-form x-update where diagonalisable)"
        return spec
    if method == "admm_vxu" and spec.has_external:
        spec.reason = "ADMM_vxu with an external prox: composed node by node (closed-form x-update)"
        return spec
    # can the psi side be fused?
    for t in psi:
        if t.kind == "identity":
            # LinearizedADMM forms b_i = x - K^T(Kx - v + u) = (1 - s^2) x + s (v - u): ADMM's v - u only for |s| = 1
            # (algo/admm.py:82-90), so scaled identity terms keep the reference's literal update
            if method == "ladmm" and abs(float(t.scale)) != 1.0:
                spec.reason = "LADMM with a scaled identity psi linop: composed node by node"
                return spec
            continue
        if t.kind in ("grad", "grad2d") and spec.xupdate == "freq" and method in ("admm", "hqs") \
                and (t.kind == "grad" or t.prox_kind != cabi.PROX_EXTERNAL):
            continue
        spec.reason = f"psi linop of kind {t.kind!r} under {method}: composed node by node"
        return spec
    spec.tier = "native"
    return spec


# ------------------------------------------------------------------------------------------------
#  helpers shared by the engines
# ------------------------------------------------------------------------------------------------

def _sched(t: torch.Tensor, B: int, device) -> tuple:
    """schedule tensor [T] or [B,T] -> (contiguous device fp32 tensor, stride)"""
    t = torch.as_tensor(t, dtype=torch.float32).to(device)
    if t.ndim == 0:
        t = t.reshape(1)
    if t.ndim == 1:
        return t.contiguous(), 0
    if t.ndim == 2:
        if t.shape[0] != B:
            raise ValueError(f"per-sample schedule has {t.shape[0]} rows but the batch is {B}")
        return t.contiguous(), t.shape[1]
    raise ValueError(f"schedule must be [T] or [B,T], got {tuple(t.shape)}")


def _placeholder_versions(fns):
    """(constants, parameters): versions of the Placeholders that enter the objective as additive constants / explicit
    `b` (a change only moves the right-hand side K^T b and psi offsets) and of those that parametrise an operator
    (`conv_doe` PSF, `mul_elementwise` weight: a change also moves the diagonals)."""
    consts, params = [], []
    for fn in fns:
        stack = [fn.linop] if fn.linop is not None else []
        while stack:
            n = stack.pop()
            if isinstance(n, Placeholder):
                consts.append((id(n), n.version))
            for attr in ("_psf", "_w"):
                q = getattr(n, attr, None)
                if isinstance(q, Placeholder):
                    params.append((id(q), q.version))
            stack += list(n.input_nodes)
        b = getattr(fn, "_b", None)
        if isinstance(b, Placeholder):
            consts.append((id(b), b.version))
    return tuple(consts), tuple(params)


def _quad_rhs(quad: List[TermSpec], like: torch.Tensor) -> Optional[torch.Tensor]:
    """sum_q A_q^T b_q  — iteration-invariant, hoisted (the reference recomputes it every iteration,
    sum_square.py:127-132)."""
    acc = None
    for t in quad:
        fn = t.fn
        b = torch.as_tensor(fn.offset).to(like.device)
        if b.is_complex():                                   # complex measurements (k-space) go straight to the plugin's adjoint
            b = b.to(torch.complex64)
        else:
            b = b.to(torch.float32)
            if b.shape != like.shape:
                b = torch.broadcast_to(b, like.shape).contiguous()
        grads = {}
        evaluate_adjoint(fn.linop, b, grads)
        g = next(iter(grads.values())) if grads else None
        if g is None:
            continue
        acc = g if acc is None else ops.axpby(1.0, acc, 1.0, g)
    return acc


def _gram_diag(terms: List[TermSpec], shape4, device, freq: bool, include_identity: bool = True):
    """sum_t scale^2 * gram_t as one real array ([1|B,C,H,Wc] for freq, [1|B,C,H,W] for spatial), or None."""
    B, Cc, H, W = shape4
    acc = None
    const = 0.0
    for t in terms:
        s2 = float(t.scale) ** 2
        if t.kind == "identity":
            if include_identity:
                const += s2
            continue
        if freq:
            g = t.low.gram_fn(shape4)
            g = torch.as_tensor(g).to(device, torch.float32)
        else:
            g = t.low.mask(shape4, device)
            g = torch.as_tensor(g).to(device, torch.float32)
        if g.ndim == 3:
            g = g.unsqueeze(0)
        acc = g * s2 if acc is None else acc + g * s2        # cold path: a handful of constant arrays
    if acc is None:
        if const == 0.0 and not include_identity:
            return None
        last = W // 2 + 1 if freq else W
        acc = torch.zeros(1, Cc, H, last, device=device, dtype=torch.float32)
    if const != 0.0:
        acc = acc + const
    if acc.shape[1] != Cc:
        acc = acc.expand(acc.shape[0], Cc, *acc.shape[2:])
    return acc.contiguous()


class _EngineBase:
    def __init__(self, spec: PlanSpec, shape, device, eps=1e-7):
        self.spec, self.device, self.eps = spec, torch.device(device), eps
        self.shape = tuple(shape)                               # user-facing shape of x
        self.shape4 = tuple(as_bchw(torch.empty(self.shape, device="meta")).shape)
        self.key = None
        self.const_key = None

    def _v(self, t):
        return t.reshape(self.shape4)


# ------------------------------------------------------------------------------------------------
#  Native engine: one plan, everything in the C-ABI hot loop
# ------------------------------------------------------------------------------------------------

class NativeEngine(_EngineBase):
    def __init__(self, spec: PlanSpec, x0: torch.Tensor, fft_backend=cabi.FFT_AUTO, eps=1e-7):
        super().__init__(spec, x0.shape, x0.device, eps)
        B, Cc, H, W = self.shape4
        d = cabi.ProblemDesc()
        d.abi_version = cabi.ABI_VERSION
        d.batch, d.channels, d.height, d.width = B, Cc, H, W
        d.algo = ALGO_IDS[spec.method]
        d.xupdate = cabi.X_FREQ_DIAG if spec.xupdate == "freq" else cabi.X_SPATIAL_DIAG
        d.eps_delta = 1 if spec.xupdate == "scalar" else 0
        d.n_psi = len(spec.psi)
        if d.n_psi > cabi.MAX_PSI:
            raise NotImplementedError(f"at most {cabi.MAX_PSI} prox terms are supported")
        for i, t in enumerate(spec.psi):
            p = d.psi[i]
            p.prox = t.prox_kind
            p.linop = {"identity": cabi.LINOP_IDENTITY, "grad": cabi.LINOP_GRAD_H, "grad2d": cabi.LINOP_GRAD_HW}[t.kind]
            if t.kind == "grad":
                p.linop = cabi.LINOP_GRAD_H if t.low.axis == 0 else cabi.LINOP_GRAD_W
            p.scale, p.alpha, p.beta = float(t.scale), float(t.fn.alpha), float(t.fn.beta)
            p.box_lo, p.box_hi = t.fn.box
        d.eps, d.fft_backend = eps, fft_backend
        self.plan = cabi.NativePlan(d, self.device)
        self._set_constants(x0)

    # constants -----------------------------------------------------------------------------------
    def _set_constants(self, x0):
        spec, dev = self.spec, self.device
        like = torch.zeros(self.shape, device=dev, dtype=torch.float32)
        for t in spec.psi + spec.quad:                          # offsets are evaluated against the variable's shape
            for v in t.fn.linop.variables:
                if v._value is None or tuple(v._value.shape) != self.shape or v._value.device != dev:
                    v._value = like
        ktb = _quad_rhs(spec.quad, like)
        ktb4 = None if ktb is None else self._v(ktb).contiguous()
        s = cabi.stream_ptr(dev)
        lib = cabi.lib()
        with torch.cuda.device(dev):
            if spec.xupdate == "freq":
                dq = _gram_diag(spec.quad, self.shape4, dev, True)
                dpsi = _gram_diag([t for t in spec.psi if t.kind != "identity"], self.shape4, dev, True, include_identity=False)
                if dpsi is not None and dpsi.shape[0] != 1:
                    raise NotImplementedError("per-sample psi OTFs")
                # a grey PSF gives every channel the same diagonal: then any two planes can share a complex transform, which
                # is what lets a single RGB image (odd plane count) use the plane-pair engine.  One comparison per constant set.
                B, Cc = self.shape4[0], self.shape4[1]
                shared = dq.shape[0] == 1 and Cc > 1 and B % 2 == 1 and bool((dq[:, 1:] == dq[:, :1]).all())
                cabi.check(lib.dpx_plan_set_hint(self.plan.handle, cabi.HINT_CHANNEL_SHARED_DIAG, int(shared)), "dpx_plan_set_hint")
                cabi.check(lib.dpx_plan_set_freq_constants(self.plan.handle, cabi.ptr(ktb4), cabi.ptr(dq), dq.shape[0],
                                                           cabi.ptr(dpsi), s), "dpx_plan_set_freq_constants")
            else:
                dq = _gram_diag(spec.quad, self.shape4, dev, False)
                cabi.check(lib.dpx_plan_set_spatial_constants(self.plan.handle, cabi.ptr(ktb4), cabi.ptr(dq), dq.shape[0], s),
                           "dpx_plan_set_spatial_constants")
            for i, t in enumerate(spec.psi):
                c = t.low.const_tensor(like)
                if c is not None:
                    off = ops.axpby(-1.0, self._v(c).contiguous())
                    cabi.check(lib.dpx_plan_set_psi_offset(self.plan.handle, i, cabi.ptr(off), s), "dpx_plan_set_psi_offset")
        self._keep = (ktb, dq)

    def _rhs_in_fourier_domain(self) -> bool:
        """sum_squares(s * conv(x) - b) as the only quadratic term: F(K^T b) = s conj(OTF) F(b) without leaving the
        Fourier domain (dpx_plan_set_rhs_spectral) -- two FFTs less per new batch of measurements."""
        spec = self.spec
        if spec.xupdate != "freq" or len(spec.quad) != 1 or spec.quad[0].kind != "spectral" or spec.quad[0].low.otf_fn is None:
            return False
        t = spec.quad[0]
        b = torch.as_tensor(t.fn.offset)
        if b.is_complex() or not b.is_cuda:
            return False
        b = cabi.require_cuda_f32(b.to(self.device, torch.float32).expand(self.shape), "b")
        otf = torch.as_tensor(t.low.otf_fn(self.shape4))
        key = (id(t.low.otf_fn.__self__) if hasattr(t.low.otf_fn, "__self__") else id(t.low.otf_fn), otf.data_ptr())
        if getattr(self, "_otf_dev_key", None) != key:
            self._otf_dev = otf.to(self.device, torch.complex64).contiguous()
            self._otf_dev_key = key
        otf = self._otf_dev
        with torch.cuda.device(self.device):
            cabi.check(cabi.lib().dpx_plan_set_rhs_spectral(self.plan.handle, cabi.ptr(self._v(b).contiguous()), C.c_void_p(otf.data_ptr()),
                                                            otf.shape[0] if otf.ndim == 4 else 1, float(t.scale),
                                                            cabi.stream_ptr(self.device)), "dpx_plan_set_rhs_spectral")
        return True

    def update_rhs(self, x0):
        """A Placeholder-fed measurement changed: re-hoist K^T b (and psi offsets) only; diagonals and plan stay."""
        spec, dev = self.spec, self.device
        if not any(t.low is not None and t.low.const for t in spec.psi) and self._rhs_in_fourier_domain():
            return
        like = torch.zeros(self.shape, device=dev, dtype=torch.float32)
        for t in spec.psi + spec.quad:
            for v in t.fn.linop.variables:
                if v._value is None or tuple(v._value.shape) != self.shape or v._value.device != dev:
                    v._value = like
        ktb = _quad_rhs(spec.quad, like)
        ktb4 = None if ktb is None else self._v(ktb).contiguous()
        s, lib = cabi.stream_ptr(dev), cabi.lib()
        with torch.cuda.device(dev):
            cabi.check(lib.dpx_plan_set_rhs(self.plan.handle, cabi.ptr(ktb4), s), "dpx_plan_set_rhs")
            for i, t in enumerate(spec.psi):
                c = t.low.const_tensor(like)
                if c is not None:
                    off = ops.axpby(-1.0, self._v(c).contiguous())
                    cabi.check(lib.dpx_plan_set_psi_offset(self.plan.handle, i, cabi.ptr(off), s), "dpx_plan_set_psi_offset")
        self._keep = (ktb, self._keep[1])

    # state ---------------------------------------------------------------------------------------
    def initialize(self, x0):
        x = cabi.require_cuda_f32(x0.to(self.device, torch.float32), "x0").clone()
        m, method = len(self.spec.psi), self.spec.method
        if method == "pgd":
            return [x]
        def term_like(t):                               # a stacked-gradient term carries [B,2C,H,W] state
            if t.kind != "grad2d":
                return torch.empty_like(x)
            B, Cc, H, W = self.shape4
            return torch.empty(B, 2 * Cc, H, W, device=x.device, dtype=x.dtype)
        v = [term_like(t) for t in self.spec.psi]
        u = [term_like(t) for t in self.spec.psi] if method != "hqs" else None
        with torch.cuda.device(self.device):
            cabi.check(cabi.lib().dpx_init_state(self.plan.handle, cabi.ptr(x), cabi.ptr_array(v), cabi.ptr_array(u),
                                                 cabi.stream_ptr(self.device)), "dpx_init_state")
        return (x, v) if method == "hqs" else (x, v, u)

    def _unpack(self, state):
        if self.spec.method == "pgd":
            return state[0], None, None
        if self.spec.method == "hqs":
            return state[0], state[1], None
        return state

    # hot loop ------------------------------------------------------------------------------------
    def run(self, state, rhos, lams: Dict, it0: int, n_iters: int, resid: Optional[torch.Tensor] = None):
        """Iterations it0..it0+n_iters-1 of the schedules, in place on `state`."""
        x, v, u = self._unpack(state)
        B = self.shape4[0]
        rho_t, rho_s = _sched(rhos, B, self.device)
        lam_t, lam_s = [], []
        for t in self.spec.psi:
            lt, ls = _sched(lams[t.fn], B, self.device)
            lam_t.append(lt)
            lam_s.append(ls)
        lib, s = cabi.lib(), cabi.stream_ptr(self.device)
        h = self.plan.handle
        with torch.cuda.device(self.device):
            if not self.spec.has_external:
                cabi.check(lib.dpx_iters(h, cabi.ptr(x), cabi.ptr_array(v), cabi.ptr_array(u), cabi.ptr(rho_t), rho_s,
                                         cabi.ptr_array(lam_t), cabi.int_array(lam_s), it0, n_iters, cabi.ptr(resid), s),
                           "dpx_iters")
                return state
            for it in range(it0, it0 + n_iters):
                self._staged_iteration(x, v, u, rho_t, rho_s, lam_t, lam_s, it)
        return state

    def _staged_iteration(self, x, v, u, rho_t, rho_s, lam_t, lam_s, it):
        lib, s, h = cabi.lib(), cabi.stream_ptr(self.device), self.plan.handle
        method = self.spec.method
        if method == "admm_vxu":
            raise NotImplementedError("ADMM_vxu with an external prox")
        B = self.shape4[0]
        if method == "pgd":
            scratch = [torch.empty_like(x)]
            cabi.check(lib.dpx_stage_xupdate(h, cabi.ptr(x), cabi.ptr_array(scratch), None, cabi.ptr(rho_t), rho_s, it, s),
                       "dpx_stage_xupdate")
            t = self.spec.psi[0]
            lam = lam_t[0][..., it] if lam_s[0] else lam_t[0][it]
            t.fn.step = it
            x.copy_(t.fn.prox(scratch[0].reshape(self.shape), lam).reshape(x.shape))
            return
        cabi.check(lib.dpx_stage_xupdate(h, cabi.ptr(x), cabi.ptr_array(v), cabi.ptr_array(u), cabi.ptr(rho_t), rho_s, it, s),
                   "dpx_stage_xupdate")
        cabi.check(lib.dpx_stage_prox(h, cabi.ptr(x), cabi.ptr_array(v), cabi.ptr_array(u), cabi.ptr_array(lam_t),
                                      cabi.int_array(lam_s), it, s), "dpx_stage_prox")
        for i, t in enumerate(self.spec.psi):
            if t.prox_kind != cabi.PROX_EXTERNAL:
                continue
            lam = lam_t[i][..., it] if lam_s[i] else lam_t[i][it]
            t.fn.step = it
            w = v[i]                                             # stage_prox left w = K x + u here
            v_new = t.fn.prox(w.reshape(self.shape), lam)
            v_new = cabi.require_cuda_f32(v_new.reshape(w.shape), "external prox output")
            cabi.check(lib.dpx_stage_dual_external(h, i, cabi.ptr(w), cabi.ptr(v_new), cabi.ptr(v[i]),
                                                   cabi.ptr(u[i]) if u is not None else None, s), "dpx_stage_dual_external")

    def solve_host(self, x0_host: torch.Tensor, rho_host: torch.Tensor, lam_host: torch.Tensor, n_iters: int, out_host=None):
        """End-to-end entry with HOST buffers (pinned preferred): H2D, init, all iterations, D2H, sync."""
        if out_host is None:
            out_host = torch.empty_like(x0_host)
        with torch.cuda.device(self.device):
            cabi.check(cabi.lib().dpx_solve_host(self.plan.handle, C.c_void_p(x0_host.data_ptr()), C.c_void_p(out_host.data_ptr()),
                                                 C.c_void_p(rho_host.data_ptr()),
                                                 C.c_void_p(lam_host.data_ptr()) if lam_host is not None else None,
                                                 n_iters, cabi.stream_ptr(self.device)), "dpx_solve_host")
        return out_host


# ------------------------------------------------------------------------------------------------
#  Generic engine: the algorithm composed from stand-alone kernels + plugin callables
# ------------------------------------------------------------------------------------------------

class GenericEngine(_EngineBase):
    """Also the differentiable engine: every kernel it composes has an autograd Function (`dprox_b200.autograd`), so
    under grad mode `step()` records a tape whose backward is native as well (SURVEY §8b autograd contract)."""

    def __init__(self, spec: PlanSpec, x0: torch.Tensor, linear_solve_config: LinearSolveConfig = LinearSolveConfig(), eps=1e-7):
        super().__init__(spec, x0.shape, x0.device, eps)
        self.cfg = linear_solve_config
        self.plan = None
        self.ktb = None
        self._step, self._cg_rho, self._cg_cache = 0, None, {}
        self._has_blackbox = False
        for t in spec.psi + spec.quad:
            stack = [t.fn.linop] if t.fn.linop is not None else []
            while stack:
                n = stack.pop()
                self._has_blackbox = self._has_blackbox or type(n).__name__ == "BlackBox"
                stack += list(n.input_nodes)
        if spec.xupdate in ("freq", "spatial", "scalar"):
            B, Cc, H, W = self.shape4
            d = cabi.ProblemDesc()
            d.abi_version = cabi.ABI_VERSION
            d.batch, d.channels, d.height, d.width = B, Cc, H, W
            d.algo, d.n_psi, d.eps, d.fft_backend = cabi.ALGO_ADMM, 0, self.eps, cabi.FFT_AUTO     # fused x-update for 2^k sizes
            if spec.xupdate == "freq":
                d.xupdate, d.eps_delta = cabi.X_FREQ_DIAG, 0
            else:
                # spatial closed form (ktb + rho t) / (dq + rho dpsi + eps) (sum_square.py:142-148, 154); identity-only
                # objectives take the reference's Fourier branch with a constant diagonal = the same quotient + eps * delta_0
                d.xupdate, d.eps_delta = cabi.X_SPATIAL_DIAG, int(spec.xupdate == "scalar")
            self.plan = cabi.NativePlan(d, self.device)
        self.set_constants()

    def set_constants(self):
        """(Re-)hoist K^T b and the diagonals into the existing plan: called when a Placeholder changed and, under grad
        mode, before every solve so that `ktb` carries a fresh tape to the measurements."""
        spec, dev = self.spec, self.device
        like = torch.zeros(self.shape, device=dev, dtype=torch.float32)
        for t in spec.psi + spec.quad:
            for v in t.fn.linop.variables:
                if v._value is None or tuple(v._value.shape) != self.shape or v._value.device != dev:
                    v._value = like
        self.ktb = _quad_rhs(spec.quad, like) if spec.xupdate not in ("ext",) else None
        if self.plan is None:
            return
        ktb4 = None if self.ktb is None else self._v(self.ktb.detach()).contiguous()
        freq = spec.xupdate == "freq"
        dq = _gram_diag(spec.quad, self.shape4, dev, freq)
        dpsi = _gram_diag(spec.psi, self.shape4, dev, freq)        # identity contributions folded in (plan has wid = 0)
        if dpsi is not None and dpsi.shape[0] != 1:
            raise NotImplementedError("per-sample psi diagonals")
        lib, s = cabi.lib(), cabi.stream_ptr(dev)
        with torch.cuda.device(dev):
            if freq:
                cabi.check(lib.dpx_plan_set_freq_constants(self.plan.handle, cabi.ptr(ktb4), cabi.ptr(dq), dq.shape[0],
                                                           cabi.ptr(dpsi), s), "dpx_plan_set_freq_constants")
            else:
                cabi.check(lib.dpx_plan_set_spatial_constants(self.plan.handle, cabi.ptr(ktb4), cabi.ptr(dq), dq.shape[0], s),
                           "dpx_plan_set_spatial_constants")
                cabi.check(lib.dpx_plan_set_spatial_psi_diag(self.plan.handle, cabi.ptr(dpsi), s), "dpx_plan_set_spatial_psi_diag")
        self.plan.const_version += 1                               # autograd.XSolve refuses a backward across a constant change
        self._keep = (ktb4, dq, dpsi)

    # linop application through the tree -------------------------------------------------------------
    @staticmethod
    def _K(fn, x, zero_const=False):
        var = fn.linop.variables[0]
        return evaluate(fn.linop, {var: x}, zero_const)

    @staticmethod
    def _Kt(fn, y):
        grads = {}
        evaluate_adjoint(fn.linop, y, grads)
        return next(iter(grads.values()))

    def _rho4(self, rho):
        rho = torch.as_tensor(rho, dtype=torch.float32, device=self.device)
        return rho.reshape(-1)

    def solve_x(self, b: List[torch.Tensor], rho, like):
        """least_squares.solve (sum_square.py:115-197)."""
        spec = self.spec
        rho = self._rho4(rho)
        if spec.xupdate == "ext":
            fn = spec.quad[0].fn
            return fn.solve(b, rho)
        t = None
        for term, bi in zip(spec.psi, b):
            g = self._Kt(term.fn, bi)
            t = g if t is None else ops.axpby(1.0, t, 1.0, g)
        if self.plan is not None:
            B = self.shape4[0]
            if t is None:
                t = torch.zeros_like(like)
            rt, rs = _sched(rho.reshape(-1, 1) if rho.numel() > 1 else rho.reshape(1), B, self.device)
            return ops.xsolve(self.plan, t.contiguous(), rt, rs, self.ktb)
        # CG on the normal equations (solve_cg, sum_square.py:158-197)
        rhs = self.ktb if t is None else (ops.lincomb(t, rho) if self.ktb is None else ops.lincomb(self.ktb, None, t, rho))
        # The CG step is captured as a CUDA graph on the first solve and replayed afterwards (linalg._capture): everything the
        # normal operator reads must therefore live in storage that persists between solves -- rho is copied into a buffer the
        # engine owns.  A BlackBox receives the iteration index as a Python argument (linop/blackbox.py:38-41), which a graph
        # would freeze: trees with BlackBox nodes get one graph per iteration index.
        cache = None if torch.is_grad_enabled() else self._cg_cache      # replay only outside autograd (backward re-runs KtK later)
        if cache is not None:
            if self._cg_rho is None or self._cg_rho.shape != rho.shape:
                self._cg_rho = rho.clone()
                cache.clear()
            else:
                self._cg_rho.copy_(rho)
            rho = self._cg_rho
        step_key = self._step if self._has_blackbox else None

        def KtK(x):
            out = None
            for q in spec.quad:
                g = self._Kt(q.fn, self._K(q.fn, x, zero_const=True))
                out = g if out is None else ops.axpby(1.0, out, 1.0, g)
            acc = None
            for p_ in spec.psi:
                g = self._Kt(p_.fn, self._K(p_.fn, x, zero_const=True))
                acc = g if acc is None else ops.axpby(1.0, acc, 1.0, g)
            if acc is not None:
                out = ops.lincomb(acc, rho) if out is None else ops.lincomb(out, None, acc, rho)
            return out

        return linear_solve(KtK, rhs, self.cfg, cache=cache, cache_key=step_key)

    # state + one iteration ----------------------------------------------------------------------------
    def initialize(self, x0):
        if x0.is_complex():                                  # complex iterates (CS-MRI): glue kernels work on (re, im) pairs
            if not x0.is_cuda:
                raise RuntimeError("dprox_b200 computes on CUDA devices only (no CPU fallback)")
            x = x0.to(self.device, torch.complex64).contiguous().clone()
        else:
            x = cabi.require_cuda_f32(x0.to(self.device, torch.float32), "x0").clone()
        if self.spec.method == "pgd":
            return [x]
        v = [self._K(t.fn, x) for t in self.spec.psi]
        if self.spec.method == "hqs":
            return x, v
        if self.spec.method == "pc":
            return x, v, x.clone()
        return x, v, [torch.zeros_like(e) for e in v]

    def step(self, state, rho, lam: Dict, it: int):
        spec, m = self.spec, self.spec.method
        for t in spec.psi + spec.quad:
            _set_step(t.fn, it)
        self._step = it
        if m in ("admm", "ladmm"):
            x, v, u = state
            if m == "admm":
                b = [ops.axpby(1.0, v[i], -1.0, u[i]) for i in range(len(spec.psi))]
            else:   # LinearizedADMM: b_i = x - K_i^T (K_i x - v_i + u_i)   (admm.py:82-90)
                b = []
                for i, t in enumerate(spec.psi):
                    tmp = ops.lincomb(self._K(t.fn, x, zero_const=True), None, v[i], _const(-1.0, x), u[i], None)
                    b.append(ops.axpby(1.0, x, -1.0, self._Kt(t.fn, tmp)))
            x = self.solve_x(b, rho, x)
            for i, t in enumerate(spec.psi):
                Kx = self._K(t.fn, x)
                w = ops.axpby(1.0, Kx, 1.0, u[i])
                v[i] = t.fn.prox(w, lam[t.fn])
                u[i] = ops.axpby(1.0, w, -1.0, v[i])
            return x, v, u
        if m == "hqs":
            x, z = state
            x = self.solve_x(z, rho, x)
            for i, t in enumerate(spec.psi):
                z[i] = t.fn.prox(self._K(t.fn, x), lam[t.fn])
            return x, z
        if m == "admm_vxu":
            z, xs, u = state
            for i, t in enumerate(spec.psi):
                xs[i] = t.fn.prox(ops.axpby(1.0, self._K(t.fn, z), -1.0, u[i]), lam[t.fn])
            b = [ops.axpby(1.0, xs[i], 1.0, u[i]) for i in range(len(spec.psi))]
            z = self.solve_x(b, rho, z)
            for i in range(len(spec.psi)):
                u[i] = ops.lincomb(u[i], None, xs[i], None, z, _const(-1.0, z))
            return z, xs, u
        if m == "custom_admm":  # CustomADMM._iter (contrib/csmri.py:157-171): prox first, complex z / u, real x
            x, zs, u = state
            z = zs[0]
            xs = [t.fn.prox(ops.axpby(1.0, z, -1.0, u[i]), lam[t.fn]) for i, t in enumerate(spec.psi)]
            lift = ops.to_complex if z.is_complex() else (lambda e: e)
            b = [ops.axpby(1.0, lift(xs[i]), 1.0, u[i]) for i in range(len(spec.psi))]
            z = self.solve_x(b, rho, z)
            for i in range(len(spec.psi)):
                u[i] = ops.axpby(1.0, b[i], -1.0, z)           # u + x - z
            return xs[0], [z], u
        if m == "pc":          # PockChambolle._iter (algo/pc.py:13-36)
            x, z, xbar = state
            for i, t in enumerate(spec.psi):
                r = self._rho4(lam[t.fn])
                zi = ops.lincomb(z[i], None, self._K(t.fn, xbar), r)
                z[i] = ops.lincomb(zi, None, t.fn.prox(zi, r), -r)
            xn = [ops.axpby(1.0, x, -1.0, self._Kt(t.fn, z[i])) for i, t in enumerate(spec.psi)]
            if spec.quad:
                x_next = self.solve_x(xn, rho, x)
            else:
                x_next = xn[0]
                for e in xn[1:]:
                    x_next = ops.axpby(1.0, x_next, 1.0, e)
            return x_next, z, ops.axpby(2.0, x_next, -1.0, x)
        if m == "pgd":
            x = state[0]
            g = spec.quad[0].fn.grad(x)
            rho = self._rho4(rho)
            v = ops.lincomb(x, None, g, -rho)
            fn = spec.psi[0].fn
            return [fn.prox(v, lam[fn])]
        raise ValueError(m)


def _const(val, like):
    return torch.full((1,), float(val), device=like.device, dtype=torch.float32)


def _set_step(fn, step):
    fn.step = step
    stack = [fn.linop] if fn.linop is not None else []
    while stack:
        n = stack.pop()
        n.step = step
        stack += list(n.input_nodes)
