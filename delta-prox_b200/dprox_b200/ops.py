"""Thin, typed wrappers over the stand-alone C-ABI kernels (no torch arithmetic here).

Everything takes/returns CUDA fp32 torch tensors; torch is only the allocator and stream owner."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _cabi as cabi
from . import autograd as ag
from .tensors import as_bchw

_fft_plans = {}


def _s(t):
    return cabi.stream_ptr(t.device)


def _new_like(x):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


def axpby(a: float, x: torch.Tensor, b: float = 0.0, y: Optional[torch.Tensor] = None, out=None) -> torch.Tensor:
    """out = a*x + b*y  (scale / sum / subtraction glue; linop/scale.py, linop/sum.py).
    complex64 operands (k-space data of plugin operators) are processed as interleaved float pairs."""
    if isinstance(x, torch.Tensor) and x.is_complex():
        yr = None if y is None else torch.view_as_real(y.to(torch.complex64).expand_as(x).contiguous())
        return torch.view_as_complex(axpby(a, torch.view_as_real(x.to(torch.complex64).contiguous()), b, yr))
    if out is None and ag.needs_grad(x, y):
        return ag.Axpby.apply(x, y, float(a), float(b))
    x = cabi.require_cuda_f32(x, "x")
    if y is not None:
        y = cabi.require_cuda_f32(y, "y")
        if y.shape != x.shape:
            y = y.expand_as(x).contiguous()
    out = _new_like(x) if out is None else out
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_axpby(cabi.ptr(out), float(a), cabi.ptr(x), float(b), cabi.ptr(y), x.numel(), _s(x)),
                   "dpx_axpby")
    return out


def lincomb(x, a=None, y=None, b=None, z=None, c=None, out=None) -> torch.Tensor:
    """out = a*x + b*y + c*z with per-sample ([B]) or shared ([1]) DEVICE coefficients (None = 1)."""
    if out is None and ag.needs_grad(x, a, y, b, z, c):
        return ag.Lincomb.apply(x, a, y, b, z, c)
    x = cabi.require_cuda_f32(x, "x")
    B = x.shape[0]
    per = x.numel() // B
    coeffs = [t for t in (a, b, c) if t is not None]
    cps = 1 if any(t.numel() > 1 for t in coeffs) else 0
    fix = lambda t: None if t is None else cabi.require_cuda_f32(t.reshape(-1).expand(B) if (cps and t.numel() == 1) else t.reshape(-1), "coeff")
    a, b, c = fix(a), fix(b), fix(c)
    y = None if y is None else cabi.require_cuda_f32(y, "y")
    z = None if z is None else cabi.require_cuda_f32(z, "z")
    out = _new_like(x) if out is None else out
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_lincomb(cabi.ptr(out), cabi.ptr(a), cabi.ptr(x), cabi.ptr(b), cabi.ptr(y), cabi.ptr(c),
                                          cabi.ptr(z), cps, B, per, _s(x)), "dpx_lincomb")
    return out


def mul(x: torch.Tensor, w: torch.Tensor, out=None) -> torch.Tensor:
    """out = w * x with w of batch 1 or B (mosaic / mul_elementwise)."""
    if out is None and ag.needs_grad(x):
        return ag.Mul.apply(x, w.detach())
    x = cabi.require_cuda_f32(x, "x")
    w = cabi.require_cuda_f32(w.to(x.device), "w")
    B = x.shape[0]
    per = x.numel() // B
    if w.numel() == x.numel():
        wb = B
    elif w.numel() == per:
        wb = 1
    else:
        w = w.expand(1, *x.shape[1:]).contiguous() if w.shape[0] == 1 else w.expand_as(x).contiguous()
        wb = w.shape[0]
    out = _new_like(x) if out is None else out
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_mul_apply(cabi.ptr(out), cabi.ptr(x), cabi.ptr(w), wb, B, per, _s(x)), "dpx_mul_apply")
    return out


def grad(x: torch.Tensor, axis: int, adjoint: bool = False, scale: float = 1.0) -> torch.Tensor:
    """Circular forward difference along H (axis=0) or W (axis=1), or its adjoint (linop/grad.py:8-23)."""
    if ag.needs_grad(x):
        return ag.Grad.apply(x, int(axis), bool(adjoint), float(scale))
    x = cabi.require_cuda_f32(x, "x")
    x4 = as_bchw(x)
    out = _new_like(x)
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_grad_apply(cabi.ptr(x), cabi.ptr(out), x4.shape[0] * x4.shape[1], x4.shape[2], x4.shape[3],
                                             int(axis), int(adjoint), float(scale), _s(x)), "dpx_grad_apply")
    return out


def pad2d(x: torch.Tensor, out_hw, top: int, left: int) -> torch.Tensor:
    """out[..., y, x] = in[..., y - top, x - left] inside the input, 0 elsewhere: zero padding (offsets >= 0) or cropping
    (offsets < 0) of the last two axes -- the data movement of the `circular=False` convolutions (linop/conv.py:100-121)."""
    if ag.needs_grad(x):
        return ag.Pad2d.apply(x, tuple(out_hw), int(top), int(left))
    x = cabi.require_cuda_f32(x, "x")
    hi, wi = x.shape[-2:]
    ho, wo = out_hw
    out = torch.empty(*x.shape[:-2], ho, wo, device=x.device, dtype=torch.float32)
    planes = x.numel() // (hi * wi)
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_pad2d(cabi.ptr(x), cabi.ptr(out), planes, hi, wi, ho, wo, int(top), int(left), _s(x)), "dpx_pad2d")
    return out


def augment(x: torch.Tensor, mode: int) -> torch.Tensor:
    """One of the 8 flips / rotations of the x8 test-time augmentation (pnp/denoisers/composite.py:30-47) of [B,C,H,W]."""
    mode = int(mode) % 8
    if mode == 0:
        return x
    if ag.needs_grad(x):
        return ag.Augment.apply(x, mode)
    x = cabi.require_cuda_f32(x, "x")
    B, Cc, H, W = x.shape
    tr = mode in (1, 3, 5, 7)
    out = torch.empty(B, Cc, W if tr else H, H if tr else W, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_augment(cabi.ptr(x), cabi.ptr(out), B * Cc, H, W, mode, _s(x)), "dpx_augment")
    return out


AUGMENT_INVERSE = {0: 0, 1: 1, 2: 2, 3: 5, 4: 4, 5: 3, 6: 6, 7: 7}     # the transform that undoes mode m


def prox(kind: int, v: torch.Tensor, lam: torch.Tensor, alpha=1.0, beta=1.0, lo=0.0, hi=0.0, offset=None, out=None):
    """ProxFn.prox with the wrapper chain for a native `_prox` body (proxfn/base.py:55-64)."""
    if out is None and ag.needs_grad(v, lam, offset):
        return ag.Prox.apply(v, lam, offset, int(kind), float(alpha), float(beta), float(lo), float(hi))
    v = cabi.require_cuda_f32(v, "v")
    B = v.shape[0] if v.ndim > 0 else 1
    lam = cabi.require_cuda_f32(lam.to(v.device, torch.float32).reshape(-1), "lam")
    if lam.numel() not in (1, B):
        raise ValueError(f"lam must have 1 or B={B} entries, got {lam.numel()}")
    if offset is not None:
        offset = cabi.require_cuda_f32(offset.to(v.device).expand_as(v), "offset")
    out = _new_like(v) if out is None else out
    with torch.cuda.device(v.device):
        cabi.check(cabi.lib().dpx_prox_apply(int(kind), cabi.ptr(v), cabi.ptr(lam), int(lam.numel() > 1), float(alpha), float(beta),
                                             float(lo), float(hi), cabi.ptr(offset), cabi.ptr(out), B, v.numel() // B, _s(v)),
                   "dpx_prox_apply")
    return out


def fft_plan(shape, device) -> cabi.NativePlan:
    """A constants-free FREQ_DIAG plan used as the FFT context of stand-alone spectral filters."""
    key = (tuple(shape), str(device))
    if key not in _fft_plans:
        d = cabi.ProblemDesc()
        d.abi_version = cabi.ABI_VERSION
        d.batch, d.channels, d.height, d.width = shape
        d.algo, d.xupdate, d.n_psi, d.eps, d.fft_backend = cabi.ALGO_ADMM, cabi.X_FREQ_DIAG, 0, 1e-7, cabi.FFT_CUFFT
        _fft_plans[key] = cabi.NativePlan(d, torch.device(device))
    return _fft_plans[key]


def spectral_filter(x: torch.Tensor, otf: torch.Tensor, conj: bool = False, plan: Optional[cabi.NativePlan] = None):
    """y = Re F^-1(otf * F x) (or conj(otf)): conv.forward / conv.adjoint (linop/conv.py:31-41).
    `otf` is a complex64 half spectrum [1|B, C, H, W/2+1]."""
    if ag.needs_grad(x):
        x4 = as_bchw(x)
        plan = plan or fft_plan(tuple(x4.shape), x.device)
        return ag.SpectralFilter.apply(x, otf.detach(), bool(conj), plan)
    x = cabi.require_cuda_f32(x, "x")
    x4 = as_bchw(x)
    B, Cc, H, W = x4.shape
    if otf.dtype != torch.complex64:
        otf = otf.to(torch.complex64)
    otf = otf.to(x.device).contiguous()
    if tuple(otf.shape[-3:]) != (Cc, H, W // 2 + 1):
        raise ValueError(f"OTF shape {tuple(otf.shape)} does not match input {tuple(x4.shape)}")
    ob = otf.shape[0] if otf.ndim == 4 else 1
    plan = plan or fft_plan((B, Cc, H, W), x.device)
    out = _new_like(x)
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_spectral_filter(plan.handle, cabi.ptr(x), C.c_void_p(otf.data_ptr()), ob, int(conj),
                                                  cabi.ptr(out), _s(x)), "dpx_spectral_filter")
    return out


def xsolve(plan: cabi.NativePlan, t: torch.Tensor, rho: torch.Tensor, rho_stride: int, ktb: Optional[torch.Tensor] = None):
    """least_squares.solve closed form through a plan whose constants are set (dpx_xsolve); differentiable w.r.t. the
    psi part of the right-hand side `t`, `rho` and — through `ktb`, which the plan already holds as F(ktb) — the
    measurements (autograd.XSolve)."""
    t = cabi.require_cuda_f32(t, "t")
    if ag.needs_grad(t, rho, ktb):
        return ag.XSolve.apply(plan, t, rho, int(rho_stride), ktb)
    x = torch.empty_like(t)
    with torch.cuda.device(t.device):
        cabi.check(cabi.lib().dpx_xsolve(plan.handle, cabi.ptr(t), cabi.ptr(rho), int(rho_stride), 0, cabi.ptr(x), _s(t)),
                   "dpx_xsolve")
    return x


def dot(x: torch.Tensor, y: torch.Tensor, per_sample: bool = True) -> torch.Tensor:
    """bdot (linalg/solve/solver_cg.py:7-22): per-sample <x_b, y_b> -> device [B] (or [1] when not per_sample)."""
    x, y = cabi.require_cuda_f32(x, "x"), cabi.require_cuda_f32(y, "y")
    B = x.shape[0] if per_sample else 1
    out = torch.empty(B, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_cg_dot(cabi.ptr(x), cabi.ptr(y), cabi.ptr(out), B, x.numel() // B, _s(x)), "dpx_cg_dot")
    return out


def absmax(x: torch.Tensor, per_sample: bool = False) -> torch.Tensor:
    x = cabi.require_cuda_f32(x, "x")
    B = x.shape[0] if per_sample else 1
    out = torch.empty(B, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_absmax(cabi.ptr(x), cabi.ptr(out), B, x.numel() // B, _s(x)), "dpx_absmax")
    return out


def cg_update(x, r, p, q, gamma, pq, per_sample=True) -> torch.Tensor:
    """alpha = gamma/pq; x += alpha p; r -= alpha q; returns the new <r,r> (device [B])."""
    B = x.shape[0] if per_sample else 1
    gn = torch.empty(B, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_cg_update(cabi.ptr(x), cabi.ptr(r), cabi.ptr(p), cabi.ptr(q), cabi.ptr(gamma), cabi.ptr(pq),
                                            cabi.ptr(gn), B, x.numel() // B, _s(x)), "dpx_cg_update")
    return gn


def cg_gate(val, tol, pq, done, strict=False):
    """device-side stop test: done |= all(val <= tol) (strict: <); while done, pq := +inf so that the next cg_update is a no-op"""
    with torch.cuda.device(val.device):
        cabi.check(cabi.lib().dpx_cg_gate(cabi.ptr(val), cabi.ptr(tol), int(tol.numel()), int(strict), cabi.ptr(pq),
                                          C.c_void_p(done.data_ptr()), int(val.numel()), _s(val)), "dpx_cg_gate")


def cg_direction(p, r, gamma_new, gamma_old, per_sample=True):
    """p = r + (gamma_new/gamma_old) p, in place."""
    B = p.shape[0] if per_sample else 1
    with torch.cuda.device(p.device):
        cabi.check(cabi.lib().dpx_cg_direction(cabi.ptr(p), cabi.ptr(r), cabi.ptr(gamma_new), cabi.ptr(gamma_old), B,
                                               p.numel() // B, _s(p)), "dpx_cg_direction")
    return p


def to_complex(x: torch.Tensor) -> torch.Tensor:
    """real fp32 -> complex64 with zero imaginary part (native copy kernel); complex input is returned as is."""
    if x.is_complex():
        return x
    x = cabi.require_cuda_f32(x, "x")
    out = torch.empty(x.shape, device=x.device, dtype=torch.complex64)
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().dpx_real_to_complex(cabi.ptr(x), C.c_void_p(out.data_ptr()), x.numel(), _s(x)), "dpx_real_to_complex")
    return out


def real_part(z: torch.Tensor) -> torch.Tensor:
    """real part of a complex64 CUDA tensor as a contiguous fp32 tensor (pnp/prior.py:79 `v = v.real`)."""
    if not z.is_complex():
        return z
    if not z.is_cuda:
        raise RuntimeError("dprox_b200: complex tensor lives on the CPU; this backend only computes on CUDA devices")
    z = z.to(torch.complex64).contiguous()
    out = torch.empty(z.shape, device=z.device, dtype=torch.float32)
    with torch.cuda.device(z.device):
        cabi.check(cabi.lib().dpx_complex_real(C.c_void_p(z.data_ptr()), cabi.ptr(out), z.numel(), _s(z)), "dpx_complex_real")
    return out


def csmri_prox(v: torch.Tensor, y: torch.Tensor, mask: torch.Tensor, rho: torch.Tensor, num_psi: float) -> torch.Tensor:
    """csmri._prox (proxfn/fast/csmri.py:14-25): masked closed-form update in centred ortho k-space, complex64 in / out."""
    if not v.is_cuda:
        raise RuntimeError(f"dprox_b200: v lives on {v.device}; this backend only computes on CUDA devices (no CPU fallback)")
    v = to_complex(v).to(torch.complex64).contiguous()
    v4 = as_bchw(v)
    B, Cc, H, W = v4.shape
    y = y.to(v.device, torch.complex64).expand(v4.shape).contiguous()
    m = cabi.require_cuda_f32(mask.to(v.device, torch.float32), "mask")
    m4 = as_bchw(m)
    if m4.shape[0] not in (1, B) or tuple(m4.shape[1:]) != (Cc, H, W):
        m4 = m4.expand(B, Cc, H, W).contiguous()
    rho = cabi.require_cuda_f32(torch.as_tensor(rho, dtype=torch.float32, device=v.device).reshape(-1), "rho")
    if rho.numel() not in (1, B):
        raise ValueError(f"rho must have 1 or B={B} entries")
    out = torch.empty_like(v)
    with torch.cuda.device(v.device):
        cabi.check(cabi.lib().dpx_csmri_prox(C.c_void_p(v.data_ptr()), C.c_void_p(y.data_ptr()), cabi.ptr(m4.contiguous()), m4.shape[0],
                                             cabi.ptr(rho), int(rho.numel() > 1), float(num_psi), C.c_void_p(out.data_ptr()),
                                             B, Cc, H, W, _s(v)), "dpx_csmri_prox")
    return out
