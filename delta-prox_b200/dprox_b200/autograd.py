"""Autograd contract of the hot path.

The reference is differentiable because it is eager PyTorch: unrolled training (`specialize(..., 'unroll')`,
algo/specialization/unroll.py:42-58), the DEQ backward hook and `LinearSolve.backward` all rely on
`iter/iters/solve` propagating gradients to the state, the measurements, `rhos` and `lams`.  Here every
stand-alone kernel of `ops` gets a `torch.autograd.Function` whose backward is again a native kernel:

  axpby / lincomb / mul / grad / spectral_filter   linear maps -> their adjoint kernels (+ fused dots for the
                                                   gradients of device-resident coefficients such as rho)
  prox (nonneg, l1, l2sq, box)                     `dpx_prox_backward` (mask / shrink derivative, sum for lam)
  xsolve (Fourier-diagonal x-update)               `dpx_xsolve_backward`: the closed form of App. D
                                                   (self-adjoint in the right-hand side, one reduction for rho)

`ops.*` switches to these only when grad mode is on and an input requires grad, so inference never pays for it.
PyTorch is the tape here, nothing else: no torch arithmetic runs in forward or backward.
"""
from __future__ import annotations

import torch

from . import _cabi as cabi


def needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


def _reduce_like(g: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """gradient of a broadcast operand (cold: constants broadcast over the batch)."""
    return g if g.shape == like.shape else g.sum_to_size(like.shape)


class Axpby(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, a, b):
        from . import ops
        ctx.a, ctx.b = a, b
        ctx.yshape = None if y is None else y.shape
        ctx.ylike = y
        return ops.axpby(a, x, b, y)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        g = g.contiguous()
        gx = ops.axpby(ctx.a, g) if ctx.needs_input_grad[0] else None
        gy = None
        if ctx.ylike is not None and ctx.needs_input_grad[1]:
            gy = _reduce_like(ops.axpby(ctx.b, g), ctx.ylike)
        return gx, gy, None, None


class Lincomb(torch.autograd.Function):
    """out = a*x + b*y + c*z with device coefficients ([1] or [B]; None = 1)."""

    @staticmethod
    def forward(ctx, x, a, y, b, z, c):
        from . import ops
        ctx.save_for_backward(x, a, y, b, z, c)
        return ops.lincomb(x, a, y, b, z, c)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        g = g.contiguous()
        x, a, y, b, z, c = ctx.saved_tensors
        out = []
        for i, (t, k) in enumerate(((x, a), (y, b), (z, c))):
            gt = gk = None
            if t is not None and ctx.needs_input_grad[2 * i]:
                gt = ops.lincomb(g, k)
            if t is not None and k is not None and ctx.needs_input_grad[2 * i + 1]:
                per = k.numel() > 1
                gk = ops.dot(g, t, per_sample=per).reshape(k.shape)
            out += [gt, gk]
        return tuple(out)


class Mul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        from . import ops
        ctx.save_for_backward(w)
        return ops.mul(x, w)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        (w,) = ctx.saved_tensors
        return ops.mul(g.contiguous(), w), None


class Grad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, axis, adjoint, scale):
        from . import ops
        ctx.cfg = (axis, adjoint, scale)
        return ops.grad(x, axis, adjoint, scale)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        axis, adjoint, scale = ctx.cfg
        return ops.grad(g.contiguous(), axis, not adjoint, scale), None, None, None


class Pad2d(torch.autograd.Function):
    """zero-pad / crop copy; its adjoint is the same copy with the offsets negated back to the input size"""

    @staticmethod
    def forward(ctx, x, out_hw, top, left):
        from . import ops
        ctx.cfg = (tuple(x.shape[-2:]), top, left)
        return ops.pad2d(x, out_hw, top, left)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        in_hw, top, left = ctx.cfg
        return ops.pad2d(g.contiguous(), in_hw, -top, -left), None, None, None


class Augment(torch.autograd.Function):
    """a pixel permutation: the adjoint is the inverse permutation"""

    @staticmethod
    def forward(ctx, x, mode):
        from . import ops
        ctx.mode = mode
        return ops.augment(x, mode)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        return ops.augment(g.contiguous(), ops.AUGMENT_INVERSE[ctx.mode]), None


class SpectralFilter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, otf, conj, plan):
        from . import ops
        ctx.save_for_backward(otf)
        ctx.cfg = (conj, plan)
        return ops.spectral_filter(x, otf, conj, plan)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        (otf,) = ctx.saved_tensors
        conj, plan = ctx.cfg
        return ops.spectral_filter(g.contiguous(), otf, not conj, plan), None, None, None


class Prox(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, lam, offset, kind, alpha, beta, lo, hi):
        from . import ops
        ctx.save_for_backward(v, lam, offset)
        ctx.cfg = (kind, alpha, beta, lo, hi)
        return ops.prox(kind, v, lam, alpha, beta, lo, hi, offset)

    @staticmethod
    def backward(ctx, g):
        v, lam, offset = ctx.saved_tensors
        kind, alpha, beta, lo, hi = ctx.cfg
        if kind not in (cabi.PROX_NONNEG, cabi.PROX_L1, cabi.PROX_L2SQ, cabi.PROX_BOX):
            raise NotImplementedError(f"prox kind {kind} has no native backward")
        g = cabi.require_cuda_f32(g, "grad")
        B = v.shape[0] if v.ndim > 0 else 1
        lam_f = cabi.require_cuda_f32(lam.to(v.device, torch.float32).reshape(-1), "lam")
        off = None if offset is None else cabi.require_cuda_f32(offset.to(v.device).expand_as(v), "offset")
        gv = torch.empty_like(v)
        glam = torch.empty(lam_f.numel(), device=v.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(v.device):
            cabi.check(cabi.lib().dpx_prox_backward(int(kind), cabi.ptr(v), cabi.ptr(lam_f), int(lam_f.numel() > 1), float(alpha),
                                                    float(beta), float(lo), float(hi), cabi.ptr(off), cabi.ptr(g), cabi.ptr(gv),
                                                    cabi.ptr(glam), B, v.numel() // B, cabi.stream_ptr(v.device)),
                       "dpx_prox_backward")
        goff = None
        if offset is not None and ctx.needs_input_grad[2]:
            from . import ops
            goff = _reduce_like(ops.axpby(1.0, g, -1.0, gv), offset)
        return gv, (None if glam is None else glam.reshape(lam.shape)), goff, None, None, None, None, None


class XSolve(torch.autograd.Function):
    """x = F^-1[(F(ktb) + rho F(t) + eps) / (dq + rho (dpsi + wid) + eps)] through a FREQ_DIAG plan whose constants hold
    F(ktb) and the diagonals.  `ktb` (= sum_q A_q^T b_q, built by differentiable ops) only receives its gradient here."""

    @staticmethod
    def forward(ctx, plan, t, rho, rho_stride, ktb):
        x = torch.empty_like(t)
        with torch.cuda.device(t.device):
            cabi.check(cabi.lib().dpx_xsolve(plan.handle, cabi.ptr(t), cabi.ptr(rho), rho_stride, 0, cabi.ptr(x),
                                             cabi.stream_ptr(t.device)), "dpx_xsolve")
        ctx.plan, ctx.rho_stride, ctx.const_version = plan, rho_stride, plan.const_version
        ctx.save_for_backward(x, rho)
        ctx.ktb_shape = None if ktb is None else ktb.shape
        return x

    @staticmethod
    def backward(ctx, g):
        from . import ops
        x, rho = ctx.saved_tensors
        if ctx.plan.const_version != ctx.const_version:
            # the backward kernel reads F(K^T b) and the diagonals from the plan: a second forward with other measurements /
            # operator parameters in between would silently give wrong rho gradients
            raise RuntimeError("dprox_b200: the solver's constants (measurements / Placeholder values) changed between this "
                               "forward pass and its backward; call backward() before the next solve(), or use one compiled "
                               "solver per in-flight graph")
        g = cabi.require_cuda_f32(g, "grad")
        need_rho = ctx.needs_input_grad[2]
        gk = torch.empty_like(x)
        grho = torch.empty(rho.numel() if ctx.rho_stride else 1, device=x.device, dtype=torch.float32) if need_rho else None
        with torch.cuda.device(x.device):
            cabi.check(cabi.lib().dpx_xsolve_backward(ctx.plan.handle, cabi.ptr(g), cabi.ptr(x), cabi.ptr(rho), ctx.rho_stride, 0,
                                                      cabi.ptr(gk), cabi.ptr(grho), cabi.stream_ptr(x.device)),
                       "dpx_xsolve_backward")
        gt = ops.lincomb(gk, rho.reshape(-1)) if ctx.needs_input_grad[1] else None
        gktb = None
        if ctx.ktb_shape is not None and ctx.needs_input_grad[4]:
            gktb = gk.reshape(ctx.ktb_shape) if gk.numel() == int(torch.Size(ctx.ktb_shape).numel()) else gk.sum_to_size(ctx.ktb_shape)
        return None, gt, (None if grho is None else grho.reshape(rho.shape)), None, gktb
