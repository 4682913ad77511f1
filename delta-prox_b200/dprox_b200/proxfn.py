"""ProxFn plugin surface (mirrors dprox.proxfn).

Native `_prox` bodies (nonneg, norm1, norm2/sum_squares, box) carry a `native_kind` that the plan
builder fuses into the prox+dual kernel; anything else (deep_prior, user subclasses overriding
`_prox`) is an *external* prox: the fused loop hands `K x + u` to the Python callable between two
native stages.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _cabi as cabi
from . import ops
from .linop import CompGraph, LinOp, Placeholder, Variable, adjoint as linop_adjoint, eval as linop_eval  # noqa: F401
from .tensors import to_torch_tensor


def safe_sqrt(x, eps=1e-8):
    """dprox/utils/misc.py:151-161."""
    return torch.sqrt(torch.clamp(x, min=eps))


class ProxFn(nn.Module):
    """f(K x) with a proximal operator (dprox/proxfn/base.py:30-108)."""

    native_kind: Optional[int] = None       # DPX_PROX_* when `_prox` has a native kernel
    box = (0.0, 0.0)

    def __init__(self, linop: LinOp = None, alpha=1, beta=1):
        super().__init__()
        self.linop = linop
        self.alpha = alpha
        self.beta = beta
        self.step = 0
        self.dag = CompGraph(linop, zero_out_constant=True)

    # -- constants --------------------------------------------------------------------------------
    @property
    def offset(self):
        """-linop.offset (proxfn/base.py:43-45)."""
        off = self.linop.offset
        return ops.axpby(-1.0, off) if off.is_cuda else -off

    def unwrap(self, value):
        if isinstance(value, Placeholder):
            return value.value
        return to_torch_tensor(value, batch=True).to(self.linop.device)

    # -- prox -------------------------------------------------------------------------------------
    def _prox(self, v, lam):
        raise NotImplementedError

    def eval(self, v):
        raise NotImplementedError

    def is_native(self) -> bool:
        """True when `_prox` is one of the library's kernels (not overridden by a subclass)."""
        return self.native_kind is not None

    def _offset_or_none(self, like):
        low = self.linop.lower()
        if low is not None:
            c = low.const_tensor(like)
            return None if c is None else ops.axpby(-1.0, c)
        off = self.offset
        return off

    def prox(self, v, lam):
        """translated(affine(scaled(_prox, alpha), beta), offset)  (proxfn/base.py:12-27, 55-64)."""
        lam = torch.as_tensor(lam, dtype=torch.float32, device=v.device)
        off = self._offset_or_none(v)
        if self.is_native():
            return ops.prox(self.native_kind, v, lam, self.alpha, self.beta, self.box[0], self.box[1], off)
        # external `_prox`: the wrapper chain around the user's callable, element-wise glue on the GPU
        lam4 = lam.view(lam.shape[0], *([1] * (v.ndim - 1))) if lam.ndim == 1 else lam
        w = v if off is None else ops.axpby(1.0, v, -1.0, off)
        if self.beta != 1:
            w = ops.axpby(self.beta, w)
        out = self._prox(w, self.beta * self.beta * lam4 * self.alpha)
        if self.beta != 1:
            out = ops.axpby(1.0 / self.beta, out)
        if off is not None:
            out = ops.axpby(1.0, out, 1.0, off)
        return out

    def convex_conjugate_prox(self, v, lam):
        """Moreau identity (proxfn/base.py:66-68)."""
        lamf = float(lam) if not isinstance(lam, torch.Tensor) or lam.numel() == 1 else None
        if lamf is None:
            raise NotImplementedError("per-sample lam in convex_conjugate_prox")
        return ops.axpby(1.0, v, -1.0, self.prox(ops.axpby(1.0 / lamf, v), lam))

    # -- algebra (proxfn/base.py:78-108) ------------------------------------------------------------
    def __mul__(self, other):
        if np.isscalar(other) and other > 0:
            self.alpha = other
            return self
        raise TypeError("Can only multiply by a positive scalar.")

    __rmul__ = __mul__

    def __add__(self, other):
        if isinstance(other, ProxFn):
            return [self, other]
        if type(other) == list:
            return [self] + other
        return NotImplemented

    def __radd__(self, other):
        if type(other) == list:
            return other + [self]
        return NotImplemented

    def __str__(self):
        return self.__class__.__name__


class nonneg(ProxFn):
    """Indicator of x >= 0 (proxfn/nonneg.py)."""
    native_kind = cabi.PROX_NONNEG

    def __init__(self, linop=None):
        super().__init__(linop)


class norm1(ProxFn):
    """|x|_1, soft threshold (proxfn/norm.py:6-19)."""
    native_kind = cabi.PROX_L1

    def __init__(self, linop=None):
        super().__init__(linop)


class norm2(ProxFn):
    """|x|_2^2 shrinkage v/(1+2 lam) (proxfn/norm.py:22-27)."""
    native_kind = cabi.PROX_L2SQ

    def __init__(self, linop=None):
        super().__init__(linop)


class box(ProxFn):
    """Indicator of lo <= x <= hi (new; named by the north star, absent from the reference)."""
    native_kind = cabi.PROX_BOX

    def __init__(self, linop=None, lo=0.0, hi=1.0):
        super().__init__(linop)
        self.box = (float(lo), float(hi))


class iso_tv(ProxFn):
    """Isotropic total variation  sum_pixels |(grad_H x, grad_W x)|_2 ; prox = group soft-threshold of the gradient
    pair.  New (north star); the reference only offers the anisotropic `norm1(grad(x, dim))` per axis."""
    native_kind = cabi.PROX_ISO_TV

    def __init__(self, arg):
        from .linop import grad2d
        super().__init__(arg if isinstance(arg, grad2d) else grad2d(arg))


class sum_squares(ProxFn):
    """|K x - b|_2^2 (proxfn/sum_square.py:12-32)."""
    native_kind = cabi.PROX_L2SQ

    def __init__(self, linop, b=None, eps=1e-7):
        super().__init__(linop)
        self.eps = eps
        self._b = b

    @property
    def offset(self):
        if self._b is not None:
            return self.unwrap(self._b)
        return super().offset

    def grad(self, x):
        """K^T (K x - b)  (sum_square.py:29-32)."""
        tmp = linop_eval(self.linop, x)
        off = self.offset.to(x.device).float()
        tmp = ops.axpby(1.0, tmp, -1.0, off)
        return linop_adjoint(self.linop, tmp)


class ext_sum_squares(sum_squares):
    """Quadratic data term with its own closed-form x-update (sum_square.py:35-48); subclasses provide
    `_prox(v, rho, num_psi)`.  Runs through the generic (node-by-node) engine."""

    def __init__(self, linop, eps=1e-7):
        super().__init__(linop, eps=eps)

    def setup(self, b):
        self.quad_b = b
        return self

    def solve(self, b, rho, eps=1e-6):
        xt = None
        for v in b:
            xt = v if xt is None else ops.axpby(1.0, xt, 1.0, v)
        return self._prox(xt, rho, len(b))


class csmri(ext_sum_squares):
    """CS-MRI data term |M F x - y|^2 with its closed-form x-update on complex iterates (proxfn/fast/csmri.py:8-25):
    `z = fft2(v); z[mask] = ((rho z + y) / (1 + rho n_psi))[mask]; ifft2(z)` with the centred ortho transforms of
    utils/misc.py:164-193, as ONE native call (`dpx_csmri_prox`: cuFFT C2C + fused masked update).  `mask` / `y` may be
    Placeholders (tests/paper/test_csmri.py:29-46)."""

    def __init__(self, linop, mask, y):
        super().__init__(linop)
        self.mask, self.y = mask, y

    def _prox(self, v, lam, num_psi):
        y = self.y.value if isinstance(self.y, Placeholder) else torch.as_tensor(self.y)
        m = self.mask.value if isinstance(self.mask, Placeholder) else torch.as_tensor(self.mask)
        return ops.csmri_prox(v, y, m, lam, float(num_psi))


# ------------------------------------------------------------------------------------------------
#  Plug-and-play prior                                   dprox/proxfn/pnp/prior.py:42-89
# ------------------------------------------------------------------------------------------------

class Denoiser(nn.Module):
    """Base class of denoisers: `denoise(input[B,C,H,W], sigma)` (pnp/denoisers/base.py:5-15)."""

    def denoise(self, input: torch.Tensor, sigma: torch.Tensor):
        sigma = sigma.reshape(-1, 1, 1, 1)
        return self._denoise(input, sigma)

    def _denoise(self, x, sigma):
        raise NotImplementedError


class Augment(nn.Module):
    """x8 test-time augmentation (pnp/denoisers/composite.py:6-28): call k denoises the image under flip / rotation k mod 8
    and undoes it afterwards (native permutation kernels, `ops.augment`)."""

    def __init__(self, base_denoiser):
        super().__init__()
        self.base_denoiser = base_denoiser
        self.iter = 0

    def denoise(self, x: torch.Tensor, sigma: torch.Tensor):
        m = self.iter % 8
        y = self.base_denoiser.denoise(ops.augment(x, m), sigma)
        y = ops.augment(y.contiguous(), 8 - m if m in (3, 5) else m)
        self.iter += 1
        return y

    def reset(self):
        self.iter = 0


class deep_prior(ProxFn):
    """Deep denoiser as a proximal operator (pnp/prior.py:42-89).  `denoiser` is a `Denoiser`
    instance (or any module with `.denoise(x, sigma)`); named pretrained models need their weight
    file, which cannot be downloaded here."""

    def __init__(self, linop, denoiser="ffdnet_color", x8=False, clamp=False, trainable=False, unroll_step=None,
                 sqrt=False):
        super().__init__(linop)
        self.name = denoiser if isinstance(denoiser, str) else type(denoiser).__name__
        if isinstance(denoiser, str):
            from .denoisers import get_denoiser
            denoiser = get_denoiser(denoiser)
        self.x8 = x8
        if x8:
            denoiser = Augment(denoiser)
        self.denoiser = denoiser
        self.clamp, self.sqrt = clamp, sqrt
        if not trainable:
            self.denoiser.eval()
            self.denoiser.requires_grad_(False)
        self.unroll = unroll_step is not None
        if self.unroll:
            import copy
            self.denoisers = nn.ModuleList([copy.deepcopy(self.denoiser) for _ in range(unroll_step)])

    def _reload(self, shape=None):
        if self.x8:
            self.denoiser.reset()

    def eval(self, v):
        raise NotImplementedError("deep prior cannot be explictly evaluated")

    def _prox(self, v, lam):
        sigma = safe_sqrt(lam) if self.sqrt else lam
        if self.clamp:
            v = v.clamp(0, 1)
        if torch.is_complex(v):                            # complex iterates (CS-MRI): the denoiser sees the real part (prior.py:79)
            v = ops.real_part(v)
        inp = v.unsqueeze(1) if v.ndim == 3 else v
        den = self.denoisers[self.step] if self.unroll else self.denoiser
        out = den.denoise(inp, sigma)          # under grad mode the tape runs through the denoiser (tests/test_grad.py:6-18)
        return out.type_as(v).reshape(v.shape).contiguous()

    def __repr__(self):
        return f'deep_prior(denoiser="{self.name}", unroll={self.unroll})'
