"""LinOp plugin surface (mirrors dprox.linop) on top of the C-ABI kernels.

Design (not a port of the reference's Edge/CompGraph interpreter): a LinOp expression is a plain
tree of `nn.Module` nodes.  Two things can be done with a tree:

  * evaluate it — `evaluate()` / `evaluate_adjoint()` walk it recursively and call each node's
    `forward` / `adjoint`, which run sm_100a kernels through `dprox_b200.ops`;
  * lower it — `lower()` folds the tree into a `Lowered` normal form  K(x) = s * A x + c  with
    A in {identity, grad_H, grad_W, spectral(OTF), mask} which the plan builder
    (`dprox_b200.lowering`) turns into a fused native plan.  Trees that do not fold (BlackBox,
    mosaic(conv(x)), ...) stay "generic" and run node by node + CG.

User plugins subclass `LinOp` with `forward/adjoint[/is_diag/is_gram_diag/get_diag]` exactly as in the
reference (docs/source/api/linop.md:5-30) or use `LinOpFactory`.
"""
from __future__ import annotations

import builtins as _builtins
import copy as _copy
import uuid as _uuid
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .tensors import to_ndarray, to_torch_tensor


def _as_linop(x) -> "LinOp":
    return x if isinstance(x, LinOp) else Constant(x)


# ------------------------------------------------------------------------------------------------
#  Lowered normal form
# ------------------------------------------------------------------------------------------------

@dataclass
class Lowered:
    """K(x) = scale * A x + const, A described by `kind`.

    kind      : 'identity' | 'grad' | 'spectral' | 'mask'
    otf_fn    : shape[B,C,H,W] -> complex64 half-spectrum ndarray/tensor [1|B,C,H,W/2+1]   (spectral / grad)
    gram_fn   : shape -> real |OTF|^2 half spectrum (float32)                               (spectral / grad)
    mask      : tensor broadcastable to [B,C,H,W]                                           (mask)
    const     : list of (coefficient, tensor-or-callable) making up the constant part
    """
    kind: str = "identity"
    scale: float = 1.0
    axis: int = -1
    otf_fn: Optional[Callable] = None
    gram_fn: Optional[Callable] = None
    mask: Optional[torch.Tensor] = None
    const: List = field(default_factory=list)

    def scaled(self, s: float) -> "Lowered":
        return Lowered(self.kind, self.scale * s, self.axis, self.otf_fn, self.gram_fn, self.mask,
                       [(c * s, t) for c, t in self.const])

    def const_tensor(self, like: torch.Tensor) -> Optional[torch.Tensor]:
        """The constant part as one tensor expanded to `like` (None when there is none)."""
        if not self.const:
            return None
        acc = None
        for coef, t in self.const:
            t = t() if callable(t) else t
            # constants broadcast against the variable exactly as in `sum.forward` (no re-batching)
            t = torch.as_tensor(t).to(like.device, torch.float32)
            if t.shape != like.shape:
                t = torch.broadcast_to(t, like.shape).contiguous()
            acc = ops.axpby(coef, t) if acc is None else ops.axpby(1.0, acc, coef, t)
        return acc


# ------------------------------------------------------------------------------------------------
#  Base class
# ------------------------------------------------------------------------------------------------

class LinOp(nn.Module):
    """Abstract linear operator node (dprox/linop/base.py:18-247)."""

    class MultOutput(list):
        pass

    instanceCnt = 0
    __array_priority__ = 10000

    def __init__(self, input_nodes: Sequence = ()):
        super().__init__()
        self.input_nodes = nn.ModuleList([_as_linop(n) for n in input_nodes])
        self.linop_id = LinOp.instanceCnt
        LinOp.instanceCnt += 1
        self.dummy = nn.Parameter(torch.tensor(0.0), requires_grad=False)     # tracks the device
        self.step = 0

    # -- plugin protocol ---------------------------------------------------------------------
    def forward(self, *inputs, **kwargs):
        raise NotImplementedError

    def adjoint(self, *inputs, **kwargs):
        raise NotImplementedError

    def is_gram_diag(self, freq: bool = False) -> bool:
        return self.is_diag(freq)

    def is_diag(self, freq: bool = False) -> bool:
        return False

    def get_diag(self, ref, freq: bool = False):
        raise NotImplementedError

    def norm_bound(self, input_mags):
        return NotImplemented

    # -- lowering hook -----------------------------------------------------------------------
    def lower(self) -> Optional[Lowered]:
        """Fold this subtree into the normal form, or None if it does not fold."""
        return None

    # -- structure ---------------------------------------------------------------------------
    @property
    def device(self):
        return self.dummy.device

    @property
    def variables(self) -> List["Variable"]:
        seen, out = set(), []
        for n in self.input_nodes:
            for v in n.variables:
                if v.uuid not in seen:
                    seen.add(v.uuid)
                    out.append(v)
        return sorted(out, key=lambda v: v.uuid)

    @property
    def constants(self) -> List["Constant"]:
        out = []
        for n in self.input_nodes:
            out += n.constants
        return out

    def is_constant(self) -> bool:
        return len(self.variables) == 0

    @property
    def value(self):
        return self.forward(*[n.value for n in self.input_nodes])

    @property
    def offset(self):
        """Value of the expression with every variable set to zero (linop/base.py:118-129)."""
        low = self.lower()
        vars_ = self.variables
        if low is not None and vars_ and vars_[0]._value is not None:
            like = vars_[0].value
            c = low.const_tensor(like)
            return c if c is not None else torch.zeros_like(like)
        saved = {v: v._value for v in vars_}
        try:
            for v in vars_:
                v._value = torch.zeros_like(v.value)
            return self.value
        finally:
            for v, val in saved.items():
                v._value = val

    @property
    def T(self) -> "LinOp":
        op = self.clone()
        op.forward, op.adjoint = op.adjoint, op.forward
        return op

    @property
    def gram(self) -> "LinOp":
        op = self.clone()
        f, a = op.forward, op.adjoint
        op.forward = lambda x: a(f(x))
        op.adjoint = lambda x: f(a(x))
        return op

    def clone(self) -> "LinOp":
        return _copy.deepcopy(self)

    def unwrap(self, value):
        if isinstance(value, Placeholder):
            return value.value
        return to_torch_tensor(value, batch=True)

    # -- operator overloading (linop/base.py:181-232) -------------------------------------------
    def __add__(self, other):
        other = _as_linop(other)
        args = []
        for e in (self, other):
            args += list(e.input_nodes) if isinstance(e, sum) else [e]
        return sum(args)

    def __radd__(self, other):
        return _as_linop(other) + self

    def __mul__(self, other):
        if np.isscalar(other):
            return scale(other, self)
        raise TypeError("Can only multiply by a scalar constant.")

    __rmul__ = __mul__

    def __truediv__(self, other):
        if np.isscalar(other):
            return scale(1.0 / other, self)
        raise TypeError("Can only divide by a scalar constant.")

    def __sub__(self, other):
        return self + (-_as_linop(other))

    def __rsub__(self, other):
        return (-self) + other

    def __neg__(self):
        return -1 * self

    def __str__(self):
        return self.__class__.__name__


# ------------------------------------------------------------------------------------------------
#  Leaves
# ------------------------------------------------------------------------------------------------

class Variable(LinOp):
    """The optimisation variable (linop/variable.py:8-100)."""

    def __init__(self, shape=None, value=None, name=None):
        super().__init__([])
        self.uuid = _uuid.uuid1()
        self._value = value
        self.shape = shape
        self.varname = name

    def forward(self, inputs, **kw):
        return inputs

    def adjoint(self, inputs, **kw):
        return inputs

    def is_diag(self, freq=False):
        return True

    def get_diag(self, ref, freq=False):
        return torch.ones(ref.shape, device=ref.device)

    def lower(self):
        return Lowered("identity")

    @property
    def variables(self):
        return [self]

    @property
    def value(self):
        if self._value is None:
            raise RuntimeError("Variable has no value yet (it is set by Algorithm.solve / iter)")
        return self._value

    @value.setter
    def value(self, val):
        self._value = val

    def norm_bound(self, input_mags):
        return 1.0

    def __repr__(self):
        return f"Variable(id={self.uuid}, shape={self.shape})"


class Constant(LinOp):
    """A constant leaf (linop/constant.py:7-96)."""

    def __init__(self, value):
        super().__init__([])
        if value is not None and not isinstance(value, torch.Tensor):
            value = torch.tensor(np.asarray(value))
        self._value = value

    def forward(self, *value, **kw):
        return self.value

    def adjoint(self, value, **kw):
        return None            # constants have no variable underneath: nothing flows back

    def is_diag(self, freq=False):
        return True

    def get_diag(self, ref=None, freq=False):
        return {}

    @property
    def variables(self):
        return []

    @property
    def constants(self):
        return [self]

    @property
    def value(self):
        v = self._value
        if v is None:
            raise RuntimeError("Placeholder/Constant has no value")
        if v.is_floating_point() and v.dtype != torch.float32:
            v = v.float()
        elif not v.is_floating_point() and not v.is_complex():
            v = v.float()
        return v.to(self.device) if self.device.type != "cpu" else v

    def lower(self):
        return Lowered("const", 0.0, const=[(1.0, lambda: self.value)])

    def norm_bound(self, input_mags):
        return 0.0

    def __repr__(self):
        return "Constant(value=%s)" % ("None" if self._value is None else "somevalue")


class Placeholder(Constant):
    """A constant whose value is supplied later; watchers are notified (linop/placeholder.py:4-22)."""

    def __init__(self, default=None):
        super().__init__(default)
        self.watchers = []
        self.version = 0

    @property
    def value(self):
        return Constant.value.fget(self)

    @value.setter
    def value(self, val):
        self._value = val if isinstance(val, torch.Tensor) or val is None else torch.tensor(np.asarray(val))
        self.version += 1
        for w in self.watchers:
            w(val)

    def change(self, fn):
        self.watchers.append(fn)


# ------------------------------------------------------------------------------------------------
#  Glue nodes
# ------------------------------------------------------------------------------------------------

class scale(LinOp):
    """scalar * X (linop/scale.py:7-80)."""

    def __init__(self, scalar, arg):
        assert np.isscalar(scalar)
        self.scalar = float(scalar)
        super().__init__([arg])

    def forward(self, input, **kw):
        return ops.axpby(self.scalar, input)

    def adjoint(self, input, **kw):
        return ops.axpby(self.scalar, input)

    def is_gram_diag(self, freq=False):
        return self.input_nodes[0].is_gram_diag(freq)

    def is_diag(self, freq=False):
        return self.input_nodes[0].is_diag(freq)

    def get_diag(self, ref, freq=False):
        # Gram diagonal of s*K is s^2 * gram(K).  (The reference squares the child's Gram diagonal as well,
        # scale.py:43-57 — identical for Variable children, which is all its own tests exercise.)
        return self.input_nodes[0].get_diag(ref, freq) * (self.scalar * self.scalar)

    def lower(self):
        child = self.input_nodes[0].lower()
        return None if child is None else child.scaled(self.scalar)

    def norm_bound(self, input_mags):
        return abs(self.scalar) * input_mags[0]


class sum(LinOp):  # noqa: A001  (the reference shadows the builtin on purpose)
    """Sums its inputs (linop/sum.py:6-75)."""

    def __init__(self, input_nodes):
        super().__init__(input_nodes)

    def forward(self, *inputs, **kw):
        ins = [i for i in inputs if i is not None]
        ref = max(ins, key=lambda t: t.numel())
        out = None
        for t in ins:
            t = t.to(ref.device)
            if t.dtype != torch.float32:
                t = t.float()
            if t.shape != ref.shape:
                t = t.expand_as(ref).contiguous()
            out = ops.axpby(1.0, t) if out is None else ops.axpby(1.0, out, 1.0, t)
        return out

    def adjoint(self, input, **kw):
        outs = LinOp.MultOutput([input for _ in self.input_nodes])
        return outs if len(outs) > 1 else outs[0]

    def is_diag(self, freq=False):
        return all(a.is_diag(freq) for a in self.input_nodes)

    def is_gram_diag(self, freq=False):
        return all(a.is_gram_diag(freq) for a in self.input_nodes)

    def get_diag(self, ref, freq=False):
        for n in self.input_nodes:                       # the non-constant branch carries the diagonal
            if not isinstance(n, Constant) and n.variables:
                return n.get_diag(ref, freq)
        return self.input_nodes[0].get_diag(ref, freq)

    def lower(self):
        parts = [n.lower() for n in self.input_nodes]
        if any(p is None for p in parts):
            return None
        lin = [p for p in parts if p.kind != "const"]
        if len(lin) != 1:
            return None                                  # e.g. conv(x) + x: not folded (runs generically)
        out = Lowered(lin[0].kind, lin[0].scale, lin[0].axis, lin[0].otf_fn, lin[0].gram_fn, lin[0].mask, list(lin[0].const))
        for p in parts:
            if p.kind == "const":
                out.const += p.const
        return out


class copy(sum):  # noqa: A001
    """Fan-out node (linop/sum.py:77-109); kept for API compatibility."""

    def __init__(self, arg):
        super().__init__([arg])

    def forward(self, inputs, **kw):
        return sum.adjoint(self, inputs)

    def adjoint(self, *inputs, **kw):
        return sum.forward(self, *inputs)


class vstack(LinOp):
    """Stacks the outputs of its inputs (linop/vstack.py:6-84)."""

    def __init__(self, input_nodes):
        super().__init__(input_nodes)

    def forward(self, *inputs, **kw):
        return LinOp.MultOutput(inputs) if len(inputs) > 1 else inputs[0]

    def adjoint(self, *inputs, **kw):
        return LinOp.MultOutput(inputs) if len(inputs) > 1 else inputs[0]

    def is_gram_diag(self, freq=False):
        return all(a.is_gram_diag(freq) for a in self.input_nodes)


class split(vstack):
    def __init__(self, output_nodes):
        super().__init__(output_nodes)


# ------------------------------------------------------------------------------------------------
#  OTF construction (cold path, host)                      dprox/utils/psf2otf.py, linop/conv.py:59-80
# ------------------------------------------------------------------------------------------------

def kernel_otf(kernel: np.ndarray, H: int, W: int, C: int) -> np.ndarray:
    """Full-spectrum OTF [C,H,W] of a small HWC kernel with the reference's conventions: zero-pad after
    the kernel, roll the centre tap floor(k/2) to the origin on every axis, DFT over H, W *and* C (so a
    (k,k,1) kernel is broadcast to all channels and a genuine multi-channel kernel is channel-mixed,
    SURVEY App. A-2), in the kernel's own precision; imaginary round-off is dropped like real_if_close."""
    k = np.asarray(kernel)
    while k.ndim < 3:
        k = k[..., None]
    if k.shape[0] > H or k.shape[1] > W or k.shape[2] > C:
        raise ValueError(f"outsize {[H, W, C]} cannot be smaller than the kernel {list(k.shape)}")
    if not np.any(k):
        return np.zeros((C, H, W), dtype=np.float32)
    big = np.zeros((H, W, C), dtype=k.dtype)
    big[:k.shape[0], :k.shape[1], :k.shape[2]] = k
    big = np.roll(big, tuple(-(s // 2) for s in k.shape), axis=(0, 1, 2))
    otf = np.fft.fftn(big)
    tol = float(np.sum(big.size * np.log2(big.shape)))
    otf = np.real_if_close(otf, tol=tol)
    return np.ascontiguousarray(np.moveaxis(otf, 2, 0))


def _half(otf_chw) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(otf_chw)) if isinstance(otf_chw, np.ndarray) else otf_chw
    W = t.shape[-1]
    return t[..., : W // 2 + 1].to(torch.complex64).contiguous()


def psf2otf2(psf: torch.Tensor, out_shape) -> torch.Tensor:
    """conv_doe's OTF (linop/conv.py:59-80) incl. its even-pad off-by-one and the all-axes ifftshift
    (channel roll).  Cold path: built with torch.fft once per PSF value."""
    fh = psf.shape[2]
    if out_shape[2] != fh:
        pad = (out_shape[2] - fh) / 2
        if (out_shape[2] - fh) % 2 != 0:
            lo, hi = int(np.ceil(pad)), int(np.floor(pad))
        else:
            lo, hi = int(pad) + 1, int(pad) - 1
        psf = torch.nn.functional.pad(psf, [lo, hi, lo, hi])
    return torch.fft.fft2(torch.fft.ifftshift(psf))


# ------------------------------------------------------------------------------------------------
#  Operators
# ------------------------------------------------------------------------------------------------

class conv(LinOp):
    """Circular convolution with a fixed kernel (linop/conv.py:15-56)."""

    def __init__(self, arg, kernel):
        self.kernel = to_ndarray(kernel)
        self.cache = {}
        super().__init__([arg])

    # full-spectrum OTF as the reference exposes it ([1,C,H,W]); half spectrum for the kernels
    def _FB(self, shape):
        shape = tuple(shape)
        if shape not in self.cache:
            _, C, H, W = shape
            full = torch.from_numpy(kernel_otf(self.kernel, H, W, C)).unsqueeze(0)
            self.cache[shape] = (full, _half(full))
        return self.cache[shape][0]

    def _otf_half(self, shape):
        self._FB(shape)
        return self.cache[tuple(shape)][1]

    def _gram_half(self, shape):
        fb = self._FB(shape)
        g = (fb.conj() * fb).real if fb.is_complex() else fb * fb
        return g[..., : shape[-1] // 2 + 1].float().contiguous()

    def _otf_on(self, shape, device):
        """half-spectrum OTF resident on `device` (uploaded once per shape and device: the host copy is 8 B x H x W/2 per
        channel, and re-sending it on every forward/adjoint stalled the host for milliseconds)"""
        key = (tuple(shape), str(device))
        if key not in self.cache:
            self.cache[key] = self._otf_half(shape).to(device, torch.complex64).contiguous()
        return self.cache[key]

    def forward(self, input, **kw):
        return ops.spectral_filter(input, self._otf_on(_shape4(input), input.device), conj=False)

    def adjoint(self, input, **kw):
        return ops.spectral_filter(input, self._otf_on(_shape4(input), input.device), conj=True)

    def is_diag(self, freq=False):
        return freq and self.input_nodes[0].is_diag(freq)

    def get_diag(self, x, freq=False):
        assert freq
        fb = self._FB(_shape4(x))
        return torch.abs(torch.conj(fb) * fb).to(self.device)

    def lower(self):
        child = self.input_nodes[0].lower()
        if child is None or child.const:
            return None
        if child.kind == "identity":
            return Lowered("spectral", child.scale, otf_fn=self._otf_half, gram_fn=self._gram_half)
        if child.kind in ("spectral", "grad") and child.otf_fn is not None:
            c_otf, c_gram = child.otf_fn, child.gram_fn
            return Lowered("spectral", child.scale,
                           otf_fn=lambda s: self._otf_half(s) * _t(c_otf(s)),
                           gram_fn=lambda s: self._gram_half(s) * _t(c_gram(s)))
        return None


def _t(a):
    return torch.from_numpy(a) if isinstance(a, np.ndarray) else a


def _shape4(x):
    s = tuple(x.shape)
    return (1,) * (4 - len(s)) + s


class grad(conv):
    """Circular forward difference along H (dim=0) or W (dim=1)  (linop/grad.py:8-23).

    The reference implements it as an FFT convolution with kernel [1,-1] (in double precision,
    because the kernel tensor is int64); in pixel space that is exactly y[i] = x[i+1] - x[i] with
    wrap-around, which is what the stencil kernel computes."""

    def __init__(self, arg, dim=1):
        if dim not in (0, 1, 2):
            raise ValueError("dim must be 0(Height) or 1(Width) or 2 (Channel)")
        self.dim = dim
        k = np.array([1, -1], dtype=np.int64).reshape(1, 1, 2)
        LinOp.__init__(self, [arg])
        self.kernel = np.swapaxes(k, dim, -1)
        self.cache = {}

    def _chan_weight(self, input):
        """dim=2: the kernel [1,-1] lies on the channel axis of the HWC kernel, and `conv` transforms the kernel over all three
        axes but the image only over H, W (conv.py:31-41, psf2otf.py): channel c is multiplied by the complex constant FB[c],
        and the real part is kept -- y[c] = Re(FB[c]) x[c] for forward AND adjoint.  Kept as the reference has it."""
        shape = _shape4(input)
        key = ("cw", shape[1], str(input.device))
        if key not in self.cache:
            fb = self._FB((1, shape[1], 8, 8))                     # constant over (h, w)
            w = torch.real(fb[:, :, 0, 0]).float().reshape(1, shape[1], 1, 1)
            self.cache[key] = w.to(input.device)
        return self.cache[key].expand(1, shape[1], shape[2], shape[3]).contiguous()

    def forward(self, input, **kw):
        if self.dim == 2:
            return ops.mul(input, self._chan_weight(input))
        return ops.grad(input, self.dim, adjoint=False)

    def adjoint(self, input, **kw):
        if self.dim == 2:
            return ops.mul(input, self._chan_weight(input))
        return ops.grad(input, self.dim, adjoint=True)

    def lower(self):
        child = self.input_nodes[0].lower()
        if child is None or child.const:
            return None
        if self.dim == 2:
            # Fourier-diagonalisable with the diagonal |FB[c]|^2 (what the reference's closed form divides by) while
            # forward / adjoint apply Re(FB[c]): `otf_fn=None` keeps every application on this node's own forward/adjoint
            return Lowered("spectral", child.scale, otf_fn=None, gram_fn=self._gram_half) if child.kind == "identity" else None
        if child.kind == "identity":
            return Lowered("grad", child.scale, axis=self.dim, otf_fn=self._otf_half, gram_fn=self._gram_half)
        return conv.lower(self)


class grad2d(LinOp):
    """[grad_H x ; grad_W x] stacked on the channel axis ([B,C,H,W] -> [B,2C,H,W]): the operator of isotropic TV
    (`iso_tv`).  Not in the reference (which only has per-axis `grad`); named by the north star."""

    def __init__(self, arg):
        super().__init__([arg])
        self._gh, self._gw = grad(Variable(), dim=0), grad(Variable(), dim=1)

    def forward(self, input, **kw):
        return torch.cat([ops.grad(input, 0), ops.grad(input, 1)], dim=1)

    def adjoint(self, input, **kw):
        C = input.shape[1] // 2
        a = ops.grad(input[:, :C].contiguous(), 0, adjoint=True)
        b = ops.grad(input[:, C:].contiguous(), 1, adjoint=True)
        return ops.axpby(1.0, a, 1.0, b)

    def is_diag(self, freq=False):
        return freq and self.input_nodes[0].is_diag(freq)

    def _gram_half(self, shape):
        return self._gh._gram_half(shape) + self._gw._gram_half(shape)

    def get_diag(self, x, freq=False):
        assert freq
        return self._gh.get_diag(x, True) + self._gw.get_diag(x, True)

    def lower(self):
        child = self.input_nodes[0].lower()
        if child is None or child.const or child.kind != "identity":
            return None
        return Lowered("grad2d", child.scale, gram_fn=self._gram_half)


def linear_conv_pads(H: int, W: int):
    """zero padding of the `circular=False` mode (linop/conv.py:103-110): BOTH axes are padded towards 2 * H (the height),
    ceil before / floor after -> (top, bottom, left, right)."""
    target = 2 * H
    hp, wp = (target - H) / 2, (target - W) / 2
    return int(np.ceil(hp)), int(np.floor(hp)), int(np.ceil(wp)), int(np.floor(wp))


class conv_doe(LinOp):
    """Convolution with a (learnable / Placeholder-fed) PSF [1,C,h,w]  (linop/conv.py:83-156): circular, or -- `circular=False`
    -- linear: zero-pad to twice the size, convolve circularly there, crop (:100-121).  As in the reference the linear mode
    still reports the CIRCULAR |OTF|^2 at the image size as its Fourier diagonal (:143-152).  Gradients w.r.t. the PSF are
    not propagated through this node (the reference re-wraps it as a fresh leaf, :91-96)."""

    def __init__(self, arg, psf, circular: bool = True):
        super().__init__([arg])
        self._psf = psf
        self.circular = circular
        self._otf_cache = {}
        if isinstance(psf, Placeholder):
            def on_change(val):
                self.psf = nn.Parameter(val, requires_grad=False)
                self._otf_cache = {}
            psf.change(on_change)
            if psf._value is not None:
                on_change(psf._value)
        else:
            self.psf = nn.Parameter(to_torch_tensor(psf, batch=True).float(), requires_grad=False)

    def _otf(self, shape):
        key = (tuple(shape), self.psf.data_ptr(), self.psf._version)
        hit = self._otf_cache.get(tuple(shape))
        if hit is None or hit[0] != key:
            full = psf2otf2(self.psf.detach(), shape)
            hit = (key, full, _half(full))
            self._otf_cache[tuple(shape)] = hit
        return hit

    def _otf_half(self, shape):
        return self._otf(shape)[2]

    def _gram_half(self, shape):
        o = self._otf_half(shape)
        return (o.conj() * o).real.float().contiguous()

    def _conv(self, img, conj):
        if self.circular:
            return ops.spectral_filter(img, self._otf_half(_shape4(img)), conj=conj)
        B, Cc, H, W = _shape4(img)
        pt, pb, pl, pr = linear_conv_pads(H, W)
        big = ops.pad2d(img.reshape(B, Cc, H, W), (H + pt + pb, W + pl + pr), pt, pl)
        out = ops.spectral_filter(big, self._otf_half(tuple(big.shape)), conj=conj)
        # the reference crops with [pt:-pb, pl:-pr]
        return ops.pad2d(out, (H, W), -pt, -pl).reshape(img.shape)

    def forward(self, img, **kw):
        return self._conv(img, False)

    def adjoint(self, img, **kw):
        return self._conv(img, True)

    def is_diag(self, freq=False):
        return freq and self.input_nodes[0].is_diag(freq)

    def get_diag(self, x, freq=False):
        assert freq
        full = self._otf(_shape4(x))[1]
        return torch.abs(torch.conj(full) * full).to(self.device)

    def lower(self):
        child = self.input_nodes[0].lower()
        if child is None or child.const or child.kind != "identity":
            return None
        # linear mode: K^T b must come from this node's own (padded) adjoint, so no OTF is offered for spectral shortcuts
        return Lowered("spectral", child.scale, otf_fn=self._otf_half if self.circular else None, gram_fn=self._gram_half)


def bayer_mask(H: int, W: int) -> torch.Tensor:
    """RGGB colour-filter-array mask [1,3,H,W] (linop/subsample.py:34-48)."""
    m = torch.zeros(1, 3, H, W)
    m[0, 0, 0::2, 0::2] = 1
    m[0, 1, 0::2, 1::2] = 1
    m[0, 1, 1::2, 0::2] = 1
    m[0, 2, 1::2, 1::2] = 1
    return m


class mosaic(LinOp):
    """Bayer RGGB mosaicking mask (linop/subsample.py:8-80)."""

    def __init__(self, arg):
        super().__init__([arg])
        self.cache = {}

    def _mask(self, shape, device=None):
        key = (tuple(shape[-2:]), str(device))
        if key not in self.cache:
            self.cache[key] = bayer_mask(*shape[-2:]).to(device or "cpu")
        return self.cache[key]

    def forward(self, input, **kw):
        return ops.mul(input, self._mask(input.shape, input.device))

    def adjoint(self, input, **kw):
        return self.forward(input)

    def is_gram_diag(self, freq=False):
        return (not freq) and self.input_nodes[0].is_diag(freq)

    def is_self_diag(self, freq=False):
        return not freq

    def get_diag(self, x, freq=False):
        assert not freq
        return self._mask(x.shape, x.device)

    def lower(self):
        child = self.input_nodes[0].lower()
        if child is None or child.const or child.kind != "identity":
            return None
        return Lowered("mask", child.scale, mask=lambda shape, dev: self._mask(shape, dev))

    def norm_bound(self, input_mags):
        return input_mags[0]


class mul_elementwise(LinOp):
    """Element-wise multiplication by a fixed weight (linop/mul.py:44-73)."""

    def __init__(self, arg, w):
        super().__init__([arg])
        self._w = w
        if isinstance(w, Placeholder):
            def on_change(val):
                self.w = nn.Parameter(val, requires_grad=False)
            w.change(on_change)
            if w._value is not None:
                on_change(w._value)
        else:
            self.w = nn.Parameter(to_torch_tensor(w, batch=True).float(), requires_grad=False)

    def forward(self, x, **kw):
        return ops.mul(x, self.w)

    def adjoint(self, x, **kw):
        return self.forward(x)

    def is_diag(self, freq=False):
        return (not freq) and self.input_nodes[0].is_diag(freq)

    def get_diag(self, x, freq=False):
        return None if freq else self.w.to(x.device)

    def lower(self):
        child = self.input_nodes[0].lower()
        if child is None or child.const or child.kind != "identity":
            return None
        return Lowered("mask", child.scale, mask=lambda shape, dev: self.w.detach().to(dev))


class BlackBox(LinOp):
    """User-supplied operator `(x, step=) -> y` (linop/blackbox.py:25-78); always runs generically."""

    def __init__(self, *args, forward=None, adjoint=None, diag=None, norm_bound=None):
        self._forward, self._adjoint, self._norm_bound, self._diag = forward, adjoint, norm_bound, diag
        super().__init__(args)

    def forward(self, *inputs, **kw):
        return self._forward(*inputs, step=self.step)

    def adjoint(self, *inputs, **kw):
        return self._adjoint(*inputs, step=self.step)

    def norm_bound(self, input_mags):
        return NotImplemented if self._norm_bound is None else self._norm_bound * input_mags[0]

    def is_gram_diag(self, freq=False):
        return self._diag is not None

    def get_diag(self, x, freq=False):
        return self._diag(x, self.step)


def LinOpFactory(forward, adjoint, diag=None, norm_bound=None):
    """Returns a constructor of custom operators (linop/blackbox.py:4-22)."""
    def make(*args):
        return BlackBox(*args, forward=forward, adjoint=adjoint, diag=diag, norm_bound=norm_bound)
    return make


# ------------------------------------------------------------------------------------------------
#  Tree evaluation (replaces the reference's Edge/CompGraph interpreter, linop/comp_graph.py:16-387)
# ------------------------------------------------------------------------------------------------

def evaluate(node: LinOp, env: dict, zero_constants: bool, memo: Optional[dict] = None):
    """Forward evaluation of the tree with variables bound by `env` {Variable: tensor}."""
    memo = {} if memo is None else memo
    if id(node) in memo:
        return memo[id(node)]
    if isinstance(node, Variable):
        out = env[node]
    elif isinstance(node, Constant):
        out = None if zero_constants else node.value
        dev = next((t.device for t in env.values() if isinstance(t, torch.Tensor)), None)
        if out is not None and dev is not None and out.device != dev:
            out = out.to(dev)                            # constants follow the variable's device
    else:
        ins = [evaluate(c, env, zero_constants, memo) for c in node.input_nodes]
        if isinstance(node, (sum,)):
            ins = [i for i in ins if i is not None]
            out = node.forward(*ins) if ins else None
        elif any(i is None for i in ins):
            out = None                                   # op applied to a zeroed constant stays zero
        else:
            out = node.forward(*ins)
    memo[id(node)] = out
    return out


def evaluate_adjoint(node: LinOp, y, grads: dict):
    """Adjoint evaluation: accumulates K^T y into `grads` {Variable: tensor}; constants absorb nothing."""
    if y is None or isinstance(node, Constant):
        return
    if isinstance(node, Variable):
        grads[node] = y if node not in grads else ops.axpby(1.0, grads[node], 1.0, y)
        return
    if isinstance(node, vstack):
        ys = list(y) if isinstance(y, (list, tuple)) else [y]
        for c, yi in zip(node.input_nodes, ys):
            evaluate_adjoint(c, yi, grads)
        return
    if isinstance(node, sum):
        for c in node.input_nodes:
            evaluate_adjoint(c, y, grads)
        return
    back = node.adjoint(y)
    if isinstance(back, LinOp.MultOutput):
        for c, b in zip(node.input_nodes, back):
            evaluate_adjoint(c, b, grads)
    else:
        evaluate_adjoint(node.input_nodes[0], back, grads)


class CompGraph:
    """Facade with the reference's CompGraph API (forward/adjoint/update_vars/sanity_check)."""

    instanceCnt = 0

    def __init__(self, end: LinOp, zero_out_constant: bool = False):
        self.instanceID = CompGraph.instanceCnt
        CompGraph.instanceCnt += 1
        self.end = end
        self.zero_out_constant = zero_out_constant

    @property
    def variables(self):
        return self.end.variables

    def forward(self, *values, return_list=False):
        vars_ = self.end.variables
        env = {v: val for v, val in zip(vars_, values)}
        y = evaluate(self.end, env, self.zero_out_constant)
        if return_list and y is not None and not isinstance(y, LinOp.MultOutput):
            y = [y]
        return y

    def adjoint(self, *values, return_list=False):
        grads = {}
        y = LinOp.MultOutput(values) if len(values) > 1 else values[0]
        evaluate_adjoint(self.end, y, grads)
        outs = [grads.get(v) for v in self.end.variables]
        out = LinOp.MultOutput(outs) if len(outs) > 1 else (outs[0] if outs else None)
        if return_list and out is not None and not isinstance(out, LinOp.MultOutput):
            out = [out]
        return out

    def update_vars(self, val):
        for i, var in enumerate(self.end.variables):
            if i < len(val):
                var.value = val[i]

    def sanity_check(self, eps=1e-5, shape=(1, 3, 64, 96), device="cuda", seed=0):
        """Dot-product (adjointness) test, comp_graph.py:342-371, on a seeded random input."""
        g = torch.Generator().manual_seed(seed)
        m = torch.rand(*shape, generator=g).to(device)
        d = self.forward(m)
        if isinstance(d, LinOp.MultOutput):
            d2 = [torch.rand(e.shape, generator=g).to(device) for e in d]
            m2 = self.adjoint(*d2)
            sd = _builtins.sum(float(ops.dot(a, b, per_sample=False)) for a, b in zip(d, d2))
        else:
            d2 = torch.rand(d.shape, generator=g).to(device)
            m2 = self.adjoint(d2)
            sd = float(ops.dot(d, d2, per_sample=False))
        sm = float(ops.dot(m, m2, per_sample=False))
        rel = abs((sm - sd) / sm)
        print(f"Sanity check {'passed' if rel < eps else 'failed'}, diff={abs(sm - sd)} rel_diff={rel}")
        return rel < eps

    def __str__(self):
        return self.__class__.__name__


def eval(linop, *inputs, zero_out_constant=True):  # noqa: A001
    return CompGraph(linop, zero_out_constant).forward(*inputs)


def adjoint(linop, *inputs, zero_out_constant=True):
    return CompGraph(linop, zero_out_constant).adjoint(*inputs)


def gram(linop, *inputs, zero_out_constant=True):
    K = CompGraph(linop, zero_out_constant)
    out = K.forward(*inputs)
    return K.adjoint(*out) if isinstance(out, LinOp.MultOutput) else K.adjoint(out)


def validate(linop: LinOp, **kw) -> bool:
    return CompGraph(linop).sanity_check(**kw)
