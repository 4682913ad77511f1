"""TEST INFRASTRUCTURE — CPU restatement of the reference's proximal-iteration hot path.

This file is the *oracle*: a plain, single-file, CPU-only (torch-on-CPU + numpy) restatement of
what princeton-computational-imaging/Delta-Prox computes on the path named by
BASELINE.json:north_star.  It is NOT product code and is never
imported by the `dprox_b200` package: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` import it, and only as the checker.

Parity pin: every algorithm here is checked against outputs of the *unmodified* reference run in
the authoring container (`oracle/make_golden.py` -> `tests/golden/*.npz`,
`tests/test_oracle_golden.py`), plus the reference's own known-answer tests
(tests/problem/test_ml_problems.py:5-44, tests/linalg/test_linear_solver.py:57-111).

It deliberately keeps the reference's *cost profile* as well as its arithmetic (complex-to-complex
FFTs, iteration-invariant terms recomputed every iteration, the prox wrapper chain), because
`bench.py` times it as the CPU baseline (`cpu_baseline.kind == "port"`).

Every function cites the reference file:line it restates (paths relative to /root/reference).
A `dtype=torch.float64` run of the same code is the arbiter used for >50-iteration comparisons.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------------
#  OTF construction                                                    dprox/utils/psf2otf.py:11-98
# --------------------------------------------------------------------------------------------


def psf2otf(psf: np.ndarray, outsize: Sequence[int]) -> np.ndarray:
    """numpy OTF of a small kernel, MATLAB-psf2otf style (utils/psf2otf.py:11-40).

    The kernel is zero-padded *after* its last sample on every axis to `outsize`, rolled so that
    its centre tap (index floor(k/2)) lands on element 0, and transformed over ALL axes (including
    the channel axis of an [H, W, C] array - which is what broadcasts a (k,k,1) kernel to C
    channels).  The imaginary part is dropped when it is within round-off (`real_if_close`).
    dtype follows numpy: float32 kernel -> complex64/float32 under numpy>=2.
    """
    psf = np.asarray(psf)
    outsize = np.asarray(outsize)
    while psf.ndim < len(outsize):                       # _process_args :43-55
        psf = psf[..., None]
    ksize = np.asarray(psf.shape)
    if np.any(ksize > outsize):
        raise ValueError("outsize smaller than kernel")
    if np.all(psf == 0):
        return np.zeros(outsize)
    pad = [(0, int(o - k)) for o, k in zip(outsize, ksize)]
    big = np.pad(psf, pad, mode="constant")              # padarray 'post' :58-83
    for ax, k in enumerate(ksize):                       # circshift :86-98
        big = np.roll(big, -int(k // 2), axis=ax)
    otf = np.fft.fftn(big)
    n_ops = np.sum(big.size * np.log2(big.shape))
    return np.real_if_close(otf, tol=n_ops)


def otf_bchw(kernel: np.ndarray, shape) -> Tensor:
    """conv._FB (linop/conv.py:23-29): OTF as a [1,C,H,W] tensor (HWC->CHW iff C in {1,3})."""
    _, C, H, W = shape
    fb = torch.from_numpy(np.ascontiguousarray(psf2otf(kernel, [H, W, C])))
    if fb.ndim == 3 and fb.shape[2] in (1, 3):           # utils/misc.py:42-59 batchify
        fb = fb.permute(2, 0, 1)
    return fb.unsqueeze(0)


def psf2otf2(psf: Tensor, out_shape) -> Tensor:
    """conv_doe's torch OTF (linop/conv.py:59-80): zero-pad around the PSF (with the reference's
    off-by-one for even pads), `ifftshift` over ALL dims (rolls channels too), `fft2`."""
    fh = psf.shape[2]
    if out_shape[2] != fh:
        pad = (out_shape[2] - fh) / 2
        if (out_shape[2] - fh) % 2 != 0:
            lo, hi = int(math.ceil(pad)), int(math.floor(pad))
        else:
            lo, hi = int(pad) + 1, int(pad) - 1
        psf = F.pad(psf, [lo, hi, lo, hi])
    return torch.fft.fft2(torch.fft.ifftshift(psf))


def grad_kernel(dim: int) -> np.ndarray:
    """grad.__init__ (linop/grad.py:14-23): kernel [1,-1] laid along axis `dim` of an HWC kernel."""
    if dim not in (0, 1, 2):
        raise ValueError("dim must be 0(Height) or 1(Width) or 2 (Channel)")
    # dp.tensor([1, -1]) is an *int64* tensor, so numpy's fftn yields a complex128 OTF and every grad
    # forward/adjoint (and any diagonal it is added to) silently runs in double before `.float()`.
    k = np.array([1, -1], dtype=np.int64).reshape(1, 1, 2)
    return np.swapaxes(k, dim, -1)


def fspecial_gaussian(hsize: int, sigma: float) -> np.ndarray:
    """contrib/restoration.py:34-45 (normalised Gaussian, float64)."""
    r = (hsize - 1.0) / 2.0
    ax = np.arange(-r, r + 1)
    xx, yy = np.meshgrid(ax, ax)
    h = np.exp(-(xx * xx + yy * yy) / (2.0 * sigma * sigma))
    h[h < np.finfo(float).eps * h.max()] = 0
    s = h.sum()
    return h / s if s != 0 else h


def point_spread_function(ksize: int, sigma: float) -> np.ndarray:
    """contrib/restoration.py:21-22 -> (k,k,1) float32."""
    return fspecial_gaussian(ksize, sigma)[..., None].astype("float32")


def blurring(img_bchw: Tensor, psf_hw1: np.ndarray) -> Tensor:
    """contrib/restoration.py:25-31: scipy circular ('wrap') convolution of an HWC image."""
    import scipy.ndimage
    outs = []
    for im in img_bchw:
        hwc = im.permute(1, 2, 0).numpy()
        outs.append(torch.from_numpy(scipy.ndimage.convolve(hwc, psf_hw1, mode="wrap")).permute(2, 0, 1))
    return torch.stack(outs)


def log_descent(upper, lower, iter=24, sigma=0.255 / 255, w=1.0, lam=0.23, sqrt=False):
    """algo/tune/dpir.py:13-39."""
    s_log = np.logspace(np.log10(upper), np.log10(lower), iter).astype(np.float32)
    s_lin = np.linspace(upper, lower, iter).astype(np.float32)
    sigmas = (s_log * w + s_lin * (1 - w)) / 255.0
    rhos = [lam * (sigma ** 2) / (s ** 2) for s in sigmas]
    if not sqrt:
        sigmas = list(sigmas ** 2)
    return torch.tensor(np.array(rhos)).float(), torch.tensor(np.array(sigmas)).float()


# --------------------------------------------------------------------------------------------
#  Linear operators (single-variable chains)                     dprox/linop/*.py
# --------------------------------------------------------------------------------------------


class Op:
    """x -> A x for one node applied on top of `inner` (the reference's `input_nodes[0]`)."""
    inner: Optional["Op"] = None

    def fwd(self, x: Tensor) -> Tensor:                  # forward through the whole chain
        x = self.inner.fwd(x) if self.inner is not None else x
        return self._fwd(x)

    def adj(self, y: Tensor) -> Tensor:                  # adjoint through the whole chain
        y = self._adj(y)
        return self.inner.adj(y) if self.inner is not None else y

    def _fwd(self, x):
        return x

    def _adj(self, y):
        return y

    def freq_diag_ok(self) -> bool:                      # is_gram_diag(freq=True)
        return False

    def spatial_diag_ok(self) -> bool:                   # is_gram_diag(freq=False)
        return False

    def diag(self, ref: Tensor, freq: bool) -> Tensor:
        raise NotImplementedError


class Identity(Op):
    """Variable (linop/variable.py:8-100)."""

    def freq_diag_ok(self):
        return True

    def spatial_diag_ok(self):
        return True

    def diag(self, ref, freq):
        return torch.ones(ref.shape)                     # variable.py:47-59


class Scale(Op):
    """scale (linop/scale.py:7-80)."""

    def __init__(self, s: float, inner: Op):
        self.s, self.inner = s, inner

    def _fwd(self, x):
        return x * self.s

    _adj = _fwd

    def freq_diag_ok(self):
        return self.inner.freq_diag_ok()

    def spatial_diag_ok(self):
        return self.inner.spatial_diag_ok()

    def diag(self, ref, freq):
        d = self.inner.diag(ref, freq) * self.s          # scale.py:43-57
        return d * torch.conj(d)


class Conv(Op):
    """conv (linop/conv.py:15-56): circular convolution through full c2c FFTs."""

    def __init__(self, kernel, inner: Op):
        # to_ndarray (utils/misc.py:127-148): tensors keep their dtype, ndarrays are cast to float32
        if isinstance(kernel, Tensor):
            k = kernel.detach().cpu().numpy()
        elif isinstance(kernel, np.ndarray) and kernel.dtype != np.int64:
            k = kernel.astype("float32")
        else:
            k = np.asarray(kernel)
        self.kernel, self.inner, self._cache = k, inner, {}

    def FB(self, shape):
        shape = tuple(shape)
        if shape not in self._cache:
            self._cache[shape] = otf_bchw(self.kernel, shape)
        return self._cache[shape]

    def _fwd(self, x):
        Fx = torch.fft.fftn(x, dim=[-2, -1])
        return torch.real(torch.fft.ifftn(self.FB(x.shape) * Fx, dim=[-2, -1])).to(x.dtype)

    def _adj(self, y):
        Fy = torch.fft.fftn(y, dim=[-2, -1])
        return torch.real(torch.fft.ifftn(torch.conj(self.FB(y.shape)) * Fy, dim=[-2, -1])).to(y.dtype)

    def freq_diag_ok(self):
        return self.inner.freq_diag_ok()                 # conv.py:43-44

    def diag(self, ref, freq):
        assert freq
        fb = self.FB(ref.shape)
        return torch.abs(torch.conj(fb) * fb)            # conv.py:46-53


class Grad(Conv):
    """grad (linop/grad.py:8-23)."""

    def __init__(self, dim: int, inner: Op):
        super().__init__(grad_kernel(dim), inner)
        self.dim = dim


class Grad2D(Op):
    """[grad_H x ; grad_W x] stacked on the channel axis — NOT in the reference (it has no isotropic TV, SURVEY
    App. A-12).  Built from the reference's own `grad` so that the only new arithmetic is the group shrink below;
    parity for this operator/prox pair is therefore *unpinned* (checked against an fp64 run of the same formula)."""

    def __init__(self, inner: Op):
        self.inner, self.gh, self.gw = inner, Grad(0, Identity()), Grad(1, Identity())

    def _fwd(self, x):
        return torch.cat([self.gh._fwd(x), self.gw._fwd(x)], dim=1)

    def _adj(self, y):
        C = y.shape[1] // 2
        return self.gh._adj(y[:, :C]) + self.gw._adj(y[:, C:])

    def freq_diag_ok(self):
        return self.inner.freq_diag_ok()

    def diag(self, ref, freq):
        assert freq
        return self.gh.diag(ref, True) + self.gw.diag(ref, True)


def prox_iso_tv(v, lam):
    """group soft-threshold over the (grad_H, grad_W) pair of every pixel/channel"""
    C = v.shape[1] // 2
    nrm = torch.sqrt(v[:, :C] ** 2 + v[:, C:] ** 2)
    f = torch.where(nrm > lam, 1 - lam / nrm.clamp_min(1e-30), torch.zeros_like(nrm))
    return torch.cat([f * v[:, :C], f * v[:, C:]], dim=1)


def _linear_pad(x):
    """zero padding of the `circular=False` mode (linop/conv.py:103-110): BOTH sides are padded towards 2 * H (the height),
    ceil before / floor after."""
    target = 2 * x.shape[2]
    hp, wp = (target - x.shape[2]) / 2, (target - x.shape[3]) / 2
    pt, pb = int(np.ceil(hp)), int(np.floor(hp))
    pl, pr = int(np.ceil(wp)), int(np.floor(wp))
    return torch.nn.functional.pad(x, [pl, pr, pt, pb]), (pt, pb, pl, pr)


class ConvDOE(Op):
    """conv_doe (linop/conv.py:83-156); OTF rebuilt on every call.  `circular=False` zero-pads to twice the size, convolves
    circularly there and crops (:100-121); its `get_diag` nevertheless stays the circular |OTF|^2 at the image size (:143-152)."""

    def __init__(self, psf: Tensor, inner: Op, circular: bool = True):
        self.psf, self.inner, self.circular = psf, inner, circular

    def _apply(self, x, conj):
        crop = None
        if not self.circular:
            x, crop = _linear_pad(x)
        otf = psf2otf2(self.psf, x.shape)
        if conj:
            otf = torch.conj(otf)
        out = torch.real(torch.fft.ifftn(otf * torch.fft.fftn(x, dim=[-2, -1]), dim=[-2, -1])).float()
        if crop is not None:
            pt, pb, pl, pr = crop
            out = out[:, :, pt:-pb, pl:-pr]
        return out

    def _fwd(self, x):
        return self._apply(x, False).to(x.dtype)

    def _adj(self, y):
        return self._apply(y, True).to(y.dtype)

    def freq_diag_ok(self):
        return self.inner.freq_diag_ok()

    def diag(self, ref, freq):
        assert freq
        otf = psf2otf2(self.psf, ref.shape)
        return torch.abs(torch.conj(otf) * otf)


def img_psf_conv(img: Tensor, psf: Tensor, circular: bool = True) -> Tensor:
    """contrib/optic/common.py:85-118 (differentiable through torch's own FFT autograd, like the reference)."""
    crop = None
    if not circular:
        img, crop = _linear_pad(img)
    out = torch.fft.ifft2(torch.fft.fft2(img) * psf2otf2(psf, img.shape)).real
    if crop is not None:
        pt, pb, pl, pr = crop
        out = out[:, :, pt:-pb, pl:-pr]
    return out


def augment(img: Tensor, mode: int) -> Tensor:
    """Augment.augment (pnp/denoisers/composite.py:30-47): the 8 flips / rotations of the x8 test-time augmentation."""
    if mode == 0:
        return img
    if mode == 1:
        return img.rot90(1, [2, 3]).flip([2])
    if mode == 2:
        return img.flip([2])
    if mode == 3:
        return img.rot90(3, [2, 3])
    if mode == 4:
        return img.rot90(2, [2, 3]).flip([2])
    if mode == 5:
        return img.rot90(1, [2, 3])
    if mode == 6:
        return img.rot90(2, [2, 3])
    return img.rot90(3, [2, 3]).flip([2])


class Augment:
    """Augment (composite.py:6-28): one augmentation mode per call, cycling with the call counter."""

    def __init__(self, denoise: Callable):
        self.denoise, self.iter = denoise, 0

    def __call__(self, x, sigma):
        m = self.iter % 8
        y = self.denoise(augment(x, m), sigma)
        y = augment(y, 8 - m if m in (3, 5) else m)
        self.iter += 1
        return y


def bayer_mask(h: int, w: int) -> Tensor:
    """mosaic.masks_CFA_Bayer, RGGB (linop/subsample.py:34-48) -> [1,3,H,W] float32."""
    m = np.zeros((3, h, w), dtype="float32")
    m[0, 0::2, 0::2] = 1
    m[1, 0::2, 1::2] = 1
    m[1, 1::2, 0::2] = 1
    m[2, 1::2, 1::2] = 1
    return torch.from_numpy(m).unsqueeze(0)


class Mosaic(Op):
    """mosaic (linop/subsample.py:8-80)."""

    def __init__(self, inner: Op):
        self.inner = inner

    def _fwd(self, x):
        return bayer_mask(*x.shape[-2:]).to(x.dtype) * x

    _adj = _fwd

    def spatial_diag_ok(self):
        return self.inner.spatial_diag_ok()              # subsample.py:53-59

    def diag(self, ref, freq):
        assert not freq
        return bayer_mask(*ref.shape[-2:])


class Mul(Op):
    """mul_elementwise (linop/mul.py:44-73)."""

    def __init__(self, w: Tensor, inner: Op):
        self.w, self.inner = w, inner

    def _fwd(self, x):
        return self.w.to(x.dtype) * x

    _adj = _fwd

    def spatial_diag_ok(self):
        return self.inner.spatial_diag_ok()

    def diag(self, ref, freq):
        assert not freq
        return self.w


class BlackBox(Op):
    """LinOpFactory / BlackBox (linop/blackbox.py:4-78): user callables `(x, step=)`."""

    def __init__(self, forward: Callable, adjoint: Callable, inner: Op):
        self.f, self.a, self.inner, self.step = forward, adjoint, inner, 0

    def _fwd(self, x):
        return self.f(x, step=self.step)

    def _adj(self, y):
        return self.a(y, step=self.step)


# --------------------------------------------------------------------------------------------
#  Proximal operators                                   dprox/proxfn/{base,nonneg,norm,sum_square}.py
# --------------------------------------------------------------------------------------------


def prox_nonneg(v, lam):                                 # nonneg.py:10-11
    return torch.maximum(v, torch.zeros_like(v))


def prox_norm1(v, lam):                                  # norm.py:6-19
    return torch.sign(v) * torch.maximum(torch.abs(v) - lam, torch.zeros_like(v))


def prox_norm2(v, lam):                                  # norm.py:22-27, sum_square.py:26-27
    return v / (1 + 2 * lam)


def safe_sqrt(x, eps=1e-8):                              # utils/misc.py:151-161
    return torch.sqrt(torch.clamp(x, min=eps))


@dataclass(eq=False)
class Term:
    """One proxable function f(A x - c) of the objective.

    kind : 'sum_squares' | 'nonneg' | 'norm1' | 'norm2' | 'deep_prior' | 'custom'
    op   : linear chain A;  c : constant subtracted inside the linop (`A(x) - c`) or None
    b    : sum_squares' explicit second argument (`sum_squares(A(x), b)`) or None
    alpha: set by `c * fn` (proxfn/base.py:78-82);  beta: always 1 in the reference
    """
    kind: str
    op: Op = field(default_factory=Identity)
    c: Optional[Tensor] = None
    b: Optional[Tensor] = None
    alpha: float = 1.0
    beta: float = 1.0
    denoiser: Optional[Callable] = None                  # deep_prior: (v, sigma[B,1,1,1]) -> v
    sqrt: bool = False
    clamp: bool = False
    custom_prox: Optional[Callable] = None
    box: tuple = (0.0, 1.0)

    # -- the affine linop and its constant part ------------------------------------------
    def K(self, x):
        """Affine forward `A x - c`  (CompGraph with zero_out_constant=False)."""
        y = self.op.fwd(x)
        return y - self.c.to(y.dtype) if self.c is not None else y

    def linop_offset(self, like: Tensor):
        """LinOp.offset (linop/base.py:118-129): the DAG evaluated at x = 0 -> `A 0 - c`.
        The reference really runs the forward on zeros every call; so do we (cost profile)."""
        return self.K(torch.zeros_like(like))

    def offset(self, like: Tensor):
        """ProxFn.offset = -linop.offset (proxfn/base.py:43-45); sum_squares with an explicit
        `b` returns b instead (sum_square.py:19-23)."""
        if self.kind == "sum_squares" and self.b is not None:
            return self.b if self.b.is_complex() else self.b.to(like.dtype)      # complex k-space data stay complex
        return -self.linop_offset(like)

    # -- prox with the wrapper chain -------------------------------------------------------
    def _prox(self, v, lam):
        if self.kind == "nonneg":
            return prox_nonneg(v, lam)
        if self.kind == "norm1":
            return prox_norm1(v, lam)
        if self.kind in ("norm2", "sum_squares"):
            return prox_norm2(v, lam)
        if self.kind == "deep_prior":                    # pnp/prior.py:73-86
            sigma = safe_sqrt(lam) if self.sqrt else lam
            if self.clamp:
                v = v.clamp(0, 1)
            out = self.denoiser(v, sigma.reshape(-1, 1, 1, 1))
            return out.to(v.dtype).reshape(v.shape)
        if self.kind == "iso_tv":
            return prox_iso_tv(v, lam)
        if self.kind == "box":                           # NOT in the reference (north star): projection onto [lo, hi]
            return torch.clamp(v, self.box[0], self.box[1])
        if self.kind == "custom":
            return self.custom_prox(v, lam)
        raise ValueError(self.kind)

    def prox(self, v, lam):
        """ProxFn.prox (proxfn/base.py:55-64):
        translated(affine(scaled(_prox, alpha), beta), offset)."""
        if lam.ndim == 1:
            lam = lam.view(lam.shape[0], 1, 1, 1)
        # (the stacked-gradient operator changes the channel count, so its - always zero - offset cannot be
        #  evaluated on v's shape the way the reference does for shape-preserving linops)
        off = self.offset(v) if self.kind != "iso_tv" else torch.zeros_like(v)
        w = self.beta * (v - off)
        return 1.0 / self.beta * self._prox(w, self.beta * self.beta * lam * self.alpha) + off

    def grad(self, x):
        """sum_squares.grad (sum_square.py:29-32): A^T(A x - offset)."""
        return self.op.adj(self.op.fwd(x) - self.offset(x))


# --------------------------------------------------------------------------------------------
#  Linear solvers                                          dprox/linalg/solve/solver_cg.py
# --------------------------------------------------------------------------------------------


def _ravel(x):
    return x if x.ndim == 1 else x.reshape(x.shape[0], -1)


def _bdot(x, y):                                         # solver_cg.py:7-22
    if x.ndim == 1:
        return torch.dot(x, y)
    return torch.sum(_ravel(x) * _ravel(y), dim=-1)


def _expand(x, ref):
    while x.ndim < ref.ndim:
        x = x.unsqueeze(-1)
    return x


def cg(A: Callable, b: Tensor, x0=None, rtol=1e-6, max_iters=100):
    """solver_cg.py:56-136, including its stop test: a *matrix* 2-norm of the [B,n] residual
    compared with per-sample tolerances."""
    x = torch.zeros_like(b) if x0 is None else x0
    r = b - A(x)
    tol = rtol * torch.linalg.norm(_ravel(b), 2, dim=-1)
    gamma_prev = p = None
    n_it = int(min(max_iters, int(np.prod(b.shape))))
    for it in range(n_it):
        normr = torch.linalg.norm(_ravel(r), 2)
        if torch.all(normr <= tol):
            break
        gamma = _expand(_bdot(r, r), x)
        p = r if it == 0 else r + (gamma / gamma_prev) * p
        q = A(p)
        alpha = gamma / _expand(_bdot(p, q), x)
        x = x + alpha * p
        r = r - alpha * q
        gamma_prev = gamma
    return x


def pcg(A: Callable, b: Tensor, x0=None, rtol=1e-6, max_iters=100, Minv: Optional[Callable] = None):
    """solver_cg.py:172-233: starts from ones, whole-tensor dot products, absolute inf-norm stop."""
    Minv = Minv or (lambda t: t)
    x = torch.ones_like(b) if x0 is None else x0
    r = A(x) - b
    y = Minv(r)
    p = -y
    for _ in range(max_iters):
        Ap = A(p)
        ry = r.ravel() @ y.ravel()
        alpha = ry / (p.ravel() @ Ap.ravel())
        x = x + alpha * p
        r = r + alpha * Ap
        y = Minv(r)
        beta = (r.ravel() @ y.ravel()) / ry
        p = -y + beta * p
        if torch.linalg.vector_norm(r.ravel(), ord=float("inf")) < rtol:
            break
    return x


LINEAR_SOLVERS = {"cg": cg, "pcg": pcg}


# --------------------------------------------------------------------------------------------
#  x-update                                                 dprox/proxfn/sum_square.py:87-197
# --------------------------------------------------------------------------------------------


def fft2c(x):                                            # utils/misc.py:164-177 (centred, ortho)
    return torch.fft.fftshift(torch.fft.fft2(torch.fft.ifftshift(x, dim=(-2, -1)), norm="ortho"), dim=(-2, -1))


def ifft2c(x):                                           # utils/misc.py:180-193
    return torch.fft.fftshift(torch.fft.ifft2(torch.fft.ifftshift(x, dim=(-2, -1)), norm="ortho"), dim=(-2, -1))


def csmri_prox(v, lam, num_psi, mask, y):
    """csmri._prox (proxfn/fast/csmri.py:14-25): closed-form x-update of |M F x - y|^2 + rho sum_i |x - b_i|^2 given
    v = sum_i b_i; only the sampled k-space locations change."""
    if lam.ndim == 1:
        lam = lam.view(lam.shape[0], 1, 1, 1)
    z = fft2c(v)
    temp = (lam * z + y) / (1 + lam * num_psi)
    return ifft2c(torch.where(mask.bool(), temp, z))


def custom_admm_csmri(denoise, mask, y, x0, rhos, sigmas):
    """CustomADMM (contrib/csmri.py:156-171) with one deep prior and the csmri data term: prox first, complex z / u.
    `denoise(v_real, sigma[B,1,1,1]) -> real`; deep_prior takes the real part of a complex input (pnp/prior.py:79).
    Returns the reference's state (x, z, u) after len(rhos) iterations (ADMM.initialize: x = x0, z = x0, u = 0)."""
    x, z, u = x0, x0, torch.zeros_like(x0)
    for it in range(len(rhos)):
        w = z - u
        x = denoise(w.real if torch.is_complex(w) else w, sigmas[it].reshape(-1, 1, 1, 1)).to(torch.float32)
        b = x + u
        z = csmri_prox(b, rhos[it], 1, mask, y)          # ext_sum_squares.solve: sum of b, num_psi = 1 (sum_square.py:44-48)
        u = u + x - z
    return x, z, u


class LeastSquares:
    """least_squares (sum_square.py:87-197): argmin_x sum_q |A_q x - b_q|^2 + rho sum_i |A_i x - b_i|^2."""

    def __init__(self, quad: List[Term], other: List[Term], try_diagonalize=True, try_freq_diagonalize=True,
                 solver_type="cg", rtol=1e-6, max_iters=100):
        self.quad, self.other = quad, other
        ops = [t.op for t in quad + other]
        self.diagonalizable = all(o.spatial_diag_ok() for o in ops) and try_diagonalize          # :106
        self.freq_diagonalizable = all(o.freq_diag_ok() for o in ops) and try_diagonalize and try_freq_diagonalize
        self.solver_type, self.rtol, self.max_iters = solver_type, rtol, max_iters

    def solve(self, b: List[Tensor], rho: Tensor, like: Tensor, eps=1e-7):
        """`like` = the current value of the variable (the reference zeroes `Variable.value` to
        evaluate offsets, linop/base.py:118-129, so it only supplies shape/dtype)."""
        if rho.ndim == 1:
            rho = rho.view(rho.shape[0], 1, 1, 1)        # :116-117
        if self.diagonalizable or self.freq_diagonalizable:
            return self._direct(b, rho, like, eps)
        return self._cg(b, rho, like)

    def _rhs(self, b, rho, like):
        Ktb = 0
        for t in self.quad:
            Ktb = Ktb + t.op.adj(t.offset(like))         # :127-132 (offset re-evaluated each call)
        for i, t in enumerate(self.other):
            Ktb = Ktb + rho * t.op.adj(b[i])             # :133-134
        return Ktb

    def _direct(self, b, rho, like, eps):
        Ktb = self._rhs(b, rho, like)
        freq = self.freq_diagonalizable
        diag = 0
        for t in self.quad:
            diag = diag + t.op.diag(Ktb, freq)           # :142-146
        for t in self.other:
            diag = diag + rho * t.op.diag(Ktb, freq)
        if freq:                                         # :150-152
            FK = torch.fft.fftn(Ktb, dim=[-2, -1])
            out = torch.real(torch.fft.ifftn((FK + eps) / (diag + eps), dim=[-2, -1]))
        else:                                            # :154
            out = Ktb / (diag + eps)
        return out.to(Ktb.dtype)

    def _cg(self, b, rho, like):

        def KtK(x):                                      # :160-173
            out = 0
            for t in self.quad:
                out = out + t.op.adj(t.op.fwd(x))
            for t in self.other:
                out = out + rho * t.op.adj(t.op.fwd(x))
            return out

        Ktb = self._rhs(b, rho, like)
        solve = lambda A, rhs: LINEAR_SOLVERS[self.solver_type](A, rhs, rtol=self.rtol, max_iters=self.max_iters)
        if torch.is_grad_enabled() and Ktb.requires_grad:
            return _ImplicitSolve.apply(KtK, Ktb, solve)
        return solve(KtK, Ktb)


class _ImplicitSolve(torch.autograd.Function):
    """LinearSolve (linalg/custom.py:39-62): x = solve(A, b); backward grad_b = solve(A^T, grad_x) with A^T = A for the
    normal equations.  The reference's KtK multiplies by the closure `rho` instead of its `self.rho` parameter
    (sum_square.py:160-173), so `autograd.grad(-A(x), params(A), grad_b)` finds no path to rho: the matrix' dependence
    on rho is NOT differentiated, only the right-hand side's -- restated as is."""

    @staticmethod
    def forward(ctx, A, b, solve):
        ctx.A, ctx.solve = A, solve
        with torch.no_grad():
            return solve(A, b)

    @staticmethod
    def backward(ctx, g):
        with torch.no_grad():
            return None, ctx.solve(ctx.A, g), None


# --------------------------------------------------------------------------------------------
#  Algorithms                                   dprox/algo/{base,admm,hqs,pgd}.py
# --------------------------------------------------------------------------------------------


def partition(terms: List[Term], method: str):
    """ADMM.partition (admm.py:26-36) / PGD.partition (pgd.py:9-26)."""
    if method == "pgd":
        if len(terms) != 2:
            raise ValueError("Proximal gradient descent only supports two proximal functions for now.")
        omega = [t for t in terms if t.kind == "sum_squares"]
        if not omega:
            raise ValueError("Proximal gradient descent requires at least one proximal function is differentiable.")
        return [t for t in terms if t not in omega], omega
    omega = [t for t in terms if t.kind == "sum_squares"]
    return [t for t in terms if t not in omega], omega


def _schedule(val, default, T, psi=None):
    """Algorithm.defaults (base.py:205-218)."""
    if val is None:
        val = default
    if np.isscalar(val) or (isinstance(val, Tensor) and val.ndim == 0):
        val = torch.tensor([float(val)] * T)
    return val


class Solver:
    """compile()+Algorithm (primitives.py:40-67, base.py:58-178) for a single-variable objective."""

    def __init__(self, terms: List[Term], method="admm", dtype=torch.float32, **ls_kwargs):
        self.method, self.dtype = method, dtype
        self.psi, self.omega = partition(terms, method)
        if method != "pgd":
            self.ls = LeastSquares(self.omega, self.psi, **ls_kwargs)

    # -- state ---------------------------------------------------------------------------------
    def initialize(self, x0):
        """ADMM.initialize admm.py:61-67; HQS.initialize hqs.py:5-8; PGD pgd.py:45-46."""
        x = x0.to(self.dtype)
        if self.method == "pgd":
            return [x]
        v = [t.K(x) for t in self.psi]
        if self.method == "hqs":
            return x, v
        if self.method == "pc":                          # pc.py:7-11
            return x, v, x.clone()
        return x, v, [torch.zeros_like(e) for e in v]

    # -- one iteration -------------------------------------------------------------------------
    def iter(self, state, rho, lam: Dict[Term, Tensor]):
        m = self.method
        if m == "admm":                                  # admm.py:49-59
            x, v, u = state
            b = [v[i] - u[i] for i in range(len(self.psi))]
            x = self.ls.solve(b, rho, x)
            for i, t in enumerate(self.psi):
                Kx = t.K(x)
                v[i] = t.prox(Kx + u[i], lam[t])
                u[i] = u[i] + Kx - v[i]
            return x, v, u
        if m == "hqs":                                   # hqs.py:10-16
            x, z = state
            x = self.ls.solve(z, rho, x)
            for i, t in enumerate(self.psi):
                z[i] = t.prox(t.K(x), lam[t])
            return x, z
        if m == "ladmm":                                 # admm.py:79-100 (per-sample intent, App. A-17)
            x, v, u = state
            b = []
            for i, t in enumerate(self.psi):
                tmp = t.op.adj(t.op.fwd(x) - v[i] + u[i])
                b.append(x - tmp)
            x = self.ls.solve(b, rho, x)
            for i, t in enumerate(self.psi):
                Kx = t.K(x)
                v[i] = t.prox(Kx + u[i], lam[t])
                u[i] = u[i] + Kx - v[i]
            return x, v, u
        if m == "admm_vxu":                              # admm.py:107-120 (per-sample intent, App. A-17)
            z, x, u = state
            for i, t in enumerate(self.psi):
                x[i] = t.prox(t.K(z) - u[i], lam[t])
            b = [x[i] + u[i] for i in range(len(self.psi))]
            z = self.ls.solve(b, rho, z)
            for i in range(len(self.psi)):
                u[i] = u[i] + x[i] - z
            return z, x, u
        if m == "pc":                                    # PockChambolle._iter, pc.py:13-36
            x, z, xbar = state
            for i, t in enumerate(self.psi):
                r = lam[t].view(lam[t].shape[0], 1, 1, 1) if lam[t].ndim == 1 else lam[t]
                z[i] = z[i] + r * t.K(xbar)
                z[i] = z[i] - r * t.prox(z[i], r)
            xn = [x - t.op.adj(z[i]) for i, t in enumerate(self.psi)]
            x_next = self.ls.solve(xn, rho, x) if self.omega else sum(xn)
            return x_next, z, x_next + x_next - x
        if m == "pgd":                                   # pgd.py:39-43
            x = state[0]
            r = rho.view(rho.shape[0], 1, 1, 1) if rho.ndim == 1 else rho
            v = x - r * self.omega[0].grad(x)
            return [self.psi[0].prox(v, lam[self.psi[0]])]
        raise ValueError(m)

    # -- the loop ------------------------------------------------------------------------------
    def solve(self, x0, rhos=None, lams=None, max_iter=24, callback=None, return_full_states=False):
        """Algorithm.solve/iters (base.py:85-156): fixed `max_iter` iterations, rho=rhos[...,it]."""
        T = max_iter
        rhos = _schedule(rhos, 1.0, T).to(self.dtype)
        if lams is None or np.isscalar(lams) or (isinstance(lams, Tensor) and lams.ndim == 0):
            lams = {t: _schedule(lams, 0.02, T) for t in self.psi}
        lams = {t: _schedule(v, 0.02, T).to(self.dtype) for t, v in lams.items()}
        state = self.initialize(x0)
        for it in range(T):
            rho = rhos[..., it]
            lam = {t: v[..., it] for t, v in lams.items()}
            for t in self.psi + self.omega:
                _set_step(t.op, it)
            state = self.iter(state, rho, lam)
            if callback is not None:
                callback(iter=it, state=state, rho=rho, lam=lam)
        return state if return_full_states else state[0]


def _set_step(op, step):
    while op is not None:
        if hasattr(op, "step"):
            op.step = step
        op = op.inner


# --------------------------------------------------------------------------------------------
#  FFDNet-color forward (the deep_prior denoiser)   pnp/denoisers/models/network_ffdnet.py:27-68
# --------------------------------------------------------------------------------------------


def ffdnet_random_weights(seed: int, in_nc=3, nc=96, nb=12, dtype=torch.float32):
    """Seeded random weights with nn.Conv2d's default init, shapes as FFDNet(in_nc,in_nc,nc,nb)."""
    g = torch.Generator().manual_seed(seed)
    chans = [in_nc * 4 + 1] + [nc] * (nb - 1) + [in_nc * 4]
    ws = []
    for cin, cout in zip(chans[:-1], chans[1:]):
        bound = 1.0 / math.sqrt(cin * 9)
        w = (torch.rand(cout, cin, 3, 3, generator=g) * 2 - 1) * bound
        b = (torch.rand(cout, generator=g) * 2 - 1) * bound
        ws.append((w.to(dtype), b.to(dtype)))
    return ws


def ffdnet_forward(weights, x: Tensor, sigma: Tensor) -> Tensor:
    """FFDNet.forward (network_ffdnet.py:54-68): replicate-pad to even, PixelUnshuffle(2), append the
    sigma map, conv3x3+ReLU stack, conv3x3, PixelShuffle(2), crop."""
    h, w = x.shape[-2:]
    x = F.pad(x, (0, (-w) % 2, 0, (-h) % 2), mode="replicate")
    x = F.pixel_unshuffle(x, 2)
    m = torch.ones((x.shape[0], 1, x.shape[2], x.shape[3]), dtype=x.dtype) * sigma.reshape(-1, 1, 1, 1).to(x.dtype)
    x = torch.cat((x, m), 1)
    for i, (wt, bs) in enumerate(weights):
        x = F.conv2d(x, wt, bs, padding=1)
        if i < len(weights) - 1:
            x = F.relu(x)
    x = F.pixel_shuffle(x, 2)
    return x[..., :h, :w]
