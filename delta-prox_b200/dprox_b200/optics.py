"""DOE optics forward model that feeds the unrolled solver in every training step.

Mirrors `dprox.contrib.optic` (common.py:27-164, doe_model.py:5-187): `RGBCollimator.get_psf()` (height map -> phase
profile -> aperture -> Fresnel propagation -> intensity -> area down-sampling -> normalisation), `img_psf_conv` (circular)
and `build_doe_model`.  Every stage is a native kernel with a native backward (`csrc/dpx_optics.cu`; complex transforms
through cuFFT C2C because sizes like 2244 = 1496 + 2*374 are general), chained by `torch.autograd.Function`s; torch itself
only builds the constant grids at construction time (cold) and moves data (pad / roll).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi as cabi
from . import ops


def _s(t):
    return cabi.stream_ptr(t.device)


def _cptr(t):
    return C.c_void_p(t.data_ptr())


def _c64(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"dprox_b200: {name} lives on {t.device}; this backend only computes on CUDA devices (no CPU fallback)")
    return t.to(torch.complex64).contiguous()


# ------------------------------------------------------------------------------------------------
#  building blocks: native forward + native backward
# ------------------------------------------------------------------------------------------------

def _c2c_raw(z, inverse):
    z = _c64(z, "z")
    H, W = z.shape[-2:]
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        cabi.check(cabi.lib().dpx_c2c(_cptr(z), _cptr(out), z.numel() // (H * W), H, W, int(inverse), _s(z)), "dpx_c2c")
    return out


class C2C(torch.autograd.Function):
    """unnormalised 2-D DFT over the last two axes; the adjoint of either direction is the other one."""

    @staticmethod
    def forward(ctx, z, inverse):
        ctx.inverse = inverse
        return _c2c_raw(z, inverse)

    @staticmethod
    def backward(ctx, g):
        return _c2c_raw(g, not ctx.inverse), None


def _cmul_raw(a, b, conj_b, scale):
    a, b = _c64(a, "a"), _c64(b, "b")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        cabi.check(cabi.lib().dpx_cmul(_cptr(a), _cptr(b), _cptr(out), a.numel(), b.numel(), int(conj_b), float(scale), _s(a)), "dpx_cmul")
    return out


class CMul(torch.autograd.Function):
    """out = scale * a * b with b broadcast over the leading (batch) axis of a; gradients for both factors."""

    @staticmethod
    def forward(ctx, a, b, scale):
        ctx.save_for_backward(a, b)
        ctx.scale = scale
        return _cmul_raw(a, b, False, scale)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = _c64(g, "grad")
        ga = _cmul_raw(g, b, True, ctx.scale) if ctx.needs_input_grad[0] else None
        gb = None
        if ctx.needs_input_grad[1]:
            gb = torch.empty_like(_c64(b, "b"))
            a_ = _c64(a, "a")
            with torch.cuda.device(g.device):
                cabi.check(cabi.lib().dpx_cmul_reduce(_cptr(g), _cptr(a_), _cptr(gb), b.numel(), a.numel() // b.numel(), float(ctx.scale),
                                                      _s(g)), "dpx_cmul_reduce")
        return ga, gb, None


class ToComplex(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.to_complex(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ops.real_part(g)


class RealPart(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        return ops.real_part(z)

    @staticmethod
    def backward(ctx, g):
        return ops.to_complex(g.contiguous())


class PhaseField(torch.autograd.Function):
    """field[l] = aperture * exp(i coef[l] h^2), zero-padded (doe_model.py:37-51, 103-104; common.py:156-157)."""

    @staticmethod
    def forward(ctx, h, coef, aperture, pad):
        h = cabi.require_cuda_f32(h, "height_map_sqrt")
        N, L = h.shape[-1], coef.numel()
        out = torch.empty(1, L, N + 2 * pad, N + 2 * pad, device=h.device, dtype=torch.complex64)
        with torch.cuda.device(h.device):
            cabi.check(cabi.lib().dpx_doe_field(cabi.ptr(h), cabi.ptr(coef), cabi.ptr(aperture), _cptr(out), L, N, pad, _s(h)), "dpx_doe_field")
        ctx.save_for_backward(h, coef, aperture)
        ctx.pad = pad
        return out

    @staticmethod
    def backward(ctx, g):
        h, coef, aperture = ctx.saved_tensors
        g = _c64(g, "grad")
        gh = torch.empty_like(h)
        with torch.cuda.device(h.device):
            cabi.check(cabi.lib().dpx_doe_field_backward(cabi.ptr(h), cabi.ptr(coef), cabi.ptr(aperture), _cptr(g), cabi.ptr(gh), coef.numel(),
                                                         h.shape[-1], ctx.pad, _s(h)), "dpx_doe_field_backward")
        return gh, None, None, None


class Abs2Pool(torch.autograd.Function):
    """crop the padding, |.|^2, average over factor x factor blocks (doe_model.py:106-107, common.py:27-44)."""

    @staticmethod
    def forward(ctx, fld, n, pad, factor):
        fld = _c64(fld, "field")
        L = fld.shape[1]
        out = torch.empty(1, L, n // factor, n // factor, device=fld.device, dtype=torch.float32)
        with torch.cuda.device(fld.device):
            cabi.check(cabi.lib().dpx_abs2_pool(_cptr(fld), cabi.ptr(out), L, n, pad, factor, 1.0, _s(fld)), "dpx_abs2_pool")
        ctx.save_for_backward(fld)
        ctx.cfg = (n, pad, factor)
        return out

    @staticmethod
    def backward(ctx, g):
        (fld,) = ctx.saved_tensors
        n, pad, factor = ctx.cfg
        g = cabi.require_cuda_f32(g, "grad")
        gf = torch.empty_like(fld)
        with torch.cuda.device(fld.device):
            cabi.check(cabi.lib().dpx_abs2_pool_backward(_cptr(fld), cabi.ptr(g), _cptr(gf), fld.shape[1], n, pad, factor, 1.0, _s(fld)),
                       "dpx_abs2_pool_backward")
        return gf, None, None, None


class NormalizeSum(torch.autograd.Function):
    """x / x.sum()  (doe_model.py:109)."""

    @staticmethod
    def forward(ctx, x):
        x = cabi.require_cuda_f32(x, "x")
        out, s = torch.empty_like(x), torch.empty(1, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            cabi.check(cabi.lib().dpx_normalize_sum(cabi.ptr(x), cabi.ptr(out), cabi.ptr(s), x.numel(), _s(x)), "dpx_normalize_sum")
        ctx.save_for_backward(out, s)
        return out

    @staticmethod
    def backward(ctx, g):
        out, s = ctx.saved_tensors
        g = cabi.require_cuda_f32(g, "grad")
        gx, tmp = torch.empty_like(out), torch.empty(1, device=out.device, dtype=torch.float32)
        with torch.cuda.device(out.device):
            cabi.check(cabi.lib().dpx_normalize_sum_backward(cabi.ptr(g), cabi.ptr(out), cabi.ptr(s), cabi.ptr(tmp), cabi.ptr(gx), out.numel(),
                                                             _s(out)), "dpx_normalize_sum_backward")
        return gx


# ------------------------------------------------------------------------------------------------
#  contrib/optic/common.py
# ------------------------------------------------------------------------------------------------

def get_coordinate(nx, ny, dx, dy):
    """common.py:11-24."""
    x = (torch.arange(nx) - (nx - 1.0) / 2) * dx
    y = (torch.arange(ny) - (ny - 1.0) / 2) * dy
    return torch.meshgrid(x, y, indexing="ij")


def _padded_kernel(psf: torch.Tensor, out_shape) -> torch.Tensor:
    """pad + all-axes ifftshift of psf2otf (common.py:47-82), incl. its even-pad off-by-one: pure data movement."""
    fh = psf.shape[2]
    if out_shape[2] != fh:
        pad = (out_shape[2] - fh) / 2
        if (out_shape[2] - fh) % 2 != 0:
            lo, hi = int(np.ceil(pad)), int(np.floor(pad))
        else:
            lo, hi = int(pad) + 1, int(pad) - 1
        psf = F.pad(psf, [lo, hi, lo, hi])
    return torch.fft.ifftshift(psf)


def img_psf_conv(img: torch.Tensor, psf: torch.Tensor, circular: bool = True) -> torch.Tensor:
    """Data formation `ifft2(fft2(img) * otf).real` (common.py:85-118), differentiable w.r.t. the image and the PSF."""
    img = cabi.require_cuda_f32(img, "img")
    crop = None
    if not circular:                                     # linear convolution: zero-pad to twice the size, crop (common.py:97-117)
        from . import ops
        from .linop import linear_conv_pads
        H, W = img.shape[-2:]
        pt, pb, pl, pr = linear_conv_pads(H, W)
        img = ops.pad2d(img, (H + pt + pb, W + pl + pr), pt, pl)
        crop = (H, W, pt, pl)
    kern = _padded_kernel(psf.to(img.device, torch.float32), img.shape)
    n = img.shape[-2] * img.shape[-1]
    X = C2C.apply(ToComplex.apply(img), False)
    K = C2C.apply(ToComplex.apply(kern), False)
    Y = C2C.apply(CMul.apply(X, K, 1.0 / n), True)
    out = RealPart.apply(Y)
    if crop is not None:
        from . import ops
        out = ops.pad2d(out, crop[:2], -crop[2], -crop[3])
    return out


class FresnelPropagator(nn.Module):
    """common.py:121-164: zero-pad by a quarter, multiply the spectrum with the Fresnel transfer function, crop."""

    def __init__(self, input_shape, distance, discretization_size, wave_lengths):
        super().__init__()
        _, _, M_orig, N_orig = input_shape
        self.Mpad, self.Npad = M_orig // 4, N_orig // 4
        M, N = M_orig + 2 * self.Mpad, N_orig + 2 * self.Npad
        xx, yy = get_coordinate(M, N, 1, 1)
        fx = torch.fft.ifftshift(xx / (discretization_size * N))
        fy = torch.fft.ifftshift(yy / (discretization_size * M))
        sq = (fx ** 2 + fy ** 2)[None][None]
        phi = -torch.pi * distance * wave_lengths.view(1, -1, 1, 1) * sq
        self.register_buffer("H", torch.exp(1j * phi).to(torch.complex64), persistent=False)

    def forward(self, padded_field):
        """`padded_field`: the already zero-padded wavefront [1,L,M,M]; returns the propagated field, still padded (the crop is
        fused into the intensity kernel)."""
        M = padded_field.shape[-1]
        return C2C.apply(CMul.apply(C2C.apply(padded_field, False), self.H, 1.0 / (M * M)), True)


class HeightMap(nn.Module):
    """doe_model.py:5-69."""

    def __init__(self, wave_lengths, refractive_idcs, xx, yy, sensor_distance):
        super().__init__()
        self.wave_lengths, self.refractive_idcs = wave_lengths, refractive_idcs
        self.register_buffer("coef", ((2.0 * torch.pi / wave_lengths) * (refractive_idcs - 1.0)).float(), persistent=False)
        k = 2 * torch.pi / wave_lengths[1]
        phase = (-k * ((xx ** 2 + yy ** 2)[None][None] / (2 * sensor_distance))) % (torch.pi * 2)
        height = (phase % (2 * torch.pi)) / k / (refractive_idcs[1] - 1.0)
        self.height_map_sqrt = nn.Parameter((height ** 0.5).float())


class RGBCollimator(nn.Module):
    """doe_model.py:72-153."""

    def __init__(self, sensor_distance, refractive_idcs, wave_lengths, patch_size, sample_interval, wave_resolution):
        super().__init__()
        self.wave_res, self.patch_size = wave_resolution, patch_size
        if wave_resolution[0] != wave_resolution[1] or wave_resolution[0] % patch_size:
            raise NotImplementedError("square wavefront with an integer down-sampling factor expected (common.py:40-43)")
        xx, yy = get_coordinate(wave_resolution[0], wave_resolution[1], sample_interval, sample_interval)
        r = torch.sqrt(xx ** 2 + yy ** 2)
        self.register_buffer("aperture", (r < xx.max()).float().contiguous(), persistent=False)
        self.height_map = HeightMap(wave_lengths, refractive_idcs, xx, yy, sensor_distance)
        self.propagator = FresnelPropagator((1, len(wave_lengths), *wave_resolution), sensor_distance, sample_interval, wave_lengths)

    def get_psf(self):
        """doe_model.py:91-110 as five native stages (+ their native backward)."""
        h = self.height_map.height_map_sqrt
        N = self.wave_res[0]
        fld = PhaseField.apply(h.reshape(N, N), self.height_map.coef, self.aperture, self.propagator.Mpad)
        fld = self.propagator(fld)
        inten = Abs2Pool.apply(fld, N, self.propagator.Mpad, N // self.patch_size)
        return NormalizeSum.apply(inten)

    def forward(self, input_img, circular=True):
        psfs = self.get_psf()
        return img_psf_conv(input_img, psfs, circular=circular), psfs


@dataclass
class DOEModelConfig:
    """doe_model.py:156-168."""
    circular: bool = True
    aperture_diameter: float = 3e-3
    sensor_distance: float = 15e-3
    refractive_idcs: torch.Tensor = field(default_factory=lambda: torch.tensor([1.4648, 1.4599, 1.4568]))
    wave_lengths: torch.Tensor = field(default_factory=lambda: torch.tensor([460, 550, 640]) * 1e-9)
    patch_size: int = 748
    sample_interval: float = 2e-6
    wave_resolution: tuple = (1496, 1496)


def build_doe_model(config: DOEModelConfig = DOEModelConfig()) -> RGBCollimator:
    """doe_model.py:171-187."""
    return RGBCollimator(config.sensor_distance, refractive_idcs=config.refractive_idcs, wave_lengths=config.wave_lengths,
                         patch_size=config.patch_size, sample_interval=config.sample_interval, wave_resolution=config.wave_resolution)
