"""Denoisers behind `deep_prior`.

`FFDNetColor` has the architecture and `state_dict` keys of the reference's FFDNet-color
(pnp/denoisers/models/network_ffdnet.py:27-68: PixelUnshuffle(2) -> conv3x3(13->96)+ReLU ->
10 x [conv3x3(96->96)+ReLU] -> conv3x3(96->12) -> PixelShuffle(2)), so the published
`ffdnet_color.pth` loads unchanged.  It is launched as an *external* prox between two native
stages on the same CUDA stream; its 3x3 convolutions currently run through the library conv
(cuDNN) — a hand-written tcgen05 implicit-GEMM path is the §8f follow-up.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .proxfn import Denoiser


class FFDNet(nn.Module):
    def __init__(self, in_nc=3, out_nc=3, nc=96, nb=12):
        super().__init__()
        layers = [nn.Conv2d(in_nc * 4 + 1, nc, 3, padding=1), nn.ReLU(inplace=True)]
        for _ in range(nb - 2):
            layers += [nn.Conv2d(nc, nc, 3, padding=1), nn.ReLU(inplace=True)]
        layers += [nn.Conv2d(nc, out_nc * 4, 3, padding=1)]
        self.model = nn.Sequential(*layers)          # keys model.{0,2,...,22}.{weight,bias}

    def forward(self, x, sigma):
        h, w = x.shape[-2:]
        x = F.pad(x, (0, (-w) % 2, 0, (-h) % 2), mode="replicate")
        x = F.pixel_unshuffle(x, 2)
        m = sigma.reshape(-1, 1, 1, 1).to(x.dtype).expand(x.shape[0], 1, x.shape[2], x.shape[3])
        x = self.model(torch.cat((x, m), 1))
        return F.pixel_shuffle(x, 2)[..., :h, :w]


class NativeFFDNet:
    """FFDNet-color forward on tcgen05 tensor cores through the C-ABI (`dpx_ffdnet_*`, csrc/dpx_ffdnet.cu): NHWC bf16
    activations, implicit-GEMM 3x3 convolutions with fp32 TMEM accumulation, fused bias+ReLU, fused
    unshuffle/sigma prologue and shuffle/crop epilogue.  bf16 operands: ~1e-2 relative to the fp32 network."""

    def __init__(self, model: "FFDNet", device):
        import ctypes as C
        from . import _cabi as cabi
        self._cabi, self.device = cabi, torch.device(device)
        lib = cabi.lib()
        if not lib.dpx_ffdnet_available():
            raise RuntimeError("libdprox_b200 was built without the tcgen05 convolution (CUTLASS headers missing)")
        convs = [m for m in model.model if isinstance(m, nn.Conv2d)]
        self._h = C.c_void_p()
        cabi.check(lib.dpx_ffdnet_create(len(convs), convs[1].out_channels, C.byref(self._h)), "dpx_ffdnet_create")
        with torch.cuda.device(self.device):
            for i, c in enumerate(convs):
                w = c.weight.detach().to(self.device, torch.float32).contiguous()
                b = c.bias.detach().to(self.device, torch.float32).contiguous()
                cabi.check(lib.dpx_ffdnet_set_layer(self._h, i, cabi.ptr(w), cabi.ptr(b), w.shape[0], w.shape[1],
                                                    cabi.stream_ptr(self.device)), "dpx_ffdnet_set_layer")
            torch.cuda.current_stream(self.device).synchronize()

    def __call__(self, x: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
        cabi = self._cabi
        x = cabi.require_cuda_f32(x, "x")
        B, Cc, H, W = x.shape
        if Cc != 3:
            raise ValueError("native FFDNet-color expects 3 channels")
        sigma = cabi.require_cuda_f32(sigma.to(x.device, torch.float32).reshape(-1), "sigma")
        if sigma.numel() not in (1, B):
            raise ValueError(f"sigma must have 1 or {B} entries")
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            cabi.check(cabi.lib().dpx_ffdnet_forward(self._h, cabi.ptr(x), cabi.ptr(sigma), int(sigma.numel() > 1), cabi.ptr(y),
                                                     B, H, W, cabi.stream_ptr(x.device)), "dpx_ffdnet_forward")
        return y

    def __del__(self):
        try:
            if self._h:
                self._cabi.lib().dpx_ffdnet_destroy(self._h)
        except Exception:
            pass


class FFDNetColorDenoiser(Denoiser):
    """pnp/denoisers/wrapper.py:38-48.  `precision='fp32'` (default) keeps fp32 convolutions for 1e-5-class parity;
    `precision='bf16'` runs the native tcgen05 network (NativeFFDNet) — the fast mode."""

    def __init__(self, model_path=None, seed=None, precision="fp32"):
        super().__init__()
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision, self._native = precision, None
        self.model = FFDNet(3, 3, 96, 12)
        if model_path is not None:
            self.model.load_state_dict(torch.load(model_path, map_location="cpu"), strict=True)
        elif seed is not None:
            self.load_seeded(seed)

    def load_seeded(self, seed: int):
        """Deterministic random weights (nn.Conv2d's default init bounds) — what the parity tests use, because the
        pretrained file needs a download (pnp/prior.py:14-35)."""
        g = torch.Generator().manual_seed(seed)
        convs = [m for m in self.model.model if isinstance(m, nn.Conv2d)]
        with torch.no_grad():
            for c in convs:
                bound = 1.0 / math.sqrt(c.in_channels * 9)
                c.weight.copy_((torch.rand(c.weight.shape, generator=g) * 2 - 1) * bound)
                c.bias.copy_((torch.rand(c.bias.shape, generator=g) * 2 - 1) * bound)
        return self

    def _denoise(self, x, sigma):
        wants_grad = torch.is_grad_enabled() and (x.requires_grad or sigma.requires_grad
                                                   or any(p.requires_grad for p in self.model.parameters()))
        if self.precision == "bf16" and wants_grad:
            # training (unrolled solver, BASELINE config 5): the tcgen05 forward has no backward yet, so the tape runs
            # through the framework's convolutions with bf16 operands / fp32 accumulation
            if not getattr(self, "_nhwc", False):          # NHWC weights: cuDNN's tensor-core kernels for forward, dgrad and wgrad
                self.model.to(memory_format=torch.channels_last)
                self._nhwc = True
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.model(x.contiguous(memory_format=torch.channels_last), sigma).float().contiguous()
        if self.precision == "bf16":
            if self._native is None or self._native.device != x.device:
                self._native = NativeFFDNet(self.model, x.device)
            return self._native(x, sigma)
        # fp32 convolutions (no TF32) so that results stay within 1e-5 of the fp32 CPU reference
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            return self.model(x, sigma)
        finally:
            torch.backends.cudnn.allow_tf32 = prev


# ------------------------------------------------------------------------------------------------
#  DRUNet (pnp/denoisers/models/network_unet.py:67-116, wrapper.py:89-146) — SURVEY §8f rank 4
# ------------------------------------------------------------------------------------------------

class _ResBlock(nn.Module):
    """x + conv(relu(conv(x))), bias-free 3x3 convolutions; state_dict keys `res.0.weight`, `res.2.weight`."""

    def __init__(self, c):
        super().__init__()
        self.res = nn.Sequential(nn.Conv2d(c, c, 3, padding=1, bias=False), nn.ReLU(inplace=True), nn.Conv2d(c, c, 3, padding=1, bias=False))

    def forward(self, x):
        return x + self.res(x)


class UNetRes(nn.Module):
    """DRUNet body: head conv, 3 x (nb residual blocks + 2x2 stride-2 conv), nb residual blocks, 3 x (2x2 transposed conv +
    nb residual blocks) with additive skips, tail conv.  Module names follow the reference so published weights load."""

    def __init__(self, in_nc=2, out_nc=1, nc=(64, 128, 256, 512), nb=4):
        super().__init__()
        blocks = lambda c: [_ResBlock(c) for _ in range(nb)]
        self.m_head = nn.Conv2d(in_nc, nc[0], 3, padding=1, bias=False)
        self.m_down1 = nn.Sequential(*blocks(nc[0]), nn.Conv2d(nc[0], nc[1], 2, stride=2, bias=False))
        self.m_down2 = nn.Sequential(*blocks(nc[1]), nn.Conv2d(nc[1], nc[2], 2, stride=2, bias=False))
        self.m_down3 = nn.Sequential(*blocks(nc[2]), nn.Conv2d(nc[2], nc[3], 2, stride=2, bias=False))
        self.m_body = nn.Sequential(*blocks(nc[3]))
        self.m_up3 = nn.Sequential(nn.ConvTranspose2d(nc[3], nc[2], 2, stride=2, bias=False), *blocks(nc[2]))
        self.m_up2 = nn.Sequential(nn.ConvTranspose2d(nc[2], nc[1], 2, stride=2, bias=False), *blocks(nc[1]))
        self.m_up1 = nn.Sequential(nn.ConvTranspose2d(nc[1], nc[0], 2, stride=2, bias=False), *blocks(nc[0]))
        self.m_tail = nn.Conv2d(nc[0], out_nc, 3, padding=1, bias=False)

    def forward(self, x0):
        x1 = self.m_head(x0)
        x2 = self.m_down1(x1)
        x3 = self.m_down2(x2)
        x4 = self.m_down3(x3)
        x = self.m_body(x4)
        x = self.m_up3(x + x4)
        x = self.m_up2(x + x3)
        x = self.m_up1(x + x2)
        return self.m_tail(x + x1)


class DRUNetDenoiser(Denoiser):
    """DRUNet behind `deep_prior(x, denoiser='drunet' | 'drunet_color')` (wrapper.py:89-146): the noise level enters as an
    extra input channel; images larger than 256 x 256 are denoised as four overlapping quadrants (recursively), images up
    to that size are replicate-padded to a multiple of 16.  An opaque torch module on the prox hook, like every denoiser
    but FFDNet-color."""

    REFIELD, MIN_SIZE, MODULO = 32, 256, 16

    def __init__(self, n_channels=1, model_path=None):
        super().__init__()
        self.model = UNetRes(in_nc=n_channels + 1, out_nc=n_channels)
        if model_path is not None:
            self.model.load_state_dict(torch.load(model_path, map_location="cpu"), strict=True)

    def load_seeded(self, seed: int):
        """Deterministic random weights in state_dict order (U(-b, b), b = 1/sqrt(fan_in)) for the parity tests."""
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for _, w in self.model.state_dict().items():
                bound = 1.0 / math.sqrt(w[0].numel())
                w.copy_((torch.rand(w.shape, generator=g) * 2 - 1) * bound)
        return self

    def _denoise(self, x, sigma):
        if sigma.shape[0] != x.shape[0]:
            sigma = sigma.expand(x.shape[0], 1, 1, 1)
        inp = torch.cat((x, sigma.to(x.dtype).expand(x.shape[0], 1, x.shape[2], x.shape[3])), dim=1)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            return self._tiled(inp)
        finally:
            torch.backends.cudnn.allow_tf32 = prev

    def _tiled(self, L):
        h, w = L.shape[-2:]
        if h * w <= self.MIN_SIZE ** 2:
            ph, pw = (-h) % self.MODULO, (-w) % self.MODULO
            return self.model(F.pad(L, (0, pw, 0, ph), mode="replicate"))[..., :h, :w]
        th, tw = (h // 2 // self.REFIELD + 1) * self.REFIELD, (w // 2 // self.REFIELD + 1) * self.REFIELD
        rows, cols = (slice(0, th), slice(h - th, h)), (slice(0, tw), slice(w - tw, w))
        run = self.model if h * w <= 4 * self.MIN_SIZE ** 2 else self._tiled
        E = None
        for i, rs in enumerate(rows):
            for j, cs in enumerate(cols):
                e = run(L[..., rs, cs])
                if E is None:
                    E = torch.zeros(e.shape[0], e.shape[1], h, w, dtype=L.dtype, device=L.device)
                # each quadrant of the output takes the matching corner of its (larger, overlapping) tile
                dst_r = slice(0, h // 2) if i == 0 else slice(h // 2, h)
                dst_c = slice(0, w // 2) if j == 0 else slice(w // 2, w)
                src_r = slice(0, h // 2) if i == 0 else slice(th - (h - h // 2), th)
                src_c = slice(0, w // 2) if j == 0 else slice(tw - (w - w // 2), tw)
                E[..., dst_r, dst_c] = e[..., src_r, src_c]
        return E


def get_denoiser(name: str):
    """pnp/prior.py:14-35.  Weight files are looked up under $DPROX_WEIGHTS (no network access here)."""
    root = os.environ.get("DPROX_WEIGHTS", os.path.expanduser("~/.cache/dprox/pnp_denoisers"))
    if name == "ffdnet_color":
        path = os.path.join(root, "ffdnet_color.pth")
        if not os.path.exists(path):
            raise FileNotFoundError(f"pretrained weights for {name!r} not found at {path}; set $DPROX_WEIGHTS or pass a "
                                    f"Denoiser instance to deep_prior(x, denoiser=obj)")
        return FFDNetColorDenoiser(path)
    if name in ("drunet", "drunet_color"):
        path = os.path.join(root, "drunet_gray.pth" if name == "drunet" else "drunet_color.pth")
        if not os.path.exists(path):
            raise FileNotFoundError(f"pretrained weights for {name!r} not found at {path}; set $DPROX_WEIGHTS or pass a "
                                    f"Denoiser instance to deep_prior(x, denoiser=obj)")
        return DRUNetDenoiser(1 if name == "drunet" else 3, path)
    raise NotImplementedError(f"denoiser {name!r} is not part of the lowered path (SURVEY §2 row 12); pass a Denoiser object")
