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
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.model(x, sigma).float()
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


def get_denoiser(name: str):
    """pnp/prior.py:14-35.  Weight files are looked up under $DPROX_WEIGHTS (no network access here)."""
    root = os.environ.get("DPROX_WEIGHTS", os.path.expanduser("~/.cache/dprox/pnp_denoisers"))
    if name == "ffdnet_color":
        path = os.path.join(root, "ffdnet_color.pth")
        if not os.path.exists(path):
            raise FileNotFoundError(f"pretrained weights for {name!r} not found at {path}; set $DPROX_WEIGHTS or pass a "
                                    f"Denoiser instance to deep_prior(x, denoiser=obj)")
        return FFDNetColorDenoiser(path)
    raise NotImplementedError(f"denoiser {name!r} is not part of the lowered path (SURVEY §2 row 12); pass a Denoiser object")
