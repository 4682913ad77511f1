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


class FFDNetColorDenoiser(Denoiser):
    """pnp/denoisers/wrapper.py:38-48."""

    def __init__(self, model_path=None, seed=None):
        super().__init__()
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
