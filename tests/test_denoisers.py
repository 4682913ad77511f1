"""Denoisers behind `deep_prior` that run as torch modules on the prox hook: architecture,
state_dict compatibility and forward parity with the unmodified reference on seeded random weights (golden vectors)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def _check(device):
    from dprox_b200.denoisers import DRUNetDenoiser
    g = np.load(os.path.join(GOLDEN, "drunet_forward.npz"))
    den = DRUNetDenoiser(1).load_seeded(int(g["seed"]))
    assert list(den.model.state_dict().keys()) == list(g["keys"])           # published drunet_gray.pth loads unchanged
    den = den.to(device)
    with torch.no_grad():
        for tag in ("small", "tiled"):                                     # replicate-pad path / four-quadrant tiling path
            y = den.denoise(torch.from_numpy(g[f"{tag}_x"]).to(device), torch.tensor([0.07], device=device)).cpu()
            want = torch.from_numpy(g[f"{tag}_y"])
            assert float((y - want).norm() / want.norm()) < 1e-5, tag


def test_drunet_matches_reference_cpu():
    _check("cpu")


@pytest.mark.gpu
def test_drunet_matches_reference_gpu():
    _check("cuda")
