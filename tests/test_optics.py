"""DOE optics forward model (dprox_b200/optics.py <-> dprox/contrib/optic): constants on the CPU, the native
get_psf pipeline / img_psf_conv and their native backward kernels on the GPU, against values and autograd gradients of the
unmodified reference (tests/golden/doe_forward_model.npz)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def _golden():
    return dict(np.load(os.path.join(GOLDEN, "doe_forward_model.npz"), allow_pickle=False))


def _model(g):
    from dprox_b200.optics import RGBCollimator
    N, n = int(g["N"]), int(g["n"])
    return RGBCollimator(sensor_distance=15e-3, refractive_idcs=torch.tensor([1.4648, 1.4599, 1.4568]),
                         wave_lengths=torch.tensor([460, 550, 640]) * 1e-9, patch_size=n, sample_interval=float(g["sample_interval"]),
                         wave_resolution=(N, N))


def rel(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, np.complex128 if np.iscomplexobj(b) else np.float64)
    b = np.asarray(b, a.dtype)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_optics_constants_match_reference():
    g = _golden()
    m = _model(g)
    assert np.array_equal(m.aperture.numpy(), g["aperture"][0, 0])
    assert rel(m.propagator.H, g["H"]) < 1e-6
    assert m.height_map.height_map_sqrt.shape == (1, 1, int(g["N"]), int(g["N"]))


@pytest.mark.gpu
def test_get_psf_and_gradient_match_reference():
    g = _golden()
    m = _model(g).cuda()
    with torch.no_grad():
        m.height_map.height_map_sqrt.copy_(torch.from_numpy(g["h0"]))
    psf = m.get_psf()
    (psf * torch.from_numpy(g["wgt"]).cuda()).sum().backward()
    assert rel(psf, g["psf"]) < 2e-5, rel(psf, g["psf"])
    assert abs(float(psf.sum()) - 1.0) < 1e-5
    assert rel(m.height_map.height_map_sqrt.grad, g["g_h"]) < 2e-4, rel(m.height_map.height_map_sqrt.grad, g["g_h"])


@pytest.mark.gpu
def test_img_psf_conv_and_gradients_match_reference():
    from dprox_b200.optics import img_psf_conv
    g = _golden()
    img = torch.from_numpy(g["img"]).cuda().requires_grad_(True)
    psf = torch.from_numpy(g["psf"]).cuda().requires_grad_(True)
    y = img_psf_conv(img, psf, circular=True)
    (y * torch.from_numpy(g["w2"]).cuda()).sum().backward()
    assert rel(y, g["y"]) < 1e-5
    assert rel(img.grad, g["g_img"]) < 1e-5 and rel(psf.grad, g["g_psf"]) < 1e-5
