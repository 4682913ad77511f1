"""CPU check of the fused FFT engine's kernels (`-m "not gpu"`).

The kernels of delta-prox_b200/csrc/dpx_fused_kernels.cuh are compiled with g++ against a small CUDA-thread
emulator (tests/emu/cuda_emu.h: one OS thread per CUDA thread, __syncthreads -> std::barrier) and run on small
problems; the result is compared with the oracle.  This pins the tile / digit-reversal / packing index arithmetic
in the container that has no GPU; the same sources are what nvcc compiles for sm_100a.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import dprox_oracle as orc
from conftest import ROOT

EMU_DIR = os.path.join(ROOT, "tests", "emu")
SO = os.path.join(EMU_DIR, "_build", "libemu_fused.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(EMU_DIR, "emu_fused.cpp")
    csrc = os.path.join(ROOT, "delta-prox_b200", "csrc")
    deps = [src, os.path.join(EMU_DIR, "cuda_emu.h")] + [os.path.join(csrc, f) for f in (
        "dpx_fused_kernels.cuh", "dpx_fused_driver.cuh", "dpx_fft_core.cuh", "dpx_types.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-shared", "-fPIC", "-pthread", "-fvisibility=hidden", "-Wl,-Bsymbolic",
                               "-I", csrc, "-I", EMU_DIR, src, "-o", SO])
    lib = C.CDLL(SO)
    lib.emu_fused_run.restype = C.c_int
    return lib


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def run_emu(lib, b, psf, kinds, scales, alphas, rho, lams, T, hqs):
    """ADMM/HQS on sum_squares(conv(x,psf)-b) + sum_i prox_i(scale_i x) through the emulated fused kernels."""
    B, Cc, H, W = b.shape
    conv = orc.Conv(psf, orc.Identity())
    ktb = conv.adj(torch.from_numpy(b)).numpy()
    fb = np.ascontiguousarray(np.fft.rfft2(ktb.astype(np.float64)).astype(np.complex64).reshape(B * Cc, H, W // 2 + 1))
    otf = conv.FB(b.shape).numpy()
    dq = np.ascontiguousarray((np.abs(otf) ** 2)[0, :, :, : W // 2 + 1].astype(np.float32))
    x = b.copy()
    m = len(kinds)
    v = [np.ascontiguousarray(scales[i] * b) for i in range(m)]
    u = [np.zeros_like(b) for _ in range(m)]
    arr = lambda lst: (C.POINTER(C.c_float) * m)(*[fptr(a) for a in lst])
    rho_a = np.full(T, rho, dtype=np.float32)
    lam_a = np.ascontiguousarray(np.stack([np.full(T, l, dtype=np.float32) for l in lams]))
    rc = lib.emu_fused_run(B, Cc, H, W, m, (C.c_int * m)(*kinds), (C.c_float * m)(*scales), (C.c_float * m)(*alphas),
                           (C.c_float * m)(*([1.0] * m)), fptr(x), arr(v), arr(u), None,
                           fb.view(np.float32).ctypes.data_as(C.POINTER(C.c_float)), fptr(dq), 1,
                           C.c_float(float(sum(s * s for s in scales))), C.c_float(1e-7), fptr(rho_a), fptr(lam_a), T, int(hqs))
    assert rc in (0, 2)                # 2 = the plane-pair engine ran (even batch, shared schedules)
    run_emu.last_engine = "pairs" if rc == 2 else "planes"
    return x, v, u


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("H,W,method", [(64, 64, "admm"), (64, 128, "admm"), (128, 64, "hqs"), (256, 64, "admm"),
                                        (192, 64, "admm"), (64, 384, "hqs"),           # 3 * 2^k sides: radix-12 first pass
                                        (320, 64, "admm"), (64, 640, "hqs"),           # 5 * 2^k sides: radix-10 first pass
                                        (1080, 64, "hqs"),                             # camera height 15*9*8: radix-15 and radix-9 passes
                                        (64, 960, "admm")])                            # 12*10*8 as a row length
def test_fused_kernels_match_oracle(emu, H, W, method):
    g = torch.Generator().manual_seed(H + W)
    B, Cc, T = 2, 1 if H > 64 else 3, 4          # the reference's OTF builder only handles C in {1, 3}
    img = torch.rand(B, Cc, H, W, generator=g) - 0.3
    psf = orc.point_spread_function(5, 1.5)
    b = (orc.Conv(psf, orc.Identity()).fwd(img) + 0.01 * torch.randn(B, Cc, H, W, generator=g)).numpy()
    f1, f2 = orc.Term("norm1", alpha=0.5), orc.Term("nonneg")
    data = orc.Term("sum_squares", orc.Conv(psf, orc.Identity()), c=torch.from_numpy(b))
    want = orc.Solver([data, f1, f2], method).solve(torch.from_numpy(b), rhos=0.7, lams={f1: 0.05, f2: 0.02}, max_iter=T,
                                                   return_full_states=True)
    x, v, u = run_emu(emu, b, psf, [1, 0], [1.0, 1.0], [0.5, 1.0], 0.7, [0.05, 0.02], T, method == "hqs")
    assert rel(x, want[0].numpy()) < 5e-6
    assert rel(v[0], want[1][0].numpy()) < 5e-5 and rel(v[1], want[1][1].numpy()) < 5e-5
    if method == "admm":
        assert rel(u[0], want[2][0].numpy()) < 5e-5 and rel(u[1], want[2][1].numpy()) < 5e-5


@pytest.mark.parametrize("H,W,method", [(64, 128, "admm"), (128, 64, "hqs"), (1024, 64, "admm"),      # H = 1024: k_col_tma
                                        (192, 192, "admm"), (64, 1280, "hqs"),                   # 1280: radix-20 first pass
                                        (64, 1920, "hqs"),                                       # 20*12*8 as a row length
                                        (480, 480, "admm")])                                     # 12*10*4: radix-10 second pass over 4-point blocks
                                        # (720, 1200, 1440, 1600, 2160, 3840 points: tests/test_parity_gpu.py, against the cuFFT engine)
def test_fused_kernels_single_term_fast_path(emu, H, W, method):
    """One psi term: the in-place register path of k_row (template SINGLE)."""
    g = torch.Generator().manual_seed(H * 3 + W)
    B, Cc, T = 1, 3, 5
    img = torch.rand(B, Cc, H, W, generator=g) - 0.3
    psf = orc.point_spread_function(5, 1.5)
    b = (orc.Conv(psf, orc.Identity()).fwd(img) + 0.01 * torch.randn(B, Cc, H, W, generator=g)).numpy()
    f = orc.Term("nonneg")
    data = orc.Term("sum_squares", orc.Conv(psf, orc.Identity()), c=torch.from_numpy(b))
    want = orc.Solver([data, f], method).solve(torch.from_numpy(b), rhos=1.0, lams=0.02, max_iter=T, return_full_states=True)
    x, v, u = run_emu(emu, b, psf, [0], [1.0], [1.0], 1.0, [0.02], T, method == "hqs")
    assert rel(x, want[0].numpy()) < 5e-6 and rel(v[0], want[1][0].numpy()) < 5e-5
    if method == "admm":
        assert rel(u[0], want[2][0].numpy()) < 5e-5


def test_pair_engine_is_selected_and_matches_plane_engine(emu, monkeypatch):
    """Even batch + shared schedules -> plane-pair engine (k_rowz, paired k_col); same answer as the half-spectrum engine
    to fp32 round-off, on a non-square image with 3 channels and two prox terms."""
    g = torch.Generator().manual_seed(17)
    B, Cc, H, W, T = 4, 3, 64, 128, 3
    img = torch.rand(B, Cc, H, W, generator=g) - 0.3
    psf = orc.point_spread_function(5, 1.5)
    b = (orc.Conv(psf, orc.Identity()).fwd(img) + 0.01 * torch.randn(B, Cc, H, W, generator=g)).numpy()
    monkeypatch.delenv("DPX_EMU_NO_PAIRS", raising=False)
    got = run_emu(emu, b, psf, [1, 0], [1.0, 1.0], [0.5, 1.0], 0.7, [0.05, 0.02], T, False)
    assert run_emu.last_engine == "pairs"
    monkeypatch.setenv("DPX_EMU_NO_PAIRS", "1")
    base = run_emu(emu, b, psf, [1, 0], [1.0, 1.0], [0.5, 1.0], 0.7, [0.05, 0.02], T, False)
    assert run_emu.last_engine == "planes"
    assert rel(got[0], base[0]) < 2e-6 and rel(got[2][0], base[2][0]) < 2e-5 and rel(got[1][1], base[1][1]) < 2e-5


def test_tma_column_kernel_with_plane_pairs(emu, monkeypatch):
    """H = 1024 selects the persistent bulk-copy (TMA) column kernel; B = 2 the plane-pair engine on top of it.  The emulated
    grid has 3 'SMs' x 2 CTAs for 34 tiles, so every CTA walks its 3-buffer ring several times."""
    g = torch.Generator().manual_seed(23)
    B, Cc, H, W, T = 2, 1, 1024, 64, 3
    img = torch.rand(B, Cc, H, W, generator=g) - 0.3
    psf = orc.point_spread_function(5, 1.5)
    b = (orc.Conv(psf, orc.Identity()).fwd(img) + 0.01 * torch.randn(B, Cc, H, W, generator=g)).numpy()
    f = orc.Term("nonneg")
    data = orc.Term("sum_squares", orc.Conv(psf, orc.Identity()), c=torch.from_numpy(b))
    want = orc.Solver([data, f], "admm").solve(torch.from_numpy(b), rhos=1.0, lams=0.02, max_iter=T, return_full_states=True)
    for no_pairs in ("", "1"):
        if no_pairs:
            monkeypatch.setenv("DPX_EMU_NO_PAIRS", "1")
        else:
            monkeypatch.delenv("DPX_EMU_NO_PAIRS", raising=False)
        x, v, u = run_emu(emu, b, psf, [0], [1.0], [1.0], 1.0, [0.02], T, False)
        assert run_emu.last_engine == ("planes" if no_pairs else "pairs")
        assert rel(x, want[0].numpy()) < 5e-6 and rel(u[0], want[2][0].numpy()) < 5e-5


@pytest.mark.parametrize("B", [1, 2])
def test_staged_xupdate_matches_first_iteration(emu, B, monkeypatch):
    _staged_xupdate_first_iteration(emu, B, monkeypatch)


@pytest.mark.parametrize("H,W", [(64, 2560), (4096, 64)])
def test_solo_tiles_run_512_threads(emu, H, W):
    """Tiles of which only one fits an SM run 512 threads per CTA (RowZPersistSmem::SOLO rows from 2560 points in the persistent pair
    kernel, ColThreads for 3840- / 4096-point columns): same arithmetic, task loops strided by the larger block."""
    g = torch.Generator().manual_seed(H + 7 * W)
    B, Cc, T = 2, 1, 3
    img = torch.rand(B, Cc, H, W, generator=g) - 0.3
    psf = orc.point_spread_function(5, 1.5)
    b = (orc.Conv(psf, orc.Identity()).fwd(img) + 0.01 * torch.randn(B, Cc, H, W, generator=g)).numpy()
    f = orc.Term("nonneg")
    data = orc.Term("sum_squares", orc.Conv(psf, orc.Identity()), c=torch.from_numpy(b))
    want = orc.Solver([data, f], "admm").solve(torch.from_numpy(b), rhos=1.0, lams=0.02, max_iter=T, return_full_states=True)
    x, v, u = run_emu(emu, b, psf, [0], [1.0], [1.0], 1.0, [0.02], T, False)
    assert run_emu.last_engine == "pairs"
    assert rel(x, want[0].numpy()) < 5e-6 and rel(v[0], want[1][0].numpy()) < 5e-5 and rel(u[0], want[2][0].numpy()) < 5e-5


def _staged_xupdate_first_iteration(emu, B, monkeypatch):
    """dpx_stage_xupdate's fused form (ROW_FIRST -> k_col -> ROW_XONLY, used when an external prox sits between the stages)
    gives bit-for-bit the x of a one-iteration fused run, on both engines, and leaves v / u untouched."""
    g = torch.Generator().manual_seed(29)
    Cc, H, W = 3, 64, 128
    img = torch.rand(B, Cc, H, W, generator=g) - 0.3
    psf = orc.point_spread_function(5, 1.5)
    b = (orc.Conv(psf, orc.Identity()).fwd(img) + 0.01 * torch.randn(B, Cc, H, W, generator=g)).numpy()
    monkeypatch.delenv("DPX_EMU_XUPDATE", raising=False)
    full = run_emu(emu, b, psf, [1, 0], [1.0, 1.0], [0.5, 1.0], 0.7, [0.05, 0.02], 1, False)
    monkeypatch.setenv("DPX_EMU_XUPDATE", "1")
    x, v, u = run_emu(emu, b, psf, [1, 0], [1.0, 1.0], [0.5, 1.0], 0.7, [0.05, 0.02], 1, False)
    assert run_emu.last_engine == ("pairs" if B == 2 else "planes")
    assert np.array_equal(x, full[0])
    assert np.array_equal(v[0], b) and not u[0].any()          # state untouched by the x-update stage


@pytest.mark.parametrize("B", [1, 2])
def test_staged_xupdate_with_stencil_gradients(emu, B, monkeypatch):
    """Anisotropic TV through the fused engine: the right-hand side t = sum_i K_i^T (v_i - u_i) (stencil kernel; here the
    oracle's adjoint) enters as ONE identity term and the solve's denominator gains rho sum_i |F(K_i)|^2 (packed like the
    quadratic diagonal); x must equal the oracle's x after one ADMM iteration from the state v_i = K_i x0, u_i = 0."""
    g = torch.Generator().manual_seed(37)
    Cc, H, W = 1, 64, 128
    img = torch.rand(B, Cc, H, W, generator=g)
    psf = orc.point_spread_function(5, 1.5)
    conv = orc.Conv(psf, orc.Identity())
    b = (conv.fwd(img) + 0.01 * torch.randn(B, Cc, H, W, generator=g)).numpy()
    th, tw = orc.Term("norm1", orc.Grad(0, orc.Identity())), orc.Term("norm1", orc.Grad(1, orc.Identity()))
    data = orc.Term("sum_squares", conv, c=torch.from_numpy(b))
    want = orc.Solver([data, th, tw], "admm").solve(torch.from_numpy(b), rhos=0.7, lams=0.05, max_iter=1)
    bt = torch.from_numpy(b)
    t = sum(tm.op.adj(tm.op.fwd(bt)) for tm in (th, tw)).numpy()          # sum_i K_i^T (v_i - 0)
    v, x = [np.ascontiguousarray(t)], b.copy()
    ktb = conv.adj(bt).numpy()
    fb = np.ascontiguousarray(np.fft.rfft2(ktb.astype(np.float64)).astype(np.complex64).reshape(B * Cc, H, W // 2 + 1))
    dq = np.ascontiguousarray((np.abs(conv.FB(b.shape).numpy()) ** 2)[0, :, :, : W // 2 + 1].astype(np.float32))
    monkeypatch.setenv("DPX_EMU_XUPDATE", "1")
    monkeypatch.setenv("DPX_EMU_LINOPS", "1,2")
    arr = lambda lst: (C.POINTER(C.c_float) * 1)(*[fptr(a) for a in lst])
    rho_a, lam_a = np.full(1, 0.7, np.float32), np.full((1, 1), 0.05, np.float32)
    rc = emu.emu_fused_run(B, Cc, H, W, 1, (C.c_int * 1)(1), (C.c_float * 1)(1.0), (C.c_float * 1)(1.0), (C.c_float * 1)(1.0),
                           fptr(x), arr(v), arr([np.zeros_like(b)]), None, fb.view(np.float32).ctypes.data_as(C.POINTER(C.c_float)),
                           fptr(dq), 1, C.c_float(0.0), C.c_float(1e-7), fptr(rho_a), fptr(lam_a), 1, 1)
    assert rc == (2 if B == 2 else 0)
    assert rel(x, want.numpy()) < 5e-6
