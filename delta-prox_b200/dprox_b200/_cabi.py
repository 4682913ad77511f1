"""ctypes binding of libdprox_b200.so (include/dprox_b200.h).

This is the ONLY place the package touches native code.  There is no CPU or eager-PyTorch fallback:
if the shared library is missing, or a call returns a non-zero status, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

HINT_CHANNEL_SHARED_DIAG = 1
ABI_VERSION = 1
MAX_PSI = 8

# enums (keep in sync with include/dprox_b200.h)
ALGO_ADMM, ALGO_HQS, ALGO_ADMM_VXU, ALGO_PGD, ALGO_LADMM = 0, 1, 2, 3, 4
X_FREQ_DIAG, X_SPATIAL_DIAG = 0, 1
PROX_NONNEG, PROX_L1, PROX_L2SQ, PROX_BOX, PROX_EXTERNAL, PROX_ISO_TV = 0, 1, 2, 3, 4, 5
LINOP_IDENTITY, LINOP_GRAD_H, LINOP_GRAD_W, LINOP_GRAD_HW = 0, 1, 2, 3
FFT_AUTO, FFT_CUFFT, FFT_FUSED = 0, 1, 2
ENGINE_NONE, ENGINE_CUFFT, ENGINE_FUSED_PLANES, ENGINE_FUSED_PAIRS, ENGINE_FUSED_FLAT = -1, 0, 1, 2, 3


class PsiDesc(C.Structure):
    _fields_ = [("prox", C.c_int32), ("linop", C.c_int32), ("scale", C.c_float), ("alpha", C.c_float),
                ("beta", C.c_float), ("box_lo", C.c_float), ("box_hi", C.c_float)]


class ProblemDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("batch", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32),
                ("width", C.c_int32), ("algo", C.c_int32), ("xupdate", C.c_int32), ("n_psi", C.c_int32),
                ("psi", PsiDesc * MAX_PSI), ("eps", C.c_float), ("fft_backend", C.c_int32),
                ("eps_delta", C.c_int32)]


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DPROX_B200_LIB", os.path.join(os.path.dirname(_PKG_DIR), "lib", "libdprox_b200.so"))

_lib = None

_VP, _I, _F, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_PP = C.POINTER(C.c_void_p)
_IP = C.POINTER(C.c_int)

# name -> (restype, argtypes); this table is also what tests/test_cabi.py checks against the header
SIGNATURES = {
    "dpx_abi_version": (_I, []),
    "dpx_last_error": (C.c_char_p, []),
    "dpx_build_info": (C.c_char_p, []),
    "dpx_launch_count": (C.c_ulonglong, []),
    "dpx_plan_create": (_I, [C.POINTER(ProblemDesc), C.POINTER(_VP)]),
    "dpx_plan_destroy": (None, [_VP]),
    "dpx_plan_workspace_bytes": (_SZ, [_VP]),
    "dpx_plan_set_freq_constants": (_I, [_VP, _VP, _VP, _I, _VP, _VP]),
    "dpx_plan_set_rhs": (_I, [_VP, _VP, _VP]),
    "dpx_plan_set_hint": (_I, [_VP, _I, _I]),
    "dpx_plan_set_rhs_spectral": (_I, [_VP, _VP, _VP, _I, _F, _VP]),
    "dpx_plan_set_spatial_constants": (_I, [_VP, _VP, _VP, _I, _VP]),
    "dpx_plan_set_psi_offset": (_I, [_VP, _I, _VP, _VP]),
    "dpx_plan_set_spatial_psi_diag": (_I, [_VP, _VP, _VP]),
    "dpx_plan_engine_mode": (_I, [_VP]),
    "dpx_pad2d": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "dpx_augment": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "dpx_iters": (_I, [_VP, _VP, _PP, _PP, _VP, _I, _PP, _IP, _I, _I, _VP, _VP]),
    "dpx_stage_xupdate": (_I, [_VP, _VP, _PP, _PP, _VP, _I, _I, _VP]),
    "dpx_stage_prox": (_I, [_VP, _VP, _PP, _PP, _PP, _IP, _I, _VP]),
    "dpx_stage_dual_external": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP]),
    "dpx_xsolve": (_I, [_VP, _VP, _VP, _I, _I, _VP, _VP]),
    "dpx_xsolve_backward": (_I, [_VP, _VP, _VP, _VP, _I, _I, _VP, _VP, _VP]),
    "dpx_init_state": (_I, [_VP, _VP, _PP, _PP, _VP]),
    "dpx_spectral_filter": (_I, [_VP, _VP, _VP, _I, _I, _VP, _VP]),
    "dpx_prox_apply": (_I, [_I, _VP, _VP, _I, _F, _F, _F, _F, _VP, _VP, _I, _SZ, _VP]),
    "dpx_prox_backward": (_I, [_I, _VP, _VP, _I, _F, _F, _F, _F, _VP, _VP, _VP, _VP, _I, _SZ, _VP]),
    "dpx_lincomb": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _SZ, _VP]),
    "dpx_axpby": (_I, [_VP, _F, _VP, _F, _VP, _SZ, _VP]),
    "dpx_grad_apply": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _F, _VP]),
    "dpx_mul_apply": (_I, [_VP, _VP, _VP, _I, _I, _SZ, _VP]),
    "dpx_cg_dot": (_I, [_VP, _VP, _VP, _I, _SZ, _VP]),
    "dpx_cg_update": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _SZ, _VP]),
    "dpx_cg_direction": (_I, [_VP, _VP, _VP, _VP, _I, _SZ, _VP]),
    "dpx_absmax": (_I, [_VP, _VP, _I, _SZ, _VP]),
    "dpx_cg_gate": (_I, [_VP, _VP, _I, _I, _VP, _VP, _I, _VP]),
    "dpx_ffdnet_available": (_I, []),
    "dpx_ffdnet_create": (_I, [_I, _I, C.POINTER(_VP)]),
    "dpx_ffdnet_destroy": (None, [_VP]),
    "dpx_ffdnet_set_precision": (_I, [_VP, _I]),
    "dpx_ffdnet_set_layer": (_I, [_VP, _I, _VP, _VP, _I, _I, _VP]),
    "dpx_ffdnet_forward": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _I, _VP]),
    "dpx_ffdnet_forward_train": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _I, _VP]),
    "dpx_ffdnet_backward": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "dpx_ffdnet_backward_params": (_I, [_VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _I, _I, _VP]),
    "dpx_ffdnet_conv_layer": (_I, [_VP, _I, _I, _I, _VP, _VP, _I, _I, _I, _VP]),
    "dpx_ffdnet_wgrad_layer": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "dpx_csmri_prox": (_I, [_VP, _VP, _VP, _I, _VP, _I, _F, _VP, _I, _I, _I, _I, _VP]),
    "dpx_real_to_complex": (_I, [_VP, _VP, _SZ, _VP]),
    "dpx_complex_real": (_I, [_VP, _VP, _SZ, _VP]),
    "dpx_c2c": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "dpx_cmul": (_I, [_VP, _VP, _VP, _SZ, _SZ, _I, _F, _VP]),
    "dpx_cmul_reduce": (_I, [_VP, _VP, _VP, _SZ, _I, _F, _VP]),
    "dpx_doe_field": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "dpx_doe_field_backward": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "dpx_abs2_pool": (_I, [_VP, _VP, _I, _I, _I, _I, _F, _VP]),
    "dpx_abs2_pool_backward": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _F, _VP]),
    "dpx_normalize_sum": (_I, [_VP, _VP, _VP, _SZ, _VP]),
    "dpx_normalize_sum_backward": (_I, [_VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "dpx_solve_host": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "dpx_resid_reduce": (_I, [_VP, _VP, _I, _I, _VP]),
}


def lib():
    """Load (once) and return the native library; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"dprox_b200: native library not found at {LIB_PATH}. Build it with "
            f"`make -C delta-prox_b200/csrc` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            f"There is no CPU / eager fallback.")
    handle = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, args
    if handle.dpx_abi_version() != ABI_VERSION:
        raise RuntimeError(f"dprox_b200: ABI mismatch (python {ABI_VERSION}, library {handle.dpx_abi_version()})")
    _lib = handle
    return _lib


def last_error() -> str:
    return lib().dpx_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"dprox_b200: {what} failed (status {rc}): {last_error()}")


# ---- tensor plumbing ------------------------------------------------------------------------------

def require_cuda_f32(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"dprox_b200: {name} lives on {t.device}; this backend only computes on CUDA devices "
                           f"(no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def ptr_array(ts: Optional[Sequence[torch.Tensor]]):
    if ts is None or len(ts) == 0:
        return None
    arr = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    return C.cast(arr, _PP)


def int_array(vals: Optional[Sequence[int]]):
    if vals is None or len(vals) == 0:
        return None
    arr = (C.c_int * len(vals))(*vals)
    return C.cast(arr, _IP)


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class NativePlan:
    """RAII wrapper of a dpx_plan*."""

    def __init__(self, desc: ProblemDesc, device: torch.device):
        self.desc = desc
        self.device = device
        self.const_version = 0          # bumped by whoever re-sets the plan's constants (see autograd.XSolve)
        self._h = C.c_void_p()
        with torch.cuda.device(device):
            check(lib().dpx_plan_create(C.byref(desc), C.byref(self._h)), "dpx_plan_create")

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("plan already destroyed")
        return self._h

    def engine_mode(self) -> int:
        """ENGINE_* of the last fused call (which transform engine ran)."""
        return int(lib().dpx_plan_engine_mode(self.handle))

    def workspace_bytes(self) -> int:
        return int(lib().dpx_plan_workspace_bytes(self.handle))

    def close(self):
        if self._h:
            lib().dpx_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
