"""The C-ABI library loads on a box without a GPU and exports exactly what include/dprox_b200.h declares
(no compute calls here).  Also checks the ctypes table of the Python binding against the header."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "dprox_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpx_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    from dprox_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        g.build()
    return _cabi.LIB_PATH


def test_header_declares_the_expected_surface():
    fns = header_functions()
    for must in ("dpx_plan_create", "dpx_plan_destroy", "dpx_plan_set_freq_constants", "dpx_iters", "dpx_stage_xupdate",
                 "dpx_stage_prox", "dpx_xsolve", "dpx_cg_update", "dpx_cg_direction", "dpx_last_error", "dpx_solve_host"):
        assert must in fns
    assert 'extern "C"' in open(HEADER).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)       # no torch types in the ABI


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [f for f in header_functions() if f not in exported]
    assert not missing, missing
    extra = sorted(s for s in exported if s.startswith("dpx_") and s not in header_functions())
    assert not extra, f"exported but undeclared: {extra}"


def test_ctypes_table_matches_header_and_loads(lib_path):
    from dprox_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == header_functions()
    lib = _cabi.lib()                                   # dlopen + ABI version check, no GPU needed
    assert lib.dpx_abi_version() == _cabi.ABI_VERSION
    assert b"sm_100a" in lib.dpx_build_info()
    # argument validation happens before any CUDA call
    d = _cabi.ProblemDesc()
    d.abi_version = 999
    h = C.c_void_p()
    assert lib.dpx_plan_create(C.byref(d), C.byref(h)) == 1 and b"ABI version" in lib.dpx_last_error()
    assert C.sizeof(_cabi.PsiDesc) == 28 and C.sizeof(_cabi.ProblemDesc) == 8 * 4 + 8 * 28 + 3 * 4


def test_cuda_sources_target_sm100a():
    mk = open(os.path.join(ROOT, "delta-prox_b200", "csrc", "Makefile")).read()
    assert "arch=compute_100a,code=sm_100a" in mk and "-lineinfo" in mk
