"""TEST INFRASTRUCTURE — not product code.

Import shim that lets the *unmodified* reference (`/root/reference/dprox`) be imported in the
authoring container, where several of its import-time dependencies are absent (SURVEY.md
App. B).  It is used ONLY by `oracle/make_golden.py` to generate `tests/golden/*.npz` and by
`oracle/check_against_reference.py`.  `/root/reference` does not exist on the GPU box, so
nothing under `tests/ -m gpu`, `bench.py` or `__graft_entry__.smoke()` imports this module.

Absent third-party modules are replaced by empty stub packages whose attributes are fresh
dummy *classes* (the reference subclasses a few of them at import time); none of them
carries hot-path arithmetic.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DPROX_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = {
    "imageio", "matplotlib", "skimage", "munch", "tfpnp", "torchlight", "torchlights",
    "termcolor", "tensorboardX", "cvxpy", "proximal", "graphviz", "IPython",
}


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dprox"))


def import_reference():
    """Return the reference's `dprox` module (imports it on first call)."""
    if "dprox" in sys.modules and getattr(sys.modules["dprox"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["dprox"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for root in list(_STUB_ROOTS):
        try:
            importlib.import_module(root)
            _STUB_ROOTS.discard(root)          # really installed: do not stub
        except Exception:
            pass
    sys.meta_path.append(_StubFinder())
    sys.path.insert(0, REFERENCE_ROOT)

    import numpy as np
    import scipy
    import scipy.misc

    rng = np.random.RandomState(0)
    if not hasattr(scipy.misc, "face"):
        scipy.misc.face = lambda gray=False: (rng.rand(768, 1024, 3) * 255).astype("uint8")
        scipy.misc.ascent = lambda: (rng.rand(512, 512) * 255).astype("uint8")
    if not hasattr(scipy, "finfo"):
        scipy.finfo = np.finfo

    import dprox  # noqa: E402
    return dprox
