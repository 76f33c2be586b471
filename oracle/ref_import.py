"""TEST INFRASTRUCTURE ONLY -- imports the live archNEMESIS reference in the build container.

The reference (``/root/reference``, read-only) is pure Python but imports h5py, matplotlib and a
few other packages at module scope (``archnemesis/helpers/h5py_helper.py:7``,
``archnemesis/ForwardModel_0.py:32``, ``archnemesis/Spectroscopy_0.py:48``) that are absent here
and never touched on the legacy-text + ``.kta`` path.  We install a ``sys.meta_path`` finder that
hands out permissive stub modules for those roots so ``import archnemesis`` succeeds.

Only ``tests/`` (container-side validation) and ``oracle/make_golden.py`` may import this module.
It is never imported by the product package.  On the GPU box (no /root/reference) it resolves to the mirror that
``oracle/make_ref.py`` stages under ``oracle/_ref`` (git-ignored, travels with the snapshot).
"""
import importlib.abc
import importlib.machinery
import logging
import os
import sys
import types

def _reference_root():
    """/root/reference in the build container; on the GPU box the byte-for-byte mirror oracle/make_ref.py staged."""
    env = os.environ.get("ANSB200_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/archnemesis"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _reference_root()
_STUB_ROOTS = {"h5py", "matplotlib", "mpl_toolkits", "corner", "pymultinest", "cdsapi", "hapi",
               "bs4", "mpi4py", "pygrib"}


class _StubMeta(type):
    # annotations such as ``h5py.Group | h5py.Dataset`` are evaluated at import time
    def __or__(cls, other):
        return cls

    def __ror__(cls, other):
        return cls

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _StubMeta(name, (), {})


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _StubMeta(name, (), {"__init__": lambda self, *a, **k: None,
                                   "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "archnemesis"))


def import_reference():
    """Return the imported ``archnemesis`` package (the unmodified reference)."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    if "archnemesis" in sys.modules:
        return sys.modules["archnemesis"]
    missing = []
    for root in sorted(_STUB_ROOTS):
        try:
            __import__(root)
        except Exception:
            missing.append(root)
    if missing and not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        _STUB_ROOTS.intersection_update(missing)
        sys.meta_path.insert(0, _StubFinder())
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/ansb200_numba_cache")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    logging.disable(logging.WARNING)
    import archnemesis  # noqa: E402
    return archnemesis
