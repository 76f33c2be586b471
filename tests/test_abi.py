"""CPU: the C-ABI library loads and exports every symbol include/ansb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ansb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ansb200_\w+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for s in ("ansb200_table_create", "ansb200_kinterp", "ansb200_koverlap", "ansb200_gas_opacity", "ansb200_radiance",
              "ansb200_jacobian_project", "ansb200_lbl_absorption"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from archnemesis_dist_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libansb200.so missing: run python -m archnemesis_dist_b200.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    # and the ctypes signatures cover exactly the declared surface
    assert sorted(_lib.EXPORTS) == declared_symbols()
    loaded = _lib.load()
    assert loaded.ansb200_version() >= 100
    assert loaded.ansb200_last_error() is not None


def test_argument_errors_are_reported_without_a_gpu():
    """EINVAL paths return before any CUDA call, so they can be checked on the CPU box."""
    from archnemesis_dist_b200 import _lib
    lib = _lib.load()
    rc = lib.ansb200_koverlap(None, None, None, None, None, None, 1, 20, 1, 2, 0, None, None, None)
    assert rc == _lib.EINVAL
    with pytest.raises(ValueError):
        _lib.check(rc)
    rc = lib.ansb200_radiance(7, 0, *([None] * 20), 0, 0.0, 1, 1, 1, 1, 1, 1, 1, 1, None, None, None, None)
    assert rc == _lib.EINVAL and b"mode" in lib.ansb200_last_error()


def test_no_cpu_fallback():
    """Without a CUDA device the operators refuse to run instead of computing on the host."""
    import numpy as np
    import torch
    from archnemesis_dist_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.Table(np.zeros((1, 2, 2, 2, 1)))
