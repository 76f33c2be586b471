"""Shapes beyond the native limits of libansb200 are left to the reference implementation instead of raising from
inside CIRSrad (include/ansb200.h: ANSB200_MAX_NG, ANSB200_MAX_NGAS; csrc/klbl.cu, csrc/convolve.cu)."""
import types

import numpy as np
import pytest

from archnemesis_dist_b200 import _lib, forward_model


def _sp(ilbl, shape, ng, ngas):
    return types.SimpleNamespace(ILBL=ilbl, K=np.zeros(shape), NG=ng, NGAS=ngas)


def test_ktable_within_and_beyond_the_limits():
    ok = _sp(0, (2, 20, 3, 3, 6), 20, 6)
    assert forward_model._table_on_device(ok)
    too_many_g = _sp(0, (2, 50, 3, 3, 2), 50, 2)                  # an NG = 50 k-table: reference-valid, not native
    assert not forward_model._table_on_device(too_many_g)
    too_many_gases = _sp(0, (2, 10, 2, 2, _lib.MAX_NGAS + 1), 10, _lib.MAX_NGAS + 1)
    assert not forward_model._table_on_device(too_many_gases)


def test_lbl_table_gas_limit():
    assert forward_model._table_on_device(_sp(2, (4, 3, 3, 5), 1, 5))
    assert not forward_model._table_on_device(_sp(2, (4, 2, 2, _lib.MAX_LBL_NGAS + 1), 1, _lib.MAX_LBL_NGAS + 1))


def test_no_table_or_runtime_lbl_stays_on_the_reference():
    assert not forward_model._table_on_device(types.SimpleNamespace(ILBL=0, K=None, NG=20, NGAS=3))
    assert not forward_model._table_on_device(_sp(1, (4, 3, 3, 5), 1, 5))


@pytest.mark.gpu
def test_jacobian_project_rejects_a_mismatched_matrix():
    import torch
    from archnemesis_dist_b200 import ops
    dspec = torch.zeros((4, 2, 5, 7), dtype=torch.float64, device="cuda")
    good = torch.zeros((2, 35, 3), dtype=torch.float64, device="cuda")
    assert tuple(ops.jacobian_project(dspec, good).shape) == (4, 2, 3)
    with pytest.raises(ValueError):
        ops.jacobian_project(dspec, torch.zeros((2, 30, 3), dtype=torch.float64, device="cuda"))     # other NPAR
    with pytest.raises(ValueError):
        ops.jacobian_project(dspec, torch.zeros((1, 35, 3), dtype=torch.float64, device="cuda"))     # other path set
    with pytest.raises(ValueError):
        ops.jacobian_project(dspec, good.float())
