"""Vectorised .kta / .lta readers (archnemesis_dist_b200.table_io) against the reference's record-by-record readers
(archnemesis/Spectroscopy_0.py:2733-2852, :2626-2729) on files written by the reference's own writers: the same
tuples, dtypes and bits, for whole tables and cropped ranges, uniform and explicit wavenumber grids."""
import os
import sys
import tempfile
import time

import numpy as np
import pytest

pytestmark = pytest.mark.reference


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        if isinstance(x, np.ndarray):
            assert isinstance(y, np.ndarray) and x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y)
        else:
            assert type(x) is type(y) and x == y, (x, y)


def test_readers_return_what_the_reference_returns():
    from oracle.ref_import import import_reference
    from archnemesis_dist_b200 import synthetic, table_io
    import_reference()
    sp = sys.modules["archnemesis.Spectroscopy_0"]
    d = tempfile.mkdtemp(prefix="ansb200_io_")
    t = synthetic.make_ktable(301, 20, 9, 7, 1, seed=5, zero_fraction=0.05)
    kta = os.path.join(d, "gas.kta")
    sp.write_ktable(kta, 6, 1, t["G_ORD"], t["DELG"], t["PRESS"], t["TEMP"], 301, 40.0, 0.5, 2.5, t["K"][..., 0])
    lta = os.path.join(d, "gas.lta")
    sp.write_lbltable(lta, 9, 7, 6, 1, t["PRESS"], t["TEMP"], 301, 40.0, 0.5, t["K"][:, 0, :, :, 0])
    fast_lbl = table_io.make_read_lbltable(sp.read_lbltable)
    for lo, hi in ((0.0, 1e10), (60.0, 120.0), (100.25, 100.75), (189.9, 1e10)):
        t0 = time.perf_counter()
        ref = sp.read_ktable(kta, lo, hi)
        t1 = time.perf_counter()
        got = table_io.read_ktable(kta, lo, hi)
        t2 = time.perf_counter()
        _same(got, ref)
        _same(table_io.read_ktable(kta[:-4], lo, hi), ref)          # the extension is optional
        _same(fast_lbl(lta, lo, hi), sp.read_lbltable(lta, lo, hi))
    assert (t2 - t1) < (t1 - t0)
    # install / uninstall rebinds the two module names read_tables resolves
    ref_k, ref_l = sp.read_ktable, sp.read_lbltable
    table_io.install_readers()
    try:
        assert sp.read_ktable is not ref_k and sp.read_lbltable is not ref_l
        _same(sp.read_ktable(kta, 50.0, 90.0), ref_k(kta, 50.0, 90.0))
    finally:
        table_io.uninstall_readers()
    assert sp.read_ktable is ref_k and sp.read_lbltable is ref_l
    with pytest.raises(IndexError):
        table_io.read_ktable(kta, 1e6, 2e6)                          # nothing in range: the reference's error
