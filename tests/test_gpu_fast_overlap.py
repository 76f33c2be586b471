"""The fast overlap kernel (csrc/koverlap_fast.cu) and its hand-over to the general kernel.

The fast kernel sorts truncated keys and is exact only where a bin edge falls (k_overlapg / rankg,
archnemesis/ForwardModel_0.py:5842-6026); cells whose result depends on the order of equal keys go to the general
kernel through a device-side work list.  These tests pin (a) fast kernel == oracle (numba tie order) on cases it
handles itself, (b) the hand-over: tie-heavy, non-monotone and dead-gas cells still give the reference's numbers and
are counted, (c) fast dispatch == general kernel to rounding, with and without gradients, fused and unfused,
(d) both gradient widths (NGAS <= 6 and NGAS up to 14).
"""
import numpy as np
import pytest

from tests.util import relerr, colerr, cpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    from archnemesis_dist_b200 import ops, plan, synthetic
    from oracle import oracle
    oracle.set_sort_mode(oracle.NUMBA_ORDER)
    return dict(torch=torch, ops=ops, plan=plan, syn=synthetic, orc=oracle, nt=oracle.max_threads())


@pytest.fixture(autouse=True)
def stats_mode(mods):
    old = mods["ops"].overlap_mode(2)
    yield
    mods["ops"].overlap_mode(old)


def _case(mods, nwave, ngas, seed, nlay=40, zero_fraction=0.0):
    c = mods["syn"].make_fm_case(nwave=nwave, ng=20, ngas=ngas, nlay=nlay, npro=nlay, nx=8, nvmr=max(8, ngas), seed=seed,
                                 zero_fraction=zero_fraction)
    tab, orc, nt = c["tab"], mods["orc"], mods["nt"]
    k, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True, nthreads=nt)
    return c, k, dkdT


def _check(mods, c, k, dkdT, tol_tau=1e-13, tol_col=1e-13):
    ops, orc, nt = mods["ops"], mods["orc"], mods["nt"]
    delg = c["tab"]["DELG"]
    otab = ops.OverlapTables(delg)
    kd, dd, am = ops.to_dev(k), ops.to_dev(dkdT), ops.to_dev(c["amount"])
    rt, rd = orc.k_overlap(delg, k, c["amount"], dkdT=dkdT, nthreads=nt)
    rt0 = orc.k_overlap(delg, k, c["amount"], nthreads=nt)
    tau, dk = ops.koverlap(kd, am, otab, dkdT=dd)
    st = ops.overlap_stats()
    assert relerr(cpu(tau), rt) < tol_tau
    got = cpu(dk)
    for col in range(rd.shape[-1]):
        assert colerr(got[..., col], rd[..., col]) < tol_col, col
    tau0 = ops.koverlap(kd, am, otab)
    st0 = ops.overlap_stats()
    assert relerr(cpu(tau0), rt0) < tol_tau
    # the general kernel alone gives the same numbers to rounding
    ops.overlap_mode(1)
    tg, dg = ops.koverlap(kd, am, otab, dkdT=dd)
    ops.overlap_mode(2)
    assert relerr(cpu(tau), cpu(tg)) < 1e-13
    for col in range(rd.shape[-1]):
        assert colerr(got[..., col], cpu(dg)[..., col]) < 1e-13, col
    return st, st0


def test_fast_kernel_handles_the_config2_shape_itself(mods):
    """Config-2 recipe (6 gases, random amounts over 6 decades): every cell stays in the fast kernel, both kinds of fold
    occur, and the result is the oracle's."""
    c, k, dkdT = _case(mods, 96, 6, 7, nlay=100)
    st, st0 = _check(mods, c, k, dkdT)
    ncell = 96 * 100
    assert st["handed_over"] == 0 and st0["handed_over"] == 0
    assert st["sorted_folds"] > ncell and st["static_folds"] > ncell // 2
    assert st["sorted_folds"] + st["static_folds"] == 5 * ncell


def test_fast_kernel_wide_gradient_rows(mods):
    """11 gases: the 16-column instantiation (two mma column tiles)."""
    c, k, dkdT = _case(mods, 24, 11, 3)
    st, _ = _check(mods, c, k, dkdT)
    assert st["handed_over"] == 0
    assert st["sorted_folds"] + st["static_folds"] == 10 * 24 * 40


def test_handover_of_tied_and_dead_gases(mods):
    """A gas 1e-25 below the rest makes whole rows of keys equal (the reference's result then depends on numba's tie
    order), dead gases take the short cuts, dominant gases the data-independent orders: the fast kernel must hand the
    cells it cannot decide to the general kernel and the mixture must still equal the oracle bit pattern for bit
    pattern within rounding."""
    c, k, dkdT = _case(mods, 64, 6, 23, zero_fraction=0.2)
    rng = np.random.default_rng(5)
    regime = rng.integers(0, 4, size=(64, 6))
    f = np.choose(regime, [1.0, 1e-25, 1e9, 1e-8])
    k = k * f[:, None, None, :]
    dkdT = dkdT * f[:, None, None, :]
    st, st0 = _check(mods, c, k, dkdT)
    assert 0 < st["handed_over"] < 64 * 40
    assert st["exact_tie"] > 0


def test_handover_of_non_monotone_k(mods):
    """k not ascending in g (scrambled ordinates): nothing the fast kernel assumes holds; all such cells are handed over."""
    c, k, dkdT = _case(mods, 16, 4, 11)
    perm = np.random.default_rng(2).permutation(20)
    k = np.ascontiguousarray(k[:, perm])
    dkdT = np.ascontiguousarray(dkdT[:, perm])
    st, _ = _check(mods, c, k, dkdT)
    assert st["handed_over"] == 16 * 40 and st["non_monotone"] == 16 * 40


def test_near_equal_gases_collide_in_the_packed_keys(mods):
    """Two gases with the same k-distribution and amounts: the key matrix is symmetric, every key a_i + b_j ties
    exactly with a_j + b_i, and many more agree to 17 mantissa bits.  Edges that fall on exact ties go to the general
    kernel; the rest is resolved from the exact keys inside the fast kernel."""
    c, k, dkdT = _case(mods, 32, 3, 5)
    k[..., 1] = k[..., 0]
    dkdT[..., 1] = dkdT[..., 0]
    c["amount"][1] = c["amount"][0]
    k[..., 2] = k[..., 0] * (1.0 + 1e-7)          # nearly, not exactly, equal to the mixture's scale
    st, _ = _check(mods, c, k, dkdT)
    assert st["exact_tie"] > 0


def test_fused_entry_point_uses_the_fast_kernel(mods):
    """ansb200_gas_opacity (what bench.py times): equal to k-interp + overlap bit for bit, nothing handed over."""
    ops, plan = mods["ops"], mods["plan"]
    c, k, dkdT = _case(mods, 48, 6, 7, nlay=100)
    tab = c["tab"]
    hp = plan.kinterp_plan(tab["PRESS"], tab["TEMP"], c["press"], c["temp"], True)
    T = ops.Table(tab["K"])
    dp = ops.DevicePlan(hp, True)
    otab = ops.OverlapTables(tab["DELG"])
    am = ops.to_dev(c["amount"])
    ft, fd = ops.gas_opacity(T, dp, am, otab, True)
    assert ops.overlap_stats()["handed_over"] == 0
    kd, dd = ops.kinterp(T, dp, True)
    ut, ud = ops.koverlap(kd, am, otab, dkdT=dd)
    assert np.array_equal(cpu(ft), cpu(ut)) and np.array_equal(cpu(fd), cpu(ud))
    f0 = ops.gas_opacity(T, dp, am, otab, False)
    assert np.array_equal(cpu(f0), cpu(ops.koverlap(ops.kinterp(T, dp, False), am, otab)))
    T.close()


def test_float64_quadrature_goes_to_the_general_kernel(mods):
    """DELG as float64 (HDF5 tables): the weights are not float32 products, the fast kernel declines everything."""
    ops, orc, nt = mods["ops"], mods["orc"], mods["nt"]
    c, k, dkdT = _case(mods, 8, 3, 9)
    x, w = np.polynomial.legendre.leggauss(20)
    delg = 0.5 * w                                    # float64
    otab = ops.OverlapTables(delg)
    kd, dd, am = ops.to_dev(k), ops.to_dev(dkdT), ops.to_dev(c["amount"])
    tau, dk = ops.koverlap(kd, am, otab, dkdT=dd)
    assert ops.overlap_stats()["handed_over"] == -1
    rt, rd = orc.k_overlap(delg, k, c["amount"], dkdT=dkdT, nthreads=nt)
    assert relerr(cpu(tau), rt) < 1e-13
    for col in range(rd.shape[-1]):
        assert colerr(cpu(dk)[..., col], rd[..., col]) < 1e-13, col


def test_partly_static_orders(mods):
    """Key matrices whose rows stop interleaving from some row on: the fast kernel sorts the head only and takes the
    bins / straddlers of the tail from the row-major tables.  Every head length 2 .. 20 occurs, with smooth and with
    grid-valued operands (exact ties inside the head, between the head and the tail, and on bin edges -- the latter go
    to the general kernel); both must give the oracle's numbers."""
    rng = np.random.default_rng(11)
    ng, nlay, reps = 20, 19 * 2, 12
    _, delg = mods["syn"].gauss_legendre_01(ng)
    k = np.zeros((reps, ng, nlay, 2))
    for w in range(reps):
        for l in range(nlay):
            hl = 2 + l % 19                      # rows hl-1 | hl is the last boundary that may interleave
            grid = l >= 19
            if grid:
                b = 1.0 + np.arange(ng) * rng.integers(1, 4) / 32.0
                spread = b[-1] - b[0]
                inc = np.where(np.arange(ng - 1) < hl - 1, rng.integers(1, 16, ng - 1) / 32.0,
                               spread + rng.integers(0, 3, ng - 1) / 32.0)      # (0: the boundary ties exactly)
                a = 0.5 + np.concatenate([[0.0], np.cumsum(inc)])
            else:
                b = 10.0 ** rng.uniform(-3, 3) * 1.3 ** np.arange(ng)
                spread = b[-1] - b[0]
                inc = np.where(np.arange(ng - 1) < hl - 1, spread * rng.uniform(0.05, 0.9, ng - 1),
                               spread * (1.0 + rng.uniform(0.01, 2.0, ng - 1)))
                a = b[0] * rng.uniform(0.1, 10.0) + np.concatenate([[0.0], np.cumsum(inc)])
            k[w, :, l, 0] = a
            k[w, :, l, 1] = b
    dkdT = k * rng.uniform(-0.01, 0.01, size=k.shape)
    c = dict(tab=dict(DELG=delg), amount=np.ones((2, nlay)))
    st, st0 = _check(mods, c, k, dkdT)
    assert st["sorted_folds"] > reps * nlay // 4
    assert st["handed_over"] < reps * nlay
