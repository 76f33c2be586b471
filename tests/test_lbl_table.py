"""CPU: line-by-line tables (ILBL = LINE_BY_LINE_TABLES) -- the oracle, the host plan and the lblconv operator
against the golden vectors captured from the live reference (tests/golden/lbl_table.npz, oracle/make_golden.py:
Spectroscopy_0.calc_klbl / calc_klblg, the LBL branch of calculate_gaseous_line_opacity, Measurement_0.lblconv[g])."""
import numpy as np
import pytest

from archnemesis_dist_b200 import plan
from oracle import oracle as orc
from tests.golden_util import load


def numpy_sum_order(terms):
    """np.sum over the last axis the way numpy's pairwise kernel does it (what klbl.cu implements)."""
    n = terms.shape[-1]
    if n < 8:
        r = np.zeros(terms.shape[:-1])
        for i in range(n):
            r = r + terms[..., i]
        return r
    acc = [terms[..., j].copy() for j in range(8)]
    m8 = n - n % 8
    for i0 in range(8, m8, 8):
        for j in range(8):
            acc[j] = acc[j] + terms[..., i0 + j]
    r = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]))
    for i in range(m8, n):
        r = r + terms[..., i]
    return r


def plan_emulation(K, hp, amount, grad):
    """The device kernel's arithmetic (csrc/klbl.cu) in numpy, driven by the host plan."""
    nw, npg, ntg, ngas = K.shape
    planes = K.reshape(nw, npg * ntg, ngas)
    nlay = len(hp["corner"])
    tau = np.zeros((nw, 1, nlay))
    dk = np.zeros((nw, 1, nlay, ngas + 1))
    for l in range(nlay):
        c00, c01, c10, c11 = (planes[:, c, :] for c in hp["corner"][l])
        w0, w1, w2, w3 = hp["w4"][l]
        omv, v, d1, d2 = hp["omv"][l], hp["vv"][l], hp["du1dt"][l], hp["du2dt"][l]
        pos = (c00 > 0) & (c01 > 0) & (c10 > 0) & (c11 > 0)
        neg = (c00 <= 0) & (c01 <= 0) & (c10 <= 0) & (c11 <= 0)
        k = np.zeros((nw, ngas))
        dkdT = np.zeros((nw, ngas))
        with np.errstate(all="ignore"):
            for mask, f in ((pos, np.log), (neg, lambda a: a)):
                a00, a01, a10, a11 = f(c00), f(c01), f(c10), f(c11)
                x = ((w0 * a00 + w1 * a10) + w2 * a11) + w3 * a01
                kv = np.exp(x) if f is np.log else x
                s = (((-a00 * omv) * d1 - (a10 * v) * d2) + (a11 * v) * d2) + (a01 * omv) * d1
                k[mask] = kv[mask]
                dkdT[mask] = (kv * s if f is np.log else s)[mask]
        tau[:, 0, l] = numpy_sum_order(k * amount[:, l])
        dk[:, 0, l, :ngas] = k
        for i in range(ngas):
            dk[:, 0, l, ngas] = dk[:, 0, l, ngas] + dkdT[:, i] * amount[i, l]
    return (tau, dk) if grad else tau


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_and_plan_against_reference_goldens(tag):
    g = load("lbl_table.npz")
    K, P, T, press, temp, amount = (g["%s_%s" % (tag, n)] for n in ("K", "PRESS", "TEMP", "press", "temp", "amount"))
    assert np.array_equal(orc.calc_klbl(K, P, T, press, temp), g[tag + "_k"])
    kg, dkdT = orc.calc_klbl(K, P, T, press, temp, want_grad=True)
    assert np.array_equal(kg, g[tag + "_kg"]) and np.array_equal(dkdT, g[tag + "_dkdT"])
    assert np.array_equal(orc.lbl_table_opacity(K, P, T, press, temp, amount), g[tag + "_tau"])
    t, dk = orc.lbl_table_opacity(K, P, T, press, temp, amount, want_grad=True)
    assert np.array_equal(t, g[tag + "_taug"]) and np.array_equal(dk, g[tag + "_dk"])
    # the host plan + the kernel's arithmetic: bit for bit the reference
    assert np.array_equal(plan_emulation(K, plan.klbl_plan(P, T, press, temp, False), amount, False), g[tag + "_tau"])
    t, dk = plan_emulation(K, plan.klbl_plan(P, T, press, temp, True), amount, True)
    assert np.array_equal(t, g[tag + "_taug"]) and np.array_equal(dk, g[tag + "_dk"])


def test_klblg_first_temperature_node_wraps_like_the_reference():
    """calc_klblg leaves it = -1 for a layer on the first temperature node: the plan addresses the last and the
    first temperature planes, calc_klbl (no gradient) clamps to the first bracket instead."""
    g = load("lbl_table.npz")
    P, T = g["a_PRESS"], g["a_TEMP"]
    nt = len(T)
    hp_g = plan.klbl_plan(P, T, np.array([1e-3]), np.array([float(T[0])]), True)
    hp_0 = plan.klbl_plan(P, T, np.array([1e-3]), np.array([float(T[0])]), False)
    assert hp_g["corner"][0, 0] % nt == nt - 1 and hp_g["corner"][0, 1] % nt == 0
    assert hp_0["corner"][0, 0] % nt == 0 and hp_0["corner"][0, 1] % nt == 1


def test_pressure_dependent_temperature_grids():
    """NT < 0 tables carry one temperature grid per pressure level (TEMP[NP,NT]): the two levels bracket the
    temperature separately.  Oracle against the plan emulation (no reference golden: the reference's own read
    path for such tables needs HDF5)."""
    rng = np.random.default_rng(3)
    nw, npg, ntg, ngas, nlay = 4, 5, 4, 2, 12
    K = np.exp(rng.uniform(-60, -40, (nw, npg, ntg, ngas)))
    P = np.exp(np.linspace(np.log(1e-5), np.log(5.0), npg))
    T = np.stack([np.linspace(80.0 + 7 * i, 300.0 + 11 * i, ntg) for i in range(npg)])
    press = np.exp(rng.uniform(np.log(1e-6), np.log(9.0), nlay))
    temp = rng.uniform(60.0, 400.0, nlay)
    amount = np.exp(rng.uniform(40, 55, (ngas, nlay)))
    for grad in (False, True):
        ref = orc.lbl_table_opacity(K, P, T, press, temp, amount, want_grad=grad)
        got = plan_emulation(K, plan.klbl_plan(P, T, press, temp, grad), amount, grad)
        if grad:
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        else:
            assert np.array_equal(got, ref)


def test_lbl_conv_operator_goldens():
    g = load("lbl_table.npz")
    cw, cy, cg, vconv = g["lc_wave"], g["lc_y"], g["lc_grad"], g["lc_vconv"]
    for tag in ("sq", "tr", "ga", "ha", "ip"):
        fwhm, ishape = float(g["lc_%s_fwhm" % tag]), int(g["lc_%s_ishape" % tag])
        op = plan.lbl_conv_operator(cw, vconv, fwhm, ishape)
        assert np.array_equal(orc.apply_conv(op, cy), g["lc_%s_y" % tag]), tag
        assert np.array_equal(orc.apply_conv(op, cg), g["lc_%s_g" % tag]), tag
        op0 = plan.lbl_conv_operator(cw, vconv, fwhm, ishape, grad=False)
        with np.errstate(all="ignore"):
            assert np.array_equal(orc.apply_conv(op0, cy), g["lc_%s_y0" % tag], equal_nan=True), tag
    op = plan.lbl_conv_operator(cw, vconv, -1.0, NFIL=g["lc_nfil"], VFIL=g["lc_vfil"], AFIL=g["lc_afil"])
    assert np.array_equal(orc.apply_conv(op, cy), g["lc_fil_y"]) and np.array_equal(orc.apply_conv(op, cg), g["lc_fil_g"])
    with pytest.raises(ZeroDivisionError):                       # Hanning carries no weight in the reference
        plan.lbl_conv_operator(cw, vconv, 0.1, plan.ILS_HANNING)


def test_lbl_conv_interpolation_clamps_outside_the_grid():
    """np.interp (lblconv with FWHM == 0) returns the end values outside the calculation grid."""
    x = np.linspace(10.0, 11.0, 21)
    y = np.random.default_rng(0).normal(size=(21, 3))
    vc = np.array([9.5, 10.0, 10.26, 10.5, 11.0, 11.7])
    op = plan.lbl_conv_operator(x, vc, 0.0)
    got = orc.apply_conv(op, y)
    for c in range(3):
        assert np.array_equal(got[:, c], np.interp(vc, x, y[:, c]))


def test_grouped_plan_equals_scalar_plan():
    """The vectorised host plan (layers grouped by which coordinate is clamped, arrays of the grids' dtypes) is bit
    for bit the layer-by-layer restatement of the reference's scalar arithmetic, for float32 and float64 grids, on
    nodes, at and beyond the edges."""
    rng = np.random.default_rng(0)
    for trial in range(40):
        gdt = (np.float32, np.float64)[trial % 2]
        npg, ntg = int(rng.integers(3, 12)), int(rng.integers(3, 9))
        P = np.exp(np.sort(rng.uniform(-16, 3, npg))).astype(gdt)
        T = np.sort(rng.uniform(60, 400, ntg)).astype(gdt)
        press = np.exp(rng.uniform(-18, 5, 30))
        temp = rng.uniform(40, 450, 30)
        press[0], press[1], press[2] = float(P[0]), float(P[-1]), float(P[1])
        temp[0], temp[3], temp[4], temp[5] = float(T[0]), float(T[0]), float(T[-1]), float(T[1])
        for grad in (False, True):
            a = plan._klbl_plan_grouped(np.log(P), T, press, temp, grad)
            b = plan._klbl_plan_scalar(np.log(P), T, press, temp, grad)
            for k in a:
                assert np.array_equal(a[k], b[k]), (trial, grad, k)
