"""CPU: host-side plans (archnemesis_dist_b200/plan.py) against the oracle and, when the reference
tree is present, against the reference itself."""
import numpy as np
import pytest

from archnemesis_dist_b200 import plan, synthetic as syn
from oracle import oracle as orc
from tests.util import relerr, colerr


def _grids(dtype):
    tab = syn.make_ktable(3, 4, 9, 7, 2, seed=1)
    return tab["PRESS"].astype(dtype), tab["TEMP"].astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("grad", [False, True])
def test_kinterp_plan_matches_oracle_plan(dtype, grad):
    P, T = _grids(dtype)
    rng = np.random.default_rng(0)
    press = np.concatenate([np.exp(rng.uniform(-17, 3, 40)), P[[0, 3, -1]].astype(np.float64), [1e-9, 50.0]])
    temp = np.concatenate([rng.uniform(50, 330, 40), T[[0, 2, -1]].astype(np.float64), [10.0, 500.0]])
    a = plan.kinterp_plan(P, T, press, temp, grad)
    ip, it, w4, omv, vv, dudt = orc.kinterp_plan(P, T, press, temp, grad)
    assert np.array_equal(a["ip_lo"], ip) and np.array_equal(a["it_lo"], it)
    for x, y in ((a["w4"], w4), (a["omv"], omv), (a["vv"], vv), (a["dudt"], dudt)):
        assert np.array_equal(x, y)
    assert a["ip_lo"].min() >= 0 and a["ip_lo"].max() <= len(P) - 2
    assert a["it_lo"].min() >= 0 and a["it_lo"].max() <= len(T) - 2


def test_float32_grids_change_the_weights():
    """with float32 PRESS/TEMP the clamped branches are evaluated in float32."""
    P32, T32 = _grids(np.float32)
    press = np.array([1e-9, 0.3, 50.0])
    temp = np.array([10.0, 150.0, 500.0])
    a = plan.kinterp_plan(P32, T32, press, temp, True)
    b = plan.kinterp_plan(P32.astype(np.float64), T32.astype(np.float64), press, temp, True)
    assert np.array_equal(a["ip_lo"], b["ip_lo"])
    assert not np.array_equal(a["w4"][1], b["w4"][1])      # float32 log of the bracket pressures


def test_overlap_tables_dtype_and_seq_flag():
    g, dg = syn.gauss_legendre_01(20)
    w, e, seq = plan.overlap_tables(dg)
    ow, oe = orc.overlap_tables(dg)
    assert np.array_equal(w, ow) and np.array_equal(e, oe) and not seq
    assert e[-1] == 1.0 and w.dtype == np.float64
    w64, e64, _ = plan.overlap_tables(dg.astype(np.float64))
    assert not np.array_equal(e, e64)
    # a quadrature whose largest weight product exceeds the narrowest bin needs the sequential scan
    bad = np.array([0.01, 0.49, 0.49, 0.01])
    assert plan.overlap_tables(bad)[2]


def test_fold_projection_equals_map2pro_map2xvec():
    c = syn.make_fm_case(nwave=7, ng=4, ngas=2, nlay=12, nvmr=5, ndust=2, npro=15, nx=20, seed=4)
    rng = np.random.default_rng(1)
    layinc = np.stack([np.arange(11, -1, -1), np.r_[np.arange(11, 5, -1), np.zeros(6, int)]], axis=1).astype(np.int32)
    nlayin = np.array([12, 6], np.int32)
    xmap = c["xmap"].copy()
    xmap[2, c["NVMR"] + 1, :] = rng.uniform(size=15)
    xmap[5, c["NVMR"] + 3, :] = 0.25          # para-H2 slot: the reference re-uses the previous product
    dspec = rng.normal(size=(7, c["NPAR"], 12, 2))
    dspec[:, :, 6:, 1] = 0.0
    M = plan.fold_projection(xmap, layinc, nlayin, c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    got = np.einsum("wpe,pex->wpx", np.transpose(dspec, (0, 3, 1, 2)).reshape(7, 2, -1), M)
    d2 = orc.map2pro(dspec, 7, c["NVMR"], c["NDUST"], c["NPRO"], 2, nlayin, layinc, c["DTE"], c["DAM"], c["DCO"],
                     INCPAR=orc.included_params(xmap))
    assert colerr(got, orc.map2xvec(d2, xmap)) < 1e-13
    assert plan.included_params(xmap) == orc.included_params(xmap)


@pytest.mark.reference
def test_oracle_matches_live_reference_functions():
    """The oracle against the unmodified reference functions on fresh seeded inputs."""
    import sys
    from oracle.ref_import import import_reference
    ans = import_reference()
    fm = sys.modules["archnemesis.ForwardModel_0"]
    c = syn.make_fm_case(nwave=4, ng=20, ngas=4, nlay=11, npro=11, nx=6, nvmr=5, seed=77, zero_fraction=0.1)
    tab = c["tab"]
    S = ans.Spectroscopy_0(ILBL=0)
    for k in ("K", "PRESS", "TEMP", "G_ORD", "DELG", "WAVE", "NWAVE", "NG", "NP", "NT", "NGAS"):
        setattr(S, k, tab[k])
    press, temp = c["press"].copy(), c["temp"].copy()
    press[0], temp[-1] = 30.0, 20.0
    kr, dr = S.calc_kg(len(press), press, temp)
    ko, do = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], press, temp, want_grad=True)
    assert relerr(kr, ko) < 1e-15 and relerr(dr, do) < 1e-14
    assert relerr(S.calc_k(len(press), press, temp), orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], press, temp)) < 1e-15
    tr, gr = fm.k_overlapg(tab["DELG"], kr, dr, c["amount"])
    to, go = orc.k_overlap(tab["DELG"], kr, c["amount"], dkdT=dr)
    assert np.array_equal(tr, to) and np.array_equal(gr, go)


@pytest.mark.reference
def test_oracle_oe_algebra_matches_live_reference():
    """oracle.oe_* against OptimalEstimation_0.calc_gain_matrix / calc_phiret / calc_next_xn / calc_serr of the
    live reference (OptimalEstimation_0.py:545-720) on a seeded problem."""
    from oracle.ref_import import import_reference
    from oracle import oracle as orc
    ans = import_reference()
    rng = np.random.default_rng(3)
    NY, NX = 40, 9
    OE = ans.OptimalEstimation_0(NX=NX, NY=NY)
    OE.KK = rng.normal(size=(NY, NX))
    A = rng.normal(size=(NX, NX))
    OE.SA = A @ A.T + NX * np.eye(NX)
    OE.SE = np.diag(rng.uniform(0.5, 2.0, NY))
    OE.Y, OE.YN = rng.normal(size=NY), rng.normal(size=NY)
    OE.XA, OE.XN = rng.normal(size=NX), rng.normal(size=NX)
    OE.calc_gain_matrix()
    DD, AA = orc.oe_gain_matrix(OE.KK, OE.SA, OE.SE)
    assert np.array_equal(DD, OE.DD) and np.array_equal(AA, OE.AA)
    OE.calc_phiret()
    chisq, phi = orc.oe_phiret(OE.Y, OE.YN, OE.XN, OE.XA, OE.SE, OE.SA)
    assert chisq == OE.CHISQ and phi == OE.PHI
    assert np.array_equal(orc.oe_next_xn(OE.XA, OE.XN, OE.Y, OE.YN, OE.DD, OE.AA), OE.calc_next_xn())
    OE.calc_serr()
    SM, SN, ST = orc.oe_serr(OE.DD, OE.AA, OE.SA, OE.SE)
    assert np.array_equal(SM, OE.SM) and np.array_equal(SN, OE.SN) and np.array_equal(ST, OE.ST)
