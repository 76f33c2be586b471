"""CPU: the oracle (oracle/) against the golden vectors generated from the live reference
(tests/golden/, oracle/make_golden.py).  This is the pin of the oracle that travels to the GPU box."""
import numpy as np

from oracle import oracle as orc
from tests.golden_util import load, stage_table, jupiter_objects
from tests.util import relerr, colerr


def test_kinterp_and_overlap_stage_goldens():
    g = load("stages.npz")
    tab = stage_table()["tab"]
    k = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], g["ko_press"], g["ko_temp"])
    kg, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], g["ko_press"], g["ko_temp"], want_grad=True)
    assert relerr(k, g["ko_k"]) < 1e-15 and relerr(kg, g["ko_kg"]) < 1e-15 and relerr(dkdT, g["ko_dkdT"]) < 1e-14
    assert (g["ko_k"] == 0).any()                      # dead columns exercise the cutoff short-circuits
    # the reference's own k as input -> bit-identical mixing
    assert np.array_equal(orc.k_overlap(tab["DELG"], g["ko_k"], g["ko_amount"]), g["ko_tau"])
    t, d = orc.k_overlap(tab["DELG"], g["ko_kg"], g["ko_amount"], dkdT=g["ko_dkdT"])
    assert np.array_equal(t, g["ko_taug"]) and np.array_equal(d, g["ko_dk"])
    # float64 DELG (HDF5 path) moves the bin edges: the oracle follows the dtype
    t64 = orc.k_overlap(tab["DELG"].astype(np.float64), g["ko_k"], g["ko_amount"])
    assert np.array_equal(t64, g["ko_tau_f64delg"])
    assert relerr(g["ko_tau"], g["ko_tau_f64delg"]) > 1e-8      # the float32 quirk is visible


def test_thermal_stage_goldens():
    g = load("stages.npz")
    z, em = np.zeros(5), np.full(5, 0.9)
    wave = g["th_wave"]
    for tag, ispace, wv, tsurf, emis in (("a", 0, wave, -1.0, z), ("b", 0, wave, 150.0, em), ("c", 1, 1e4 / wave, 150.0, em)):
        s = orc.thermal(ispace, wv, g["th_tau"], None, g["th_emtemp"], g["th_empress"], tsurf, emis, z, z, 100.0, 10.0)
        assert np.array_equal(s, g["th_%s_spec" % tag])
        sg, ds, dt = orc.thermalg(ispace, wv, g["th_tau"], g["th_dtau"], int(g["th_nvmr"]), g["th_emtemp"], g["th_empress"],
                                  tsurf, emis)
        assert np.array_equal(sg, g["th_%s_specg" % tag])
        assert np.array_equal(ds, g["th_%s_dspec" % tag]) and np.array_equal(dt, g["th_%s_dts" % tag])


def test_transmission_stage_golden():
    """Three ragged limb paths through the live reference's calculate_transmission_spectrum
    (ForwardModel_0.py:4104-4129) and CIRSrad's g-integration (:4504-4507)."""
    g = load("stages.npz")
    tl, tp, dtl = orc.assemble_opacity(g["tr_tau"], g["tr_dk"], g["tr_gas_slot"], 4, g["tr_dtaucon"].shape[1], g["tr_taucon"],
                                       g["tr_dtaucon"], g["tr_layinc"], g["tr_scale"])
    S, dS = orc.transmission(tp, dtl)
    s, d, _ = orc.g_integrate(S, dS, None, stage_table()["tab"]["DELG"])
    assert np.array_equal(s, g["tr_spec"]) and np.array_equal(d, g["tr_dspec"])


def test_conv_operator_goldens():
    """plan.conv_operator + oracle.apply_conv against the live reference's Measurement_0.convg / conv in
    both k-table modes (FWHM == 0: interp1d, with points on a knot and on the last knot; FWHM < 0: filter
    mean).  Bit-identical."""
    from archnemesis_dist_b200 import plan
    g = load("stages.npz")
    op0 = plan.conv_operator(g["cv_wave"], g["cv_vconv"], 0.0)
    assert np.array_equal(orc.apply_conv(op0, g["cv_y"]), g["cv_y0"])
    assert np.array_equal(orc.apply_conv(op0, g["cv_grad"]), g["cv_g0"])
    op1 = plan.conv_operator(g["cv_wave"], g["cv_vconv1"], -1.0, g["cv_nfil"], g["cv_vfil"], g["cv_afil"])
    assert np.array_equal(orc.apply_conv(op1, g["cv_y"]), g["cv_y1"])
    assert np.array_equal(orc.apply_conv(op1, g["cv_grad"]), g["cv_g1"])
    # integrated radiance over the same filters (integrate_filterg): trapezoid sum, equal to rounding
    opi = plan.filter_integral_operator(g["cv_wave"], 11, g["cv_nfil"], g["cv_vfil"], g["cv_afil"])
    assert relerr(orc.apply_conv(opi, g["cv_y"]), g["cv_yi"]) < 1e-14
    assert colerr(orc.apply_conv(opi, g["cv_grad"]), g["cv_gi"]) < 1e-14
    import pytest
    with pytest.raises(ValueError):
        plan.conv_operator(g["cv_wave"], g["cv_vconv"], 0.5)                       # convg raises for FWHM > 0
    with pytest.raises(ValueError):
        plan.conv_operator(g["cv_wave"], np.array([599.0]), 0.0)                   # interp1d bounds_error


def test_projection_stage_goldens():
    g = load("stages.npz")
    c = stage_table()
    d2 = orc.map2pro(g["mp_dspec"], 5, c["NVMR"], c["NDUST"], c["NPRO"], 1, c["NLAYIN"], c["LAYINC"], c["DTE"], c["DAM"],
                     c["DCO"], INCPAR=list(g["mp_inc"]))
    assert np.array_equal(d2, g["mp_d2"])
    assert relerr(orc.map2xvec(d2, c["xmap"]), g["mp_dx"]) < 1e-15
    assert orc.included_params(c["xmap"]) == list(g["mp_inc"])


def test_lbl_stage_golden():
    g = load("stages.npz")
    lines = {k: g["lbl_" + k] for k in ("nu", "sw", "e_lower", "stim_ref", "broadening")}
    for i, (t, p, q) in enumerate(g["lbl_pts"]):
        out = orc.lbl_absorption(g["lbl_wn"], lines, t, p, 296.0, 1.0, q, 0.98, 28.0, g["lbl_mix"])
        assert relerr(out, g["lbl_out"][i]) < 1e-13


def test_jupiter_cirsrad_golden():
    """Whole CIRSrad chain of the oracle on the arrays captured from the reference's Jupiter CIRS run."""
    from tests import cpu_engine
    from archnemesis_dist_b200.forward_model import ArrayForwardModel
    g = load("jupiter.npz")
    objs, cont = jupiter_objects(g)
    fm = ArrayForwardModel(objs, **cont)
    fm.b200_engine = cpu_engine
    spec, dspec, dts = fm.CIRSrad(return_grad=True)
    assert relerr(spec, g["ref_SPECOUT"]) < 1e-13
    assert colerr(dspec, g["ref_dSPECOUT"]) < 1e-13
    assert relerr(dts, g["ref_dTSURF"]) < 1e-13
    assert relerr(fm.CIRSrad(), g["ref_SPECOUT"]) < 1e-13
    s1, d1 = fm.b200_forward_jacobian(g["xmap"])
    assert relerr(s1, g["ref_SPECOUT"]) < 1e-13 and colerr(d1, g["ref_dSPEC1"]) < 1e-13
