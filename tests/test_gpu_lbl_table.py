"""GPU: line-by-line tables (ILBL = LINE_BY_LINE_TABLES) on the device -- ansb200_lbl_table_opacity against the
golden vectors of the live reference and the oracle, the engine end to end on a 4-D table, and the lblconv[g]
operators through ansb200_convolve.  Tolerances: 1e-13 on the opacities (CUDA exp/log vs libm differ by an ulp;
sums and products are in the reference's order), bit-exact for the line-shape operators."""
import numpy as np
import pytest

from tests.golden_util import load
from tests.util import relerr, colerr, cpu

pytestmark = pytest.mark.gpu


def _device_opacity(K, P, T, press, temp, amount, grad):
    from archnemesis_dist_b200 import ops, plan
    tab = ops.Table(K.reshape(K.shape[0], 1, *K.shape[1:]))
    dplan = ops.LblDevicePlan(plan.klbl_plan(P, T, press, temp, grad))
    out = ops.lbl_table_opacity(tab, dplan, ops.to_dev(amount), grad)
    out = tuple(cpu(o) for o in out) if grad else cpu(out)
    tab.close()
    return out


@pytest.mark.parametrize("tag", ["a", "b"])
def test_lbl_table_opacity_goldens(tag):
    g = load("lbl_table.npz")
    K, P, T, press, temp, amount = (g["%s_%s" % (tag, n)] for n in ("K", "PRESS", "TEMP", "press", "temp", "amount"))
    tau = _device_opacity(K, P, T, press, temp, amount, False)
    assert relerr(tau, g[tag + "_tau"]) < 1e-13
    taug, dk = _device_opacity(K, P, T, press, temp, amount, True)
    assert relerr(taug, g[tag + "_taug"]) < 1e-13
    ngas = K.shape[3]
    assert relerr(dk[..., :ngas], g[tag + "_dk"][..., :ngas]) < 1e-13
    assert colerr(dk[..., ngas], g[tag + "_dk"][..., ngas]) < 1e-13
    # zero / linear-branch entries are exact
    z = g[tag + "_dk"][..., :ngas] <= 0.0
    assert np.array_equal(dk[..., :ngas][z], g[tag + "_dk"][..., :ngas][z])


@pytest.mark.parametrize("ngas", [1, 2, 4, 6, 9, 16])
def test_lbl_table_opacity_matches_oracle(ngas):
    from oracle import oracle as orc
    rng = np.random.default_rng(100 + ngas)
    nw, npg, ntg, nlay = 517, 9, 7, 61
    K = np.exp(rng.uniform(-60.0, -38.0, size=(nw, npg, ntg, ngas)))
    K[rng.uniform(size=K.shape) < 0.03] = 0.0
    P = np.exp(np.linspace(np.log(1e-7), np.log(20.0), npg)).astype(np.float32)
    T = np.linspace(70.0, 400.0, ntg).astype(np.float32)
    press = np.exp(rng.uniform(np.log(1e-8), np.log(50.0), nlay))
    temp = rng.uniform(50.0, 450.0, nlay)
    temp[7] = float(T[0])
    amount = np.exp(rng.uniform(40.0, 58.0, size=(ngas, nlay)))
    ref = orc.lbl_table_opacity(K, P, T, press, temp, amount)
    assert relerr(_device_opacity(K, P, T, press, temp, amount, False), ref) < 1e-13
    rt, rdk = orc.lbl_table_opacity(K, P, T, press, temp, amount, want_grad=True)
    taug, dk = _device_opacity(K, P, T, press, temp, amount, True)
    assert relerr(taug, rt) < 1e-13 and relerr(dk[..., :ngas], rdk[..., :ngas]) < 1e-13
    assert colerr(dk[..., ngas], rdk[..., ngas]) < 1e-13


def test_lbl_table_requires_single_g_ordinate():
    from archnemesis_dist_b200 import ops, plan
    K = np.full((4, 2, 3, 3, 2), 1e-20)
    tab = ops.Table(K)
    dplan = ops.LblDevicePlan(plan.klbl_plan(np.array([1e-3, 1e-1, 1.0]), np.array([100.0, 200.0, 300.0]),
                                             np.array([0.01]), np.array([150.0]), False))
    with pytest.raises(ValueError):
        ops.lbl_table_opacity(tab, dplan, ops.to_dev(np.ones((2, 1))), False)
    tab.close()


@pytest.mark.parametrize("mode", ["thermal", "transmission"])
def test_engine_on_a_line_by_line_table(mode):
    """HotPath on a 4-D table: calc_klblg + gas sum + radiance + projection + lblconvg (Gaussian ILS) against the
    oracle-backed engine of tests/cpu_engine.py, and the convolved block bit for bit against the oracle's operator."""
    from archnemesis_dist_b200 import engine, plan, synthetic as syn
    from oracle import oracle as orc
    from tests import cpu_engine
    c = syn.make_fm_case(nwave=240, ng=1, ngas=3, nlay=20, npro=20, nx=9, nvmr=4, seed=77, tsurf=150.0)
    tab = c["tab"]
    K4 = np.ascontiguousarray(tab["K"][:, 0])
    delg = np.array([1.0])
    md = engine.THERMAL if mode == "thermal" else engine.TRANSMISSION
    ev = engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                           NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"],
                           EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"],
                           TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"], mode=md)
    M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    hp = engine.HotPath(K4, tab["PRESS"], tab["TEMP"], delg, tab["WAVE"])
    ref = cpu_engine.HotPath(K4, tab["PRESS"], tab["TEMP"], delg, tab["WAVE"])
    assert hp.lbl_table and hp.NG == 1
    spec, dx, dts = (cpu(t) for t in hp.forward_jacobian(ev, M))
    rs, rdx, rdts = ref.forward_jacobian(ev, M)
    assert relerr(spec, rs) < 1e-9 and colerr(dx, rdx) < 1e-9 and colerr(dts, rdts) < 1e-9
    import dataclasses
    ev0 = dataclasses.replace(ev, SOL_ANG=np.array([100.0]), EMISS_ANG=np.array([10.0]))
    assert relerr(cpu(hp.cirsrad(ev0)), ref.cirsrad(ev0)) < 1e-9
    wave = tab["WAVE"]
    step = wave[1] - wave[0]
    vconv = np.linspace(wave[30], wave[-31], 11)
    op = plan.lbl_conv_operator(wave, vconv, 6.0 * step, plan.ILS_GAUSSIAN)
    out = cpu(hp.forward_jacobian_conv(ev, M, hp.conv_operator(op), 2, 0.5))
    block = np.concatenate([spec[:, :1], dx[:, 0, :]], axis=1)
    block[:, 1 + 2] = dts[:, 0]
    block = block * 0.5
    assert np.array_equal(out[:, 0], orc.apply_conv(op, block[:, 0]))
    assert np.array_equal(out[:, 1:], orc.apply_conv(op, block[:, 1:]))
    hp.close()


def test_lblconv_operators_goldens():
    """ansb200_convolve with the lblconv / lblconvg operators: bit for bit the live reference's outputs."""
    from archnemesis_dist_b200 import ops, plan
    g = load("lbl_table.npz")
    cw, cy, cg, vconv = g["lc_wave"], g["lc_y"], g["lc_grad"], g["lc_vconv"]
    block = ops.to_dev(np.concatenate([cy[:, None], cg], axis=1))
    for tag in ("sq", "tr", "ga", "ha", "ip"):
        op = plan.lbl_conv_operator(cw, vconv, float(g["lc_%s_fwhm" % tag]), int(g["lc_%s_ishape" % tag]))
        out = cpu(ops.convolve(ops.ConvOperator(op), block))
        assert np.array_equal(out[:, 0], g["lc_%s_y" % tag]), tag
        assert np.array_equal(out[:, 1:], g["lc_%s_g" % tag]), tag
        op0 = plan.lbl_conv_operator(cw, vconv, float(g["lc_%s_fwhm" % tag]), int(g["lc_%s_ishape" % tag]), grad=False)
        out0 = cpu(ops.convolve(ops.ConvOperator(op0), block[:, :1].contiguous()))
        assert np.array_equal(out0[:, 0], g["lc_%s_y0" % tag], equal_nan=True), tag
    op = plan.lbl_conv_operator(cw, vconv, -1.0, NFIL=g["lc_nfil"], VFIL=g["lc_vfil"], AFIL=g["lc_afil"])
    out = cpu(ops.convolve(ops.ConvOperator(op), block))
    assert np.array_equal(out[:, 0], g["lc_fil_y"]) and np.array_equal(out[:, 1:], g["lc_fil_g"])


@pytest.mark.parametrize("mode", ["transmission", "thermal"])
def test_engine_lbl_table_limb_paths(mode):
    """Solar-occultation / limb geometry on a line-by-line table (the reference's mars_solocc example is this
    shape: ILBL = 2, one path per tangent height): six ragged limb paths, NG = 1, through the warp-per-path
    kernels, against the oracle-backed engine."""
    from archnemesis_dist_b200 import engine, plan, synthetic as syn
    from tests import cpu_engine
    c = syn.make_fm_case(nwave=150, ng=1, ngas=3, nlay=24, npro=24, nx=8, nvmr=4, seed=91, tsurf=-1.0)
    tab = c["tab"]
    K4, delg = np.ascontiguousarray(tab["K"][:, 0]), np.array([1.0])
    nlay, npath = 24, 6
    nlm = 2 * nlay
    layinc = np.zeros((nlm, npath), np.int32)
    scale = np.zeros((nlm, npath))
    nlayin = np.zeros(npath, np.int32)
    for p in range(npath):
        t = 3 * p + 1
        seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay))
        nlayin[p] = len(seq)
        layinc[:len(seq), p] = seq
        scale[:len(seq), p] = 1.0 + 15.0 / (1.0 + np.abs(np.array(seq) - t))
    md = engine.THERMAL if mode == "thermal" else engine.TRANSMISSION
    ev = engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                           NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=layinc, SCALE=scale, NLAYIN=nlayin,
                           EMTEMP=c["temp"][layinc], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"],
                           dtaucon=c["dtaucon"], TSURF=-1.0, EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"], mode=md)
    M = plan.fold_projection(c["xmap"], layinc, nlayin, c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    hp = engine.HotPath(K4, tab["PRESS"], tab["TEMP"], delg, tab["WAVE"])
    ref = cpu_engine.HotPath(K4, tab["PRESS"], tab["TEMP"], delg, tab["WAVE"])
    spec, dx, dts = (cpu(t) for t in hp.forward_jacobian(ev, M))
    rs, rdx, rdts = ref.forward_jacobian(ev, M)
    assert spec.shape == (150, npath) and dx.shape == (150, npath, 8)
    assert relerr(spec, rs) < 1e-9
    for p in range(npath):
        assert colerr(dx[:, p], rdx[:, p]) < 1e-9, p
    hp.close()
