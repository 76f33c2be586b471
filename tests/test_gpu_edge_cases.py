"""GPU: edge cases of the hot path -- degenerate shapes, dead gases, saturated and empty atmospheres,
unsupported shapes -- against the CPU oracle."""
import numpy as np
import pytest

from tests.util import relerr, colerr, cpu

pytestmark = pytest.mark.gpu


def _mods():
    from archnemesis_dist_b200 import ops, plan, synthetic, engine
    from oracle import oracle
    return ops, plan, synthetic, engine, oracle


@pytest.mark.parametrize("nwave,ng,ngas,nlay", [(1, 20, 2, 1), (3, 1, 3, 5), (2, 2, 2, 2), (5, 22, 7, 3), (4, 20, 9, 2)])
def test_degenerate_shapes(nwave, ng, ngas, nlay):
    ops, plan, syn, engine, orc = _mods()
    c = syn.make_fm_case(nwave=nwave, ng=ng, ngas=ngas, nlay=nlay, nvmr=ngas + 1, npro=max(nlay, 2), nx=3, seed=nwave + ng)
    tab = c["tab"]
    hp = plan.kinterp_plan(tab["PRESS"], tab["TEMP"], c["press"], c["temp"], True)
    T = ops.Table(tab["K"])
    otab = ops.OverlapTables(tab["DELG"])
    tau, dk = ops.gas_opacity(T, ops.DevicePlan(hp, True), ops.to_dev(c["amount"]), otab, True, force_seq=otab.seq)
    k, d = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    rt, rd = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=d)
    if ng == 1:
        # one g-ordinate with several gases: the single sort element reaches the first bin edge at once, the
        # reference's gdist[iloop-1] wraps to gdist[-1] and frac = 0/0: NaN, reproduced
        assert otab.seq and np.isnan(rt).all() and np.isnan(cpu(tau)).all()
        return
    assert relerr(cpu(tau), rt) < 1e-11
    for col in range(ngas + 1):
        assert colerr(cpu(dk)[..., col], rd[..., col]) < 1e-10, col


def test_all_gases_dead_and_single_live_gas():
    """Cut-off branches of k_overlap[g] (ForwardModel_0.py:5897-5937): zero tables leave tau = 0."""
    ops, plan, syn, engine, orc = _mods()
    c = syn.make_fm_case(nwave=4, ng=20, ngas=4, nlay=6, nvmr=5, npro=6, nx=3, seed=2)
    tab = c["tab"]
    for live in ([], [2], [0, 3]):
        K = np.zeros_like(tab["K"])
        for g in live:
            K[..., g] = tab["K"][..., g]
        hp = plan.kinterp_plan(tab["PRESS"], tab["TEMP"], c["press"], c["temp"], True)
        tau, dk = ops.gas_opacity(ops.Table(K), ops.DevicePlan(hp, True), ops.to_dev(c["amount"]),
                                  ops.OverlapTables(tab["DELG"]), True)
        k, d = orc.calc_k(K, tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
        rt, rd = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=d)
        assert relerr(cpu(tau), rt) < 1e-12, live
        assert np.abs(cpu(dk) - rd).max() <= 1e-12 * max(np.abs(rd).max(), 1e-300), live
        if not live:
            assert not cpu(tau).any()


def test_saturated_and_transparent_paths():
    """Optical depths of 1e4 (transmission underflows to 0) and 1e-30: no NaN, matches the oracle."""
    ops, plan, syn, engine, orc = _mods()
    for scale in (1e12, 1e-25):
        c = syn.make_fm_case(nwave=6, ng=20, ngas=3, nlay=12, nvmr=4, npro=12, nx=5, seed=8)
        c["amount"] = c["amount"] * scale
        tab = c["tab"]
        hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
        ev = engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                               NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"],
                               EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"],
                               TSURF=200.0, EMISSIVITY=np.full(6, 0.8), xfac=c["xfac"])
        spec, dspec, dts = hp.cirsrad(ev, True)
        k, d = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
        tau, dk = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=d)
        tl, tp, dtl = orc.assemble_opacity(tau, dk, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"], c["dtaucon"],
                                           c["LAYINC"], c["SCALE"])
        S, dS, dT = orc.thermal_paths(0, tab["WAVE"], tl, dtl, c["NVMR"], c["NLAYIN"], c["EMTEMP"], c["LAYPRESS"],
                                      c["LAYINC"], 200.0, np.full(6, 0.8), c["xfac"])
        s_ref, d_ref, t_ref = orc.g_integrate(S, dS, dT, tab["DELG"])
        assert np.isfinite(cpu(spec)).all() and np.isfinite(cpu(dspec)).all()
        assert relerr(cpu(spec), s_ref) < 1e-11, scale
        got = np.transpose(cpu(dspec), (0, 2, 3, 1))
        assert np.abs(got - d_ref).max() <= 1e-10 * max(np.abs(d_ref).max(), 1e-300), scale
        assert relerr(cpu(dts), t_ref) < 1e-11


def test_unsupported_shapes_raise():
    ops, plan, syn, engine, orc = _mods()
    with pytest.raises(ValueError, match="NG"):
        ops.Table(np.ones((1, 23, 2, 2, 1)))
    T = ops.Table(np.ones((2, 4, 2, 2, 16)) * 1e-22)
    c = syn.make_fm_case(nwave=2, ng=4, npress=2, ntemp=2, ngas=16, nlay=2, nvmr=17, npro=2, nx=2, seed=1)
    hp = plan.kinterp_plan(c["tab"]["PRESS"], c["tab"]["TEMP"], c["press"], c["temp"], False)
    with pytest.raises(ValueError, match="NGAS"):
        ops.gas_opacity(T, ops.DevicePlan(hp, False), ops.to_dev(c["amount"]), ops.OverlapTables(c["tab"]["DELG"]))


def test_wide_quadrature_uses_sequential_rebin():
    """A quadrature whose largest weight product spans more than one bin: the host flags it and the
    kernel runs the literal bin-edge scan; the result is the reference's (quirks included)."""
    ops, plan, syn, engine, orc = _mods()
    c = syn.make_fm_case(nwave=5, ng=4, ngas=3, nlay=4, nvmr=4, npro=4, nx=2, seed=12)
    dg = np.array([0.02, 0.47, 0.49, 0.02], np.float32)
    k, d = orc.calc_k(c["tab"]["K"], c["tab"]["PRESS"], c["tab"]["TEMP"], c["press"], c["temp"], want_grad=True)
    otab = ops.OverlapTables(dg)
    assert otab.seq
    tau, dk = ops.koverlap(ops.to_dev(k), ops.to_dev(c["amount"]), otab, dkdT=ops.to_dev(d))
    rt, rd = orc.k_overlap(dg, k, c["amount"], dkdT=d)
    assert np.array_equal(cpu(tau), rt) and np.array_equal(cpu(dk), rd)


def test_stager_chunked_upload_round_trips():
    """Large per-evaluation inputs are copied into pinned memory and sent chunk by chunk (engine._Stager): ragged
    last chunk, int32 payloads, read-only sources and buffer re-use between evaluations."""
    import torch
    from archnemesis_dist_b200.engine import _Stager
    st = _Stager()
    st.CHUNK_BYTES = 1 << 20
    rng = np.random.default_rng(2)
    a = rng.normal(size=(701, 13, 97))                     # 7.07 MB: seven chunks, the last one ragged
    d = st("x", a)
    assert torch.equal(d.cpu(), torch.from_numpy(a)) and st.bytes == a.nbytes
    b = rng.normal(size=a.shape)
    b.setflags(write=False)
    d2 = st("x", b)                                        # same slot: waits for the first copy, then overwrites
    assert torch.equal(d2.cpu(), torch.from_numpy(b.copy())) and torch.equal(d.cpu(), torch.from_numpy(a))
    i = rng.integers(-5, 5, size=(300001,)).astype(np.int32)
    di = st("i", i, torch.int32)
    assert di.dtype == torch.int32 and torch.equal(di.cpu(), torch.from_numpy(i))
    small = rng.normal(size=(17, 3))
    assert torch.equal(st("s", small).cpu(), torch.from_numpy(small))
