"""GPU parity tests: every CUDA entry point of libansb200.so against the CPU oracle on the same
seeded inputs.  Tolerances: k-interp 1e-13 (exp/log differ by an ulp between CUDA and libm),
overlap bit-exact given identical k, radiance/Jacobian 1e-9 (the BASELINE.json tolerance; the
observed error is ~1e-13), Voigt 1e-12 against SciPy."""
import numpy as np
import pytest

from tests.util import relerr, colerr, cpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def reference_tie_order():
    """The oracle's default reproduces numba's unstable quicksort for tied keys, like the kernels do when
    gradients are requested (see test_tied_keys_follow_numba_order)."""
    from oracle import oracle
    oracle.set_sort_mode(oracle.NUMBA_ORDER)
    yield
    oracle.set_sort_mode(oracle.NUMBA_ORDER)


@pytest.fixture(scope="module")
def mods():
    import torch
    from archnemesis_dist_b200 import ops, plan, synthetic
    from oracle import oracle
    return dict(torch=torch, ops=ops, plan=plan, syn=synthetic, orc=oracle)


def _case(mods, **kw):
    return mods["syn"].make_fm_case(**kw)


@pytest.mark.parametrize("want_grad", [False, True])
@pytest.mark.parametrize("zero_fraction", [0.0, 0.2])
def test_kinterp_matches_oracle(mods, want_grad, zero_fraction):
    ops, plan, orc = mods["ops"], mods["plan"], mods["orc"]
    c = _case(mods, nwave=24, ng=20, ngas=5, nlay=30, npro=30, nx=10, seed=11, zero_fraction=zero_fraction)
    tab = c["tab"]
    press, temp = c["press"].copy(), c["temp"].copy()
    press[0], press[-1], temp[3], temp[5] = 20.0, 1e-8, 50.0, 400.0   # clamp on all four sides
    hp = plan.kinterp_plan(tab["PRESS"], tab["TEMP"], press, temp, want_grad)
    T = ops.Table(tab["K"])
    out = ops.kinterp(T, ops.DevicePlan(hp, want_grad), want_grad)
    ref = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], press, temp, want_grad=want_grad)
    if want_grad:
        assert relerr(cpu(out[0]), ref[0]) < 1e-13
        assert relerr(cpu(out[1]), ref[1]) < 1e-12
    else:
        assert relerr(cpu(out), ref) < 1e-13


def test_kinterp_negative_and_mixed_corners(mods):
    ops, plan, orc, syn = mods["ops"], mods["plan"], mods["orc"], mods["syn"]
    tab = syn.make_ktable(6, 8, 6, 5, 3, seed=5)
    K = tab["K"].copy()
    rng = np.random.default_rng(0)
    K[rng.uniform(size=K.shape) < 0.15] = 0.0          # mixed corners -> 0
    K[:, :, :, :, 1] = -np.abs(K[:, :, :, :, 1])        # all corners <= 0 -> linear branch
    press = np.exp(np.linspace(1.0, -12.0, 9))
    temp = np.linspace(80.0, 280.0, 9)
    hp = plan.kinterp_plan(tab["PRESS"], tab["TEMP"], press, temp, True)
    k, d = ops.kinterp(ops.Table(K), ops.DevicePlan(hp, True), True)
    kr, dr = orc.calc_k(K, tab["PRESS"], tab["TEMP"], press, temp, want_grad=True)
    assert relerr(cpu(k), kr) < 1e-13 and relerr(cpu(d), dr) < 1e-12
    assert (kr == 0).any() and (kr < 0).any()


@pytest.mark.parametrize("ng,ngas", [(20, 6), (10, 3), (16, 2), (5, 4), (20, 1)])
@pytest.mark.parametrize("want_grad", [False, True])
def test_koverlap_matches_oracle(mods, ng, ngas, want_grad):
    ops, orc, syn, torch = mods["ops"], mods["orc"], mods["syn"], mods["torch"]
    c = _case(mods, nwave=10, ng=ng, ngas=ngas, nlay=12, npro=12, nx=6, nvmr=max(ngas, 2), seed=100 + ng + ngas,
              zero_fraction=0.15)
    tab = c["tab"]
    k, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    otab = ops.OverlapTables(tab["DELG"])
    assert not otab.seq
    kd, dd, am = ops.to_dev(k), ops.to_dev(dkdT), ops.to_dev(c["amount"])
    # default path: rebin in parallel over the sorted elements (same arithmetic, other summation order);
    # force_seq: lane-per-bin walk in the reference's order, bit-identical to the oracle
    if want_grad:
        tau, dk = ops.koverlap(kd, am, otab, dkdT=dd)
        rt, rd = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT)
        assert relerr(cpu(tau), rt) < 1e-13
        for col in range(rd.shape[-1]):
            assert colerr(cpu(dk)[..., col], rd[..., col]) < 1e-13, col
        big = np.abs(rd) > 1e-6 * np.abs(rd).max()
        assert relerr(cpu(dk)[big], rd[big]) < 1e-11
        tau2, dk2 = ops.koverlap(kd, am, otab, dkdT=dd, force_seq=True)
        assert np.array_equal(cpu(tau2), rt) and np.array_equal(cpu(dk2), rd)
    else:
        rt = orc.k_overlap(tab["DELG"], k, c["amount"])
        assert relerr(cpu(ops.koverlap(kd, am, otab)), rt) < 1e-13
        # (without gradients tied keys stay in index order: identical up to the rounding of equal terms)
        assert relerr(cpu(ops.koverlap(kd, am, otab, force_seq=True)), rt) < 1e-14


@pytest.mark.parametrize("ng,ngas", [(20, 6), (16, 3), (7, 4)])
def test_koverlap_non_monotone_gas(mods, ng, ngas):
    """k(g) of one gas scrambled (not ascending in g): the shortcut orders (row-/column-major) must not
    fire and the sort + rebin still match the oracle, with and without gradients."""
    ops, orc = mods["ops"], mods["orc"]
    c = _case(mods, nwave=12, ng=ng, ngas=ngas, nlay=10, npro=10, nx=4, nvmr=max(ngas, 2), seed=400 + ng)
    tab = c["tab"]
    k, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    perm = np.random.default_rng(5).permutation(ng)
    k[:, :, :, 1] = k[:, :, :, 1][:, perm, :]
    dkdT[:, :, :, 1] = dkdT[:, :, :, 1][:, perm, :]
    otab = ops.OverlapTables(tab["DELG"])
    kd, dd, am = ops.to_dev(k), ops.to_dev(dkdT), ops.to_dev(c["amount"])
    rt, rd = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT)
    t1, d1 = ops.koverlap(kd, am, otab, dkdT=dd)
    assert relerr(cpu(t1), rt) < 1e-13
    for col in range(rd.shape[-1]):
        assert colerr(cpu(d1)[..., col], rd[..., col]) < 1e-13, col
    t2, d2 = ops.koverlap(kd, am, otab, dkdT=dd, force_seq=True)
    assert np.array_equal(cpu(t2), rt) and np.array_equal(cpu(d2), rd)
    assert relerr(cpu(ops.koverlap(kd, am, otab)), orc.k_overlap(tab["DELG"], k, c["amount"])) < 1e-13


@pytest.mark.parametrize("want_grad", [False, True])
def test_koverlap_dominated_folds_static_orders(mods, want_grad):
    """Folds whose order is data-independent: a next gas far weaker than the running opacity gives the
    row-major order, a far stronger one the column-major order.  The kernel then applies the quadrature's
    fixed rebin matrices (ov_rebin_static) instead of sorting; results must still match the oracle, and
    the literal sequential rebin must stay bit-identical."""
    ops, orc = mods["ops"], mods["orc"]
    c = _case(mods, nwave=16, ng=20, ngas=5, nlay=9, npro=9, nx=4, nvmr=5, seed=77)
    tab = c["tab"]
    k, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    for gas, f in ((1, 1e-7), (2, 1e9), (3, 1e-9), (4, 1e11)):     # alternate weak / dominant gases
        k[..., gas] *= f
        dkdT[..., gas] *= f
    otab = ops.OverlapTables(tab["DELG"])
    kd, dd, am = ops.to_dev(k), ops.to_dev(dkdT), ops.to_dev(c["amount"])
    if want_grad:
        rt, rd = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT)
        tau, dk = ops.koverlap(kd, am, otab, dkdT=dd)
        assert relerr(cpu(tau), rt) < 1e-13
        for col in range(rd.shape[-1]):
            assert colerr(cpu(dk)[..., col], rd[..., col]) < 1e-13, col
        ts, ds = ops.koverlap(kd, am, otab, dkdT=dd, force_seq=True)
        assert np.array_equal(cpu(ts), rt) and np.array_equal(cpu(ds), rd)
    else:
        rt = orc.k_overlap(tab["DELG"], k, c["amount"])
        assert relerr(cpu(ops.koverlap(kd, am, otab)), rt) < 1e-13


def test_koverlap_float64_delg_and_ties(mods):
    """float64 DELG (HDF5 tables) changes the bin edges; exact ties (gas far below another) keep tau exact."""
    ops, orc = mods["ops"], mods["orc"]
    c = _case(mods, nwave=6, ng=20, ngas=3, nlay=8, npro=8, nx=4, nvmr=3, seed=9)
    tab = c["tab"]
    k = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"])
    k[:, :, :, 2] *= 1e-25      # a_i + b_j == a_i : whole rows tie
    dg = tab["DELG"].astype(np.float64)
    otab = ops.OverlapTables(dg)
    tau = ops.koverlap(ops.to_dev(k), ops.to_dev(c["amount"]), otab)
    assert relerr(cpu(tau), orc.k_overlap(dg, k, c["amount"])) < 1e-13


def test_tied_keys_follow_numba_order(mods):
    """Exact ties (a gas 25 orders of magnitude below another: a_i + b_j == a_i, whole rows of keys tie).
    The reference's order of tied keys is whatever numba's unstable quicksort produces; the kernel replays
    that algorithm (ov_numba_order), so even the gradient columns of the tied gas match the oracle in
    its default NUMBA_ORDER mode -- bit for bit with the reference-order rebin (force_seq)."""
    ops, orc = mods["ops"], mods["orc"]
    c = _case(mods, nwave=6, ng=20, ngas=3, nlay=8, npro=8, nx=4, nvmr=3, seed=9)
    tab = c["tab"]
    k, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    k[:, :, :, 2] *= 1e-25
    dkdT[:, :, :, 2] *= 1e-25
    otab = ops.OverlapTables(tab["DELG"])
    kd, dd, am = ops.to_dev(k), ops.to_dev(dkdT), ops.to_dev(c["amount"])
    orc.set_sort_mode(orc.NUMBA_ORDER)
    nt, nd = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT)
    orc.set_sort_mode(orc.STABLE_ORDER)
    st, sd = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT)
    assert colerr(sd[..., 2], nd[..., 2]) > 1e-6          # the two tie orders really differ for the tied gas
    tau, dk = ops.koverlap(kd, am, otab, dkdT=dd)
    assert relerr(cpu(tau), nt) < 1e-13
    for col in range(4):
        assert colerr(cpu(dk)[..., col], nd[..., col]) < 1e-13, col
    ts, ds = ops.koverlap(kd, am, otab, dkdT=dd, force_seq=True)
    assert np.array_equal(cpu(ts), nt) and np.array_equal(cpu(ds), nd)
    # without gradients the tie order is irrelevant (equal keys contribute equally wherever they fall)
    assert relerr(cpu(ops.koverlap(kd, am, otab)), nt) < 1e-13


@pytest.mark.parametrize("want_grad", [False, True])
def test_gas_opacity_fused(mods, want_grad):
    ops, plan, orc = mods["ops"], mods["plan"], mods["orc"]
    c = _case(mods, nwave=12, ng=20, ngas=6, nlay=20, npro=20, nx=8, seed=21, zero_fraction=0.1)
    tab = c["tab"]
    hp = plan.kinterp_plan(tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad)
    T = ops.Table(tab["K"])
    dp = ops.DevicePlan(hp, want_grad)
    otab = ops.OverlapTables(tab["DELG"])
    am = ops.to_dev(c["amount"])
    fused = ops.gas_opacity(T, dp, am, otab, want_grad)
    # fused == unfused on the device, bit for bit
    if want_grad:
        k, d = ops.kinterp(T, dp, True)
        tau, dk = ops.koverlap(k, am, otab, dkdT=d)
        assert np.array_equal(cpu(fused[0]), cpu(tau)) and np.array_equal(cpu(fused[1]), cpu(dk))
        kr, dr = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
        rt, rd = orc.k_overlap(tab["DELG"], kr, c["amount"], dkdT=dr)
        assert relerr(cpu(fused[0]), rt) < 1e-11
        assert colerr(cpu(fused[1]), rd) < 1e-10
    else:
        tau = ops.koverlap(ops.kinterp(T, dp, False), am, otab)
        assert np.array_equal(cpu(fused), cpu(tau))
        kr = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"])
        assert relerr(cpu(fused), orc.k_overlap(tab["DELG"], kr, c["amount"])) < 1e-11


def _radiance_inputs(mods, c, tau, dk):
    ops, torch = mods["ops"], mods["torch"]
    d = ops.to_dev
    return dict(tau=d(tau), dk=d(dk) if dk is not None else None, gas_slot=d(c["gas_slot"], torch.int32),
                taucia=d(c["taucon"]), dtaucon=d(c["dtaucon"]), layinc=d(c["LAYINC"], torch.int32), scale=d(c["SCALE"]),
                nlayin=d(c["NLAYIN"], torch.int32), emtemp=d(c["EMTEMP"]), laypress=d(c["LAYPRESS"]),
                wave=d(c["tab"]["WAVE"]), delg=d(c["tab"]["DELG"].astype(np.float64)), emissivity=d(c["EMISSIVITY"]),
                xfac=d(c["xfac"]))


@pytest.mark.parametrize("tsurf", [-1.0, 180.0])
@pytest.mark.parametrize("want_grad", [False, True])
def test_radiance_thermal(mods, tsurf, want_grad):
    ops, orc = mods["ops"], mods["orc"]
    c = _case(mods, nwave=14, ng=20, ngas=4, nlay=37, npro=37, nx=12, nvmr=6, ndust=1, seed=31, tsurf=tsurf)
    c["xfac"] = np.linspace(0.5, 2.0, 14)
    tab = c["tab"]
    kr, dr = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    tau, dk = orc.k_overlap(tab["DELG"], kr, c["amount"], dkdT=dr)
    a = _radiance_inputs(mods, c, tau, dk if want_grad else None)
    out = ops.radiance(ops.THERMAL, a["tau"], a["dk"], a["gas_slot"], a["taucia"], None, None,
                       a["dtaucon"] if want_grad else None, a["layinc"], a["scale"], a["nlayin"], a["emtemp"],
                       a["laypress"], a["wave"], a["delg"], a["emissivity"], a["xfac"], None, None, None, None,
                       c["ISPACE"], tsurf, c["NVMR"], c["NPAR"], want_grad)
    tl, tp, dtl = orc.assemble_opacity(tau, dk if want_grad else None, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"],
                                       c["dtaucon"], c["LAYINC"], c["SCALE"])
    z = np.zeros(14)
    S, dS, dT = orc.thermal_paths(c["ISPACE"], tab["WAVE"], tl, dtl, c["NVMR"], c["NLAYIN"], c["EMTEMP"], c["LAYPRESS"],
                                  c["LAYINC"], tsurf, c["EMISSIVITY"], c["xfac"], z, z, np.array([100.0]),
                                  np.array([10.0]))
    dg = tab["DELG"]
    if want_grad:
        s_ref, d_ref, t_ref = orc.g_integrate(S, dS, dT, dg)
        spec, dspec, dts = out
        assert relerr(cpu(spec), s_ref) < 1e-12
        got = np.transpose(cpu(dspec), (0, 2, 3, 1))     # -> (NWAVE,NPAR,NLAYIN,NPATH)
        for kpar in range(c["NPAR"]):
            assert colerr(got[:, kpar], d_ref[:, kpar]) < 1e-11, kpar
        big = np.abs(d_ref) > 1e-6 * np.abs(d_ref).max()
        assert relerr(got[big], d_ref[big]) < 1e-9
        assert relerr(cpu(dts), t_ref) < 1e-12
    else:
        assert relerr(cpu(out), orc.g_integrate(S, None, None, dg)) < 1e-12


@pytest.mark.parametrize("want_grad", [False, True])
def test_radiance_transmission_multipath(mods, want_grad):
    ops, orc, torch = mods["ops"], mods["orc"], mods["torch"]
    c = _case(mods, nwave=9, ng=10, ngas=3, nlay=16, npro=16, nx=5, nvmr=4, seed=41)
    rng = np.random.default_rng(3)
    # three limb-like paths of different length (down and up again), zero padded like Path_0 does
    nlm, npath = 24, 3
    layinc = np.zeros((nlm, npath), np.int32)
    scale = np.zeros((nlm, npath))
    nlayin = np.array([24, 16, 6], np.int32)
    for p, n in enumerate(nlayin):
        half = n // 2
        seq = list(range(15, 15 - half, -1))
        layinc[:n, p] = seq + seq[::-1]
        scale[:n, p] = rng.uniform(1.0, 30.0, n)
    c.update(LAYINC=layinc, SCALE=scale, NLAYIN=nlayin, EMTEMP=np.zeros((nlm, npath)))
    c["xfac"] = np.linspace(1.0, 3.0, 9)
    tab = c["tab"]
    kr, dr = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    tau, dk = orc.k_overlap(tab["DELG"], kr, c["amount"], dkdT=dr)
    tau *= 1e-3   # keep the transmissions away from underflow
    dk *= 1e-3
    a = _radiance_inputs(mods, c, tau, dk if want_grad else None)
    out = ops.radiance(ops.TRANSMISSION, a["tau"], a["dk"], a["gas_slot"], a["taucia"], None, None,
                       a["dtaucon"] if want_grad else None, a["layinc"], a["scale"], a["nlayin"], None, None, None,
                       a["delg"], None, a["xfac"], None, None, None, None, 0, -1.0, c["NVMR"], c["NPAR"], want_grad)
    tl, tp, dtl = orc.assemble_opacity(tau, dk if want_grad else None, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"],
                                       c["dtaucon"], layinc, scale)
    S, dS = orc.transmission(tp, dtl, c["xfac"])
    dg = tab["DELG"]
    if want_grad:
        s_ref, d_ref, _ = orc.g_integrate(S, dS, None, dg)
        spec, dspec, _ = out
        assert relerr(cpu(spec), s_ref) < 1e-12
        got = np.transpose(cpu(dspec), (0, 2, 3, 1))
        assert colerr(got, d_ref) < 1e-12
    else:
        assert relerr(cpu(out), orc.g_integrate(S, None, None, dg)) < 1e-12


def test_jacobian_project(mods):
    ops, plan, orc = mods["ops"], mods["plan"], mods["orc"]
    c = _case(mods, nwave=50, ng=4, ngas=2, nlay=21, npro=33, nx=70, nvmr=5, ndust=2, seed=51)
    rng = np.random.default_rng(8)
    npath, nlm = 2, 21
    layinc = np.stack([np.arange(20, -1, -1), np.r_[np.arange(20, 10, -1), np.zeros(11, int)]], axis=1).astype(np.int32)
    nlayin = np.array([21, 10], np.int32)
    dspec_ref = rng.normal(size=(50, c["NPAR"], nlm, npath))       # reference layout
    dspec_ref[:, :, 10:, 1] = 0.0
    xmap = c["xmap"].copy()
    xmap[3, c["NVMR"] + 1, :] = 0.5      # a dust parameter too
    M = plan.fold_projection(xmap, layinc, nlayin, c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    out = ops.jacobian_project(ops.to_dev(np.transpose(dspec_ref, (0, 3, 1, 2))), ops.to_dev(M))
    inc = orc.included_params(xmap)
    d2 = orc.map2pro(dspec_ref, 50, c["NVMR"], c["NDUST"], c["NPRO"], npath, nlayin, layinc, c["DTE"], c["DAM"], c["DCO"],
                     INCPAR=inc)
    ref = orc.map2xvec(d2, xmap)
    assert colerr(cpu(out), ref) < 1e-13


def test_voigt_matches_scipy(mods):
    import ctypes
    from scipy.special import voigt_profile
    from archnemesis_dist_b200 import _lib
    ops, torch = mods["ops"], mods["torch"]
    rng = np.random.default_rng(2)
    n = 200000
    dwn = np.concatenate([rng.uniform(-25, 25, n // 2), rng.normal(0, 0.01, n // 4), 10.0 ** rng.uniform(-8, 1.4, n // 4)])
    ad = 10.0 ** rng.uniform(-4, -1.5, n)
    gl = 10.0 ** rng.uniform(-6, 0.5, n)
    dwn[:5] = [0.0, 25.0, -25.0, 1e-300, 24.999]
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    d = ops.to_dev
    a, b, g = d(dwn), d(ad), d(gl)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.load().ansb200_voigt(p(a), p(b), p(g), n, p(out), None))
    ref = voigt_profile(dwn, ad / np.sqrt(2.0 * np.log(2.0)), gl)
    assert relerr(cpu(out), ref) < 1e-12


def test_lbl_absorption_matches_oracle(mods):
    ops, orc, syn, torch = mods["ops"], mods["orc"], mods["syn"], mods["torch"]
    from archnemesis_dist_b200 import lbl
    wn = np.linspace(1000.0, 1010.0, 3001)
    lines = syn.make_line_list(400, 1000.0, 1010.0, seed=4)
    pts = [(200.0, 0.1, 1.3), (296.0, 1.0, 1.0), (120.0, 1e-4, 4.0)]
    mix = np.array([0.05, 0.95])
    out = lbl.lbl_absorption(wn, lines, pts, t_ref=296.0, p_ref=1.0, abundance=0.98, mass=28.0, mix=mix)
    for i, (t, p, q) in enumerate(pts):
        ref = orc.lbl_absorption(wn, lines, t, p, 296.0, 1.0, q, 0.98, 28.0, mix)
        assert relerr(cpu(out[i]), ref) < 1e-11, i
    # a device-resident line list gives the same launch
    again = lbl.lbl_absorption(torch.from_numpy(wn).cuda(), lbl.resident_lines(lines), pts, t_ref=296.0, p_ref=1.0,
                               abundance=0.98, mass=28.0, mix=mix)
    assert torch.equal(again, out)


def test_lbl_uniform_window_classes_match_per_pair_tests(mods):
    """Lines whose 25 / 75 cm-1 windows end inside a CTA's 1024 grid points take the per-pair tests, all others the
    branch-free loops: a wide, coarse grid (every CTA spans 40 cm-1, so most lines are mixed) and a narrow, fine one
    (every line uniform) against the oracle, plus window edges that fall exactly on grid points."""
    orc, syn = mods["orc"], mods["syn"]
    from archnemesis_dist_b200 import lbl
    mix = np.array([0.2, 0.8])
    for wn in (np.linspace(900.0, 1100.0, 5121), np.linspace(1000.0, 1000.5, 2049)):
        lines = syn.make_line_list(300, wn[0], wn[-1], seed=9, pad=80.0)
        lines["nu"][:4] = [wn[100] - 25.0, wn[200] + 25.0, wn[300] - 75.0, wn[50] + 75.0]   # edges on grid points
        pts = [(220.0, 0.3, 1.0), (90.0, 2e-5, 1.5)]
        out = lbl.lbl_absorption(wn, lines, pts, t_ref=296.0, p_ref=1.0, abundance=1.0, mass=44.0, mix=mix)
        for i, (t, p, q) in enumerate(pts):
            ref = orc.lbl_absorption(wn, lines, t, p, 296.0, 1.0, q, 1.0, 44.0, mix)
            assert relerr(cpu(out[i]), ref) < 1e-11, (len(wn), i)


@pytest.mark.parametrize("fwhm", [0.0, -1.0])
def test_engine_forward_jacobian_conv(mods, fwhm):
    """HotPath.forward_jacobian_conv (spectrum + Jacobian + JSURF column + WGEOM + instrument line shape on the
    device, only [NCONV, 1+NX] returned) equals the same evaluation convolved on the host by the oracle's
    restatement of Measurement_0.convg -- bit for bit, since the operator arithmetic is the reference's."""
    from archnemesis_dist_b200 import engine
    plan, orc = mods["plan"], mods["orc"]
    c = _case(mods, nwave=40, ng=20, ngas=3, nlay=15, npro=15, nx=9, nvmr=4, seed=61, tsurf=160.0)
    tab = c["tab"]
    hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    ev = engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                           NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"],
                           EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"],
                           TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"])
    M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    wave = tab["WAVE"]
    vconv = np.linspace(wave[4], wave[-5], 7)          # filters of +-0.6 stay inside the 0.25-spaced grid
    vconv[2] = wave[11]
    if fwhm == 0.0:
        op = plan.conv_operator(wave, vconv, 0.0)
    else:
        nfil = np.full(7, 5, np.int32)
        vfil = np.stack([np.linspace(v - 0.6, v + 0.6, 5) for v in vconv], axis=1)
        afil = np.tile(np.array([0.0, 0.5, 1.0, 0.5, 0.0])[:, None], (1, 7))
        op = plan.conv_operator(wave, vconv, -1.0, nfil, vfil, afil)
    jsurf, wgeom = 4, 0.75
    out = cpu(hp.forward_jacobian_conv(ev, M, hp.conv_operator(op), jsurf, wgeom))
    spec, dx, dts = (cpu(t) for t in hp.forward_jacobian(ev, M))
    block = np.concatenate([spec[:, :1], dx[:, 0, :]], axis=1)
    block[:, 1 + jsurf] = dts[:, 0]
    block = block * wgeom
    assert np.array_equal(out[:, 0], orc.apply_conv(op, block[:, 0]))
    assert np.array_equal(out[:, 1:], orc.apply_conv(op, block[:, 1:]))
    hp.close()


def test_oe_algebra_on_device(mods):
    """oe.calc_gain_matrix / calc_phiret / calc_next_xn / calc_serr (cuBLAS / cuSOLVER through torch.linalg) against
    the numpy restatement of OptimalEstimation_0.py:545-720; tolerance scaled to the conditioning of the solve."""
    from archnemesis_dist_b200 import oe
    orc = mods["orc"]
    rng = np.random.default_rng(11)
    NY, NX = 300, 24
    KK = rng.normal(size=(NY, NX))
    A = rng.normal(size=(NX, NX))
    SA = A @ A.T + NX * np.eye(NX)
    Y, YN, XA, XN = rng.normal(size=NY), rng.normal(size=NY), rng.normal(size=NX), rng.normal(size=NX)
    SE = np.diag(rng.uniform(0.5, 2.0, NY))
    DD, AA = oe.calc_gain_matrix(KK, SA, SE)
    rDD, rAA = orc.oe_gain_matrix(KK, SA, SE)
    # two LU solves of the same NY x NY system (LAPACK on the host, cuSOLVER on the device) agree to
    # eps * cond(M); everything downstream of DD inherits that
    tol = 100 * np.finfo(float).eps * np.linalg.cond(KK @ SA @ KK.T + SE)
    assert 1e-13 < tol < 1e-8
    assert colerr(cpu(DD), rDD) < tol and colerr(cpu(AA), rAA) < tol
    xn = oe.calc_next_xn(XA, XN, Y, YN, DD, AA)
    assert colerr(cpu(xn), orc.oe_next_xn(XA, XN, Y, YN, rDD, rAA)) < tol
    # SE diagonal, and the reference's (1,1) shorthand (phiret and calc_serr(simple=True); with a (1,1) SE the
    # reference's gain matrix adds the scalar to EVERY element of kk sa kk^T, which is singular for NY > NX + 1)
    for SE2, simple in ((SE, False), (np.array([[0.7]]), True)):
        chisq, phi = oe.calc_phiret(Y, YN, XN, XA, SE2, SA)
        rc, rp = orc.oe_phiret(Y, YN, XN, XA, SE2, SA)
        assert abs(chisq - rc) < 1e-12 * abs(rc) and abs(phi - rp) < 1e-12 * abs(rp)
        SM, SN, ST = oe.calc_serr(DD, AA, SA, SE2, simple=simple)
        rSM, rSN, rST = orc.oe_serr(rDD, rAA, SA, SE2, simple=simple)
        assert colerr(cpu(SM), rSM) < tol and colerr(cpu(SN), rSN) < tol and colerr(cpu(ST), rST) < tol


def _limb_paths(nlay, npath, nlm, rng, long_path=False):
    """Ragged limb-like paths (down to a tangent layer and up again), zero padded like Path_0 pads them."""
    layinc = np.zeros((nlm, npath), np.int32)
    scale = np.zeros((nlm, npath))
    nlayin = np.zeros(npath, np.int32)
    for p in range(npath):
        t = (p * (nlay - 2)) // npath
        seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay))
        if long_path and p == npath - 1:
            seq = (seq * (nlm // len(seq) + 1))[:nlm]          # a path that revisits layers: more than 256 entries
        n = min(len(seq), nlm)
        nlayin[p] = n
        layinc[:n, p] = seq[:n]
        scale[:n, p] = rng.uniform(1.0, 20.0, n)
    return layinc, scale, nlayin


@pytest.mark.parametrize("want_grad", [False, True])
@pytest.mark.parametrize("npath,nlm,long_path", [(7, 40, False), (5, 300, True), (40, 40, False)])
def test_transmission_many_paths_warp_per_path_kernel(mods, want_grad, npath, nlm, long_path):
    """NPATH >= 4 in transmission mode takes ans_transmission_paths_kernel (one warp per path on the staged
    slabs, including more paths than warps and a path of more than 256 entries): against the oracle."""
    ops, orc, torch = mods["ops"], mods["orc"], mods["torch"]
    c = _case(mods, nwave=6, ng=20, ngas=3, nlay=18, npro=18, nx=5, nvmr=4, seed=43)
    rng = np.random.default_rng(8)
    layinc, scale, nlayin = _limb_paths(18, npath, nlm, rng, long_path)
    c.update(LAYINC=layinc, SCALE=scale, NLAYIN=nlayin, EMTEMP=np.zeros((nlm, npath)))
    c["xfac"] = np.linspace(1.0, 3.0, 6)
    tab = c["tab"]
    kr, dr = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    tau, dk = orc.k_overlap(tab["DELG"], kr, c["amount"], dkdT=dr)
    tau *= 2e-4
    dk *= 2e-4
    a = _radiance_inputs(mods, c, tau, dk if want_grad else None)
    out = ops.radiance(ops.TRANSMISSION, a["tau"], a["dk"], a["gas_slot"], a["taucia"], None, None,
                       a["dtaucon"] if want_grad else None, a["layinc"], a["scale"], a["nlayin"], None, None, None,
                       a["delg"], None, a["xfac"], None, None, None, None, 0, -1.0, c["NVMR"], c["NPAR"], want_grad)
    tl, tp, dtl = orc.assemble_opacity(tau, dk if want_grad else None, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"],
                                       c["dtaucon"], layinc, scale)
    S, dS = orc.transmission(tp, dtl, c["xfac"])
    dg = tab["DELG"]
    if want_grad:
        s_ref, d_ref, _ = orc.g_integrate(S, dS, None, dg)
        spec, dspec, _ = out
        assert relerr(cpu(spec), s_ref) < 1e-12
        got = np.transpose(cpu(dspec), (0, 2, 3, 1))
        assert colerr(got, d_ref) < 1e-12
        for p in range(npath):                    # rows past NLAYIN are written as zeros
            assert not np.any(got[:, :, nlayin[p]:, p])
    else:
        assert relerr(cpu(out), orc.g_integrate(S, None, None, dg)) < 1e-12


@pytest.mark.parametrize("want_grad", [False, True])
def test_thermal_many_limb_paths_staged_kernel(mods, want_grad):
    """NPATH >= 4 in thermal mode: the CTA walks its paths on the staged tau / dk slabs (limb emission: no
    ground term, layers seen twice)."""
    ops, orc = mods["ops"], mods["orc"]
    c = _case(mods, nwave=6, ng=20, ngas=3, nlay=18, npro=18, nx=5, nvmr=4, seed=47)
    rng = np.random.default_rng(9)
    npath, nlm = 6, 36
    layinc, scale, nlayin = _limb_paths(18, npath, nlm, rng)
    emtemp = np.zeros((nlm, npath))
    for p in range(npath):
        emtemp[:nlayin[p], p] = c["temp"][layinc[:nlayin[p], p]]
    c.update(LAYINC=layinc, SCALE=scale, NLAYIN=nlayin, EMTEMP=emtemp)
    c["xfac"] = np.linspace(0.5, 2.0, 6)
    tab = c["tab"]
    kr, dr = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    tau, dk = orc.k_overlap(tab["DELG"], kr, c["amount"], dkdT=dr)
    tau *= 2e-3
    dk *= 2e-3
    a = _radiance_inputs(mods, c, tau, dk if want_grad else None)
    sol = ops.to_dev(np.full(npath, 100.0))
    emi = ops.to_dev(np.full(npath, 10.0))
    z6 = ops.to_dev(np.zeros(6))
    out = ops.radiance(ops.THERMAL, a["tau"], a["dk"], a["gas_slot"], a["taucia"], None, None,
                       a["dtaucon"] if want_grad else None, a["layinc"], a["scale"], a["nlayin"], a["emtemp"],
                       a["laypress"], a["wave"], a["delg"], a["emissivity"], a["xfac"], None if want_grad else z6,
                       None if want_grad else z6, None if want_grad else sol, None if want_grad else emi,
                       c["ISPACE"], -1.0, c["NVMR"], c["NPAR"], want_grad)
    tl, tp, dtl = orc.assemble_opacity(tau, dk if want_grad else None, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"],
                                       c["dtaucon"], layinc, scale)
    z = np.zeros(6)
    S, dS, dT = orc.thermal_paths(c["ISPACE"], tab["WAVE"], tl, dtl, c["NVMR"], nlayin, emtemp, c["LAYPRESS"], layinc,
                                  -1.0, c["EMISSIVITY"], c["xfac"], z, z, np.full(npath, 100.0), np.full(npath, 10.0))
    dg = tab["DELG"]
    if want_grad:
        s_ref, d_ref, t_ref = orc.g_integrate(S, dS, dT, dg)
        spec, dspec, dts = out
        assert relerr(cpu(spec), s_ref) < 1e-12
        got = np.transpose(cpu(dspec), (0, 2, 3, 1))
        for kpar in range(d_ref.shape[1]):
            assert colerr(got[:, kpar], d_ref[:, kpar]) < 1e-11, kpar
        assert relerr(cpu(dts), t_ref) < 1e-12
    else:
        assert relerr(cpu(out), orc.g_integrate(S, None, None, dg)) < 1e-12


@pytest.mark.parametrize("nlay,npath,ngas,tsurf,kind", [(13, 1, 4, 150.0, "nadir"), (37, 2, 5, -1.0, "mixed"),
                                                        (75, 3, 7, 200.0, "mixed"), (100, 1, 6, -1.0, "nadir"),
                                                        (112, 2, 3, -1.0, "limb")])
def test_thermal_one_to_three_paths_warp_per_path_kernel(mods, nlay, npath, ngas, tsurf, kind):
    """NPATH < 4 in thermal mode with gradients takes ans_thermal_nadir_kernel (a warp per (wavenumber, path), the
    g-ordinate's operands streamed through per-warp shared memory): every chunk length (1, 3, 5, 7 visits per lane),
    nadir paths with and without a surface, limb paths that cross layers twice next to a nadir path, an odd number of
    layers (8-byte copies), NGAS + 1 even and at its maximum of 8 -- against the oracle."""
    ops, orc = mods["ops"], mods["orc"]
    nw = 5
    c = _case(mods, nwave=nw, ng=20, ngas=ngas, nlay=nlay, npro=nlay, nx=5, nvmr=ngas + 1, seed=200 + nlay, tsurf=tsurf)
    rng = np.random.default_rng(nlay)
    nlm = 2 * nlay if kind != "nadir" else nlay
    layinc = np.zeros((nlm, npath), np.int32)
    scale = np.zeros((nlm, npath))
    nlayin = np.zeros(npath, np.int32)
    for p in range(npath):
        if kind == "nadir" or (kind == "mixed" and p == 0):
            seq = list(range(nlay - 1, -1, -1))                       # top to bottom: ends on the ground
        else:
            t = int(rng.integers(0, nlay - 1))
            seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay))
        n = len(seq)
        nlayin[p] = n
        layinc[:n, p] = seq
        scale[:n, p] = rng.uniform(1.0, 4.0, n)
    emtemp = np.zeros((nlm, npath))
    for p in range(npath):
        emtemp[:nlayin[p], p] = c["temp"][layinc[:nlayin[p], p]]
    c.update(LAYINC=layinc, SCALE=scale, NLAYIN=nlayin, EMTEMP=emtemp)
    c["xfac"] = np.linspace(0.5, 2.0, nw)
    tab = c["tab"]
    kr, dr = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    tau, dk = orc.k_overlap(tab["DELG"], kr, c["amount"], dkdT=dr)
    tau *= 2e-3
    dk *= 2e-3
    a = _radiance_inputs(mods, c, tau, dk)
    spec, dspec, dts = ops.radiance(ops.THERMAL, a["tau"], a["dk"], a["gas_slot"], a["taucia"], None, None, a["dtaucon"],
                                    a["layinc"], a["scale"], a["nlayin"], a["emtemp"], a["laypress"], a["wave"], a["delg"],
                                    a["emissivity"], a["xfac"], None, None, None, None, c["ISPACE"], tsurf, c["NVMR"],
                                    c["NPAR"], True)
    tl, tp, dtl = orc.assemble_opacity(tau, dk, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"], c["dtaucon"], layinc, scale)
    z = np.zeros(nw)
    S, dS, dT = orc.thermal_paths(c["ISPACE"], tab["WAVE"], tl, dtl, c["NVMR"], nlayin, emtemp, c["LAYPRESS"], layinc,
                                  tsurf, c["EMISSIVITY"], c["xfac"], z, z, np.full(npath, 100.0), np.full(npath, 10.0))
    s_ref, d_ref, t_ref = orc.g_integrate(S, dS, dT, tab["DELG"])
    assert relerr(cpu(spec), s_ref) < 1e-12
    got = np.transpose(cpu(dspec), (0, 2, 3, 1))
    for kpar in range(d_ref.shape[1]):
        assert colerr(got[:, kpar], d_ref[:, kpar]) < 1e-11, kpar
    assert relerr(cpu(dts), t_ref) < 1e-12 or float(np.abs(t_ref).max()) == 0.0
    for p in range(npath):                                            # rows past NLAYIN are written as zeros
        assert float(np.abs(got[:, :, nlayin[p]:, p]).max(initial=0.0)) == 0.0


def test_float32_table_storage_variants():
    """ansb200_table_create_ex: K as float32 is lossless for .kta data (bit-identical k-interp and fused gas opacity);
    K and ln K as float32 -- the FP32 k-interp variant -- stays within 2e-5 of the float64 table; a table that is not
    float32-representable is refused for the lossless variant."""
    import torch
    from archnemesis_dist_b200 import ops, plan, synthetic
    c = synthetic.make_fm_case(nwave=24, ng=20, ngas=6, nlay=60, npro=60, nx=8, nvmr=8, seed=4, zero_fraction=0.1)
    tab = c["tab"]
    hp = plan.kinterp_plan(tab["PRESS"], tab["TEMP"], c["press"], c["temp"], True)
    dp = ops.DevicePlan(hp, True)
    otab = ops.OverlapTables(tab["DELG"])
    am = ops.to_dev(c["amount"])
    out = {}
    for st in ("f64", "k32", "f32"):
        T = ops.Table(tab["K"], st)
        assert T.nbytes == {"f64": 16, "k32": 12, "f32": 8}[st] * tab["K"].size
        k, d = ops.kinterp(T, dp, True)
        tau, dk = ops.gas_opacity(T, dp, am, otab, True)
        out[st] = [cpu(x) for x in (k, d, tau, dk)]
        T.close()
    for a, b in zip(out["k32"], out["f64"]):
        assert np.array_equal(a, b)
    for name, a, b in zip(("k", "dkdT", "tau", "dk"), out["f32"], out["f64"]):
        err = relerr(a, b) if name in ("k", "tau") else colerr(a, b)
        assert 0.0 < err < 2e-5, (name, err)
    K2 = tab["K"].copy()
    K2[3, 2, 1, 1, 0] *= 1.0 + 1e-12                     # not a float32 number any more
    with pytest.raises(ValueError):
        ops.Table(K2, "k32")
    ops.Table(K2, "f32").close()


def test_path_mix_matches_numpy_bit_for_bit():
    """ansb200_path_mix: the tangent-height interpolation of nemesisSOfmg / nemesisLfmg (ForwardModel_0.py:1206-1228)."""
    import torch
    from archnemesis_dist_b200 import ops, plan
    rng = np.random.default_rng(5)
    nw, npath, nx = 37, 9, 70
    spec, dx = rng.standard_normal((nw, npath)), rng.standard_normal((nw, npath, nx))
    base = np.sort(rng.uniform(10.0, 200.0, npath))
    tan = np.array([base[0], base[0] + 1.0, 0.5 * (base[3] + base[4]), base[5], base[-1] - 1e-3, base[-1], base[-1] + 30.0])
    mix = plan.tangent_mix(base, tan)
    assert (mix["hi"] < 0).sum() == 2              # at and above the highest path: that path alone
    out = ops.path_mix(ops.to_dev(spec), ops.to_dev(dx), mix).cpu().numpy()
    full = np.concatenate([spec[:, :, None], dx], axis=2)
    for i, (lo, hi, a, b) in enumerate(zip(mix["lo"], mix["hi"], mix["wlo"], mix["whi"])):
        want = full[:, lo] if hi < 0 else full[:, lo] * a + full[:, hi] * b
        assert np.array_equal(out[:, i], want), i
    with pytest.raises(ValueError):
        plan.tangent_mix(base, [base[0] - 5.0])
    assert torch.cuda.is_available()


def test_kdist_kernel_against_golden_and_oracle():
    """ansb200_kdist: k-distributions of spectral bins (tail of calc_ktable_chunk, Spectroscopy_0.py:3619-3660) against
    the reference's own output (tests/golden/kdist.npz, made by tests/test_ktable_dropin.py from the live reference),
    with and without an instrument function, and against the oracle on bins of every size class: one point, a
    non-power-of-two, the capacity of the shared-memory sort, and one beyond it (library sort on the device)."""
    import os
    from archnemesis_dist_b200 import ktable, ops
    from oracle import oracle
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kdist.npz"))
    got = ktable.k_distribution(z["plain_kabs"], z["plain_wavecalc"], z["plain_vbinmin"], z["plain_vbinmax"], z["g_ord"])
    assert relerr(got, z["plain_k"]) < 1e-11        # (g_i = (i+1)/n here, n accumulated steps in the reference)
    ils = lambda ib, wv: np.interp(wv - z["centres"][ib], z["vfil_rel"], z["afil"])      # noqa: E731
    got = ktable.k_distribution(z["ils_kabs"], z["ils_wavecalc"], z["ils_vbinmin"], z["ils_vbinmax"], z["g_ord"], ils)
    assert relerr(got, z["ils_k"]) < 1e-11
    rng = np.random.default_rng(12)
    cap = ops.kdist_capacity(False)
    assert cap == 16384 and ops.kdist_capacity(True) == 8192
    sizes = [1, 2, 33, 1000, 4097, cap, cap + 5]
    wave = np.arange(sum(sizes), dtype=np.float64) * 0.25 + 100.0
    kabs = 10.0 ** rng.uniform(-28.0, -20.0, len(wave))
    kabs[5:9] = kabs[5]                                # ties
    edges = np.concatenate([[0], np.cumsum(sizes)])
    vmin, vmax = wave[edges[:-1]], wave[edges[1:] - 1]
    g = np.array([0.0, 1e-6, 0.013, 0.25, 0.5, 0.77, 0.987, 1.0 - 1e-9, 1.0])
    want = oracle.k_distribution(kabs, wave, vmin, vmax, g)
    got = ktable.k_distribution(kabs, wave, vmin, vmax, g)
    for ib in range(len(sizes)):
        assert relerr(got[ib], want[ib]) < 1e-10, sizes[ib]
    wfun = lambda ib, wv: 0.2 + np.abs(np.sin(wv))                                        # noqa: E731
    ok = [i for i, n in enumerate(sizes) if n <= cap]
    want = oracle.k_distribution(kabs, wave, vmin, vmax, g, wfun)
    got = ktable.k_distribution(kabs, wave, vmin, vmax, g, wfun)       # (the two largest bins: library sort)
    for ib in ok + [len(sizes) - 1]:
        assert relerr(got[ib], want[ib]) < 1e-10, sizes[ib]
    with pytest.raises(ValueError):
        ktable.k_distribution(kabs, wave, [wave[3] + 0.01], [wave[3] + 0.02], g)          # a bin without grid points


@pytest.mark.parametrize("shared", [False, True])
def test_projection_with_chunk_list_equals_full_product(shared):
    """ansb200_jacobian_project_chunks: the projection that skips the 16-row chunks of M without a non-zero (whole
    parameters that no state-vector element touches) is the same sum, bit for bit; an all-zero M gives zeros."""
    import torch
    from archnemesis_dist_b200 import ops
    rng = np.random.default_rng(8)
    nw, npath, npar, nlm, nx = 150, 3, 7, 37, 45            # E = 259: not a multiple of 16
    dspec = rng.standard_normal((nw, npath, npar, nlm))
    P = 1 if shared else npath
    M = np.zeros((P, npar * nlm, nx))
    for k in (1, 4, 6):                                     # parameters 0, 2, 3, 5 have no state-vector element
        M[:, k * nlm:(k + 1) * nlm, rng.integers(0, nx, 9)] = rng.standard_normal((P, nlm, 9))
    d, Md = ops.to_dev(dspec), ops.to_dev(M)
    full = ops.jacobian_project(d, Md, shared=shared)
    chunks = ops.project_chunks(M)
    assert 0 < chunks.numel() < (npar * nlm + 15) // 16
    # rows of dspec the list leaves out may hold anything (here NaN): they are never read
    poisoned = dspec.copy()
    keep = np.zeros(((npar * nlm + 15) // 16) * 16, dtype=bool)
    keep.reshape(-1, 16)[chunks.cpu().numpy()] = True
    poisoned.reshape(nw, npath, -1)[:, :, ~keep[:npar * nlm]] = np.nan
    got = ops.jacobian_project(ops.to_dev(poisoned), Md, shared=shared, chunks=chunks)
    assert torch.equal(got, full)
    want = np.einsum("wpe,pex->wpx", dspec.reshape(nw, npath, -1), np.broadcast_to(M, (npath,) + M.shape[1:]))
    assert colerr(got.cpu().numpy(), want) < 1e-13
    zero = ops.jacobian_project(d, ops.to_dev(np.zeros_like(M)), shared=shared, chunks=ops.project_chunks(np.zeros_like(M)))
    assert float(zero.abs().max()) == 0.0


@pytest.mark.parametrize("shared", [False, True])
def test_sparse_projection_equals_dense_product(shared):
    """ansb200_jacobian_project_sparse (M by columns, a warp per dspec row) against the tiled dense product: columns of
    at most 16 entries are summed in the same order with fused multiply-adds -- bit-identical; longer ones (a scaling
    factor that touches every layer) by the whole warp -- equal to rounding; empty columns give zero."""
    import torch
    from archnemesis_dist_b200 import ops, plan
    rng = np.random.default_rng(9)
    nw, npath, npar, nlm, nx = 70, 3, 7, 37, 45
    dspec = rng.standard_normal((nw, npath, npar, nlm))
    P = 1 if shared else npath
    M = np.zeros((P, npar * nlm, nx))
    for p in range(P):
        M[p, 1 * nlm:2 * nlm, 3] = rng.standard_normal(nlm)                 # long column (37 entries)
        M[p, 4 * nlm:5 * nlm, 40] = rng.standard_normal(nlm)
        for x in range(8, 30):                                              # short columns: 2-3 neighbouring layers
            l0 = int(rng.integers(0, nlm - 3))
            M[p, 6 * nlm + l0:6 * nlm + l0 + 3, x] = rng.standard_normal(3)
        M[p, 5, 44] = 2.5                                                   # one entry; columns 0-2, 4-7, ... stay empty
    M[:, 6 * nlm + 10, 12] = 0.0                                            # (a zero inside a run is kept)
    sp_host = plan.sparse_projection(M)
    assert sp_host is not None and len(sp_host["long_cols"]) == 2 * P
    assert plan.sparse_projection(rng.standard_normal((1, 64, 8))) is None   # dense: left to the tiled kernel
    far = np.zeros((1, 400, 4))
    far[0, [3, 390], 1] = 1.0
    assert plan.sparse_projection(far) is None                              # a column spread over distant rows: likewise
    d = ops.to_dev(dspec)
    dense = ops.jacobian_project(d, ops.to_dev(M), shared=shared)
    got = ops.jacobian_project_sparse(d, ops.SparseProjection(sp_host), shared=shared)
    short = np.ones(nx, dtype=bool)
    short[[3, 40]] = False
    assert torch.equal(got[:, :, torch.as_tensor(short)], dense[:, :, torch.as_tensor(short)])
    assert colerr(got.cpu().numpy(), dense.cpu().numpy()) < 1e-14
    assert float(got[:, :, 0].abs().max()) == 0.0
