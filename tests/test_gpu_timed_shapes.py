"""GPU parity at the shapes bench.py and tools/measure_configs.py time (BASELINE.json configs 2-5).

The overlap kernel is persistent: its grid is capped at the number of resident CTAs, so a warp only takes a
second trip through the cell loop when the case has more (wavenumber, layer) cells than resident warps
(148 SMs x up to 32 warps).  The cases here have NG 20 / NGAS 6 / NLAY 100 like config 2 and 12 800 - 19 200
cells, i.e. every warp walks several cells and stale per-cell state would show.  The other tests take the
multi-path kernels, the projection and the line-by-line kernel to their benchmarked sizes:
NLAYIN = 200 x 64 limb paths, NX = 1000, 10^4 lines over many shared-memory tiles.

Tolerances as in test_gpu_kernels.py (1e-9 is the BASELINE.json bound; observed errors are ~1e-13).
"""
import numpy as np
import pytest

from tests.util import relerr, colerr, cpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def reference_tie_order():
    from oracle import oracle
    oracle.set_sort_mode(oracle.NUMBA_ORDER)
    yield
    oracle.set_sort_mode(oracle.NUMBA_ORDER)


@pytest.fixture(scope="module")
def mods():
    import torch
    from archnemesis_dist_b200 import ops, plan, synthetic, engine
    from oracle import oracle
    return dict(torch=torch, ops=ops, plan=plan, syn=synthetic, orc=oracle, engine=engine, nt=oracle.max_threads())


@pytest.fixture(scope="module")
def config2_slice(mods):
    """Config 2 with NWAVE = 128 (12 800 cells) and the oracle's k / k_overlapg on it."""
    c = mods["syn"].make_fm_case(nwave=128, ng=20, ngas=6, nlay=100, npro=100, nx=60, nvmr=8, seed=7)
    tab, orc, nt = c["tab"], mods["orc"], mods["nt"]
    k, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True, nthreads=nt)
    rt, rd = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT, nthreads=nt)
    rt0 = orc.k_overlap(tab["DELG"], k, c["amount"], nthreads=nt)
    return dict(c=c, k=k, dkdT=dkdT, tau=rt, dk=rd, tau_nograd=rt0)


def _check_grad(tau, dk, rt, rd, tol=1e-13):
    assert relerr(cpu(tau), rt) < tol
    got = cpu(dk)
    for col in range(rd.shape[-1]):
        assert colerr(got[..., col], rd[..., col]) < tol, col
    big = np.abs(rd) > 1e-6 * np.abs(rd).max()
    assert relerr(got[big], rd[big]) < 1e-10


def test_overlap_config2_cells_exceed_resident_warps(mods, config2_slice):
    """k_overlapg / k_overlap on 12 800 cells of the config-2 shape, unfused entry point: parallel rebin against
    the oracle, literal sequential rebin (force_seq) bit for bit."""
    ops = mods["ops"]
    s = config2_slice
    c = s["c"]
    otab = ops.OverlapTables(c["tab"]["DELG"])
    kd, dd, am = ops.to_dev(s["k"]), ops.to_dev(s["dkdT"]), ops.to_dev(c["amount"])
    ncell = s["k"].shape[0] * s["k"].shape[2]
    assert ncell >= 3 * 148 * 24          # several cell-loop iterations per warp at any occupancy
    tau, dk = ops.koverlap(kd, am, otab, dkdT=dd)
    _check_grad(tau, dk, s["tau"], s["dk"])
    ts, ds = ops.koverlap(kd, am, otab, dkdT=dd, force_seq=True)
    assert np.array_equal(cpu(ts), s["tau"]) and np.array_equal(cpu(ds), s["dk"])
    assert relerr(cpu(ops.koverlap(kd, am, otab)), s["tau_nograd"]) < 1e-13
    assert relerr(cpu(ops.koverlap(kd, am, otab, force_seq=True)), s["tau_nograd"]) < 1e-14
    # a second launch over the same buffers gives the same bits (no state left behind by the first)
    tau2, dk2 = ops.koverlap(kd, am, otab, dkdT=dd)
    assert mods["torch"].equal(tau, tau2) and mods["torch"].equal(dk, dk2)


@pytest.mark.parametrize("want_grad", [True, False])
def test_gas_opacity_fused_config2_cells_exceed_resident_warps(mods, config2_slice, want_grad):
    """The fused entry point (the kernel bench.py times) on the same 12 800 cells: equal to k-interp followed by
    the overlap kernel bit for bit, and to the oracle chain within the k-interp tolerance."""
    ops, plan = mods["ops"], mods["plan"]
    s = config2_slice
    c = s["c"]
    tab = c["tab"]
    hp = plan.kinterp_plan(tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad)
    T = ops.Table(tab["K"])
    dp = ops.DevicePlan(hp, want_grad)
    otab = ops.OverlapTables(tab["DELG"])
    am = ops.to_dev(c["amount"])
    fused = ops.gas_opacity(T, dp, am, otab, want_grad)
    if want_grad:
        k, d = ops.kinterp(T, dp, True)
        tau, dk = ops.koverlap(k, am, otab, dkdT=d)
        assert np.array_equal(cpu(fused[0]), cpu(tau)) and np.array_equal(cpu(fused[1]), cpu(dk))
        assert relerr(cpu(fused[0]), s["tau"]) < 1e-11
        got = cpu(fused[1])
        for col in range(s["dk"].shape[-1]):
            assert colerr(got[..., col], s["dk"][..., col]) < 1e-10, col
        seq = ops.gas_opacity(T, dp, am, otab, True, force_seq=True)
        assert relerr(cpu(seq[0]), s["tau"]) < 1e-11
    else:
        tau = ops.koverlap(ops.kinterp(T, dp, False), am, otab)
        assert np.array_equal(cpu(fused), cpu(tau))
        assert relerr(cpu(fused), s["tau_nograd"]) < 1e-11
    T.close()


def test_overlap_config2_mixed_regimes_many_cells(mods):
    """19 200 cells where the folds of one cell take different routes (dead gases, a gas 1e-25 below the rest so
    that whole rows of keys tie, dominant gases giving the data-independent orders): what a warp leaves in its
    shared-memory slots after one route must not leak into the next cell's other route."""
    ops, orc, nt = mods["ops"], mods["orc"], mods["nt"]
    c = mods["syn"].make_fm_case(nwave=192, ng=20, ngas=6, nlay=100, npro=100, nx=8, nvmr=8, seed=23, zero_fraction=0.2)
    tab = c["tab"]
    k, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True, nthreads=nt)
    rng = np.random.default_rng(5)
    # per (wavenumber, gas) regime: 0 normal, 1 negligible (ties), 2 dominant, 3 weak (row-major order)
    regime = rng.integers(0, 4, size=(192, 6))
    f = np.choose(regime, [1.0, 1e-25, 1e9, 1e-8])
    k *= f[:, None, None, :]
    dkdT *= f[:, None, None, :]
    otab = ops.OverlapTables(tab["DELG"])
    kd, dd, am = ops.to_dev(k), ops.to_dev(dkdT), ops.to_dev(c["amount"])
    rt, rd = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT, nthreads=nt)
    tau, dk = ops.koverlap(kd, am, otab, dkdT=dd)
    _check_grad(tau, dk, rt, rd)
    ts, ds = ops.koverlap(kd, am, otab, dkdT=dd, force_seq=True)
    assert np.array_equal(cpu(ts), rt) and np.array_equal(cpu(ds), rd)
    assert relerr(cpu(ops.koverlap(kd, am, otab)), orc.k_overlap(tab["DELG"], k, c["amount"], nthreads=nt)) < 1e-13


def _evaluation(mods, c, **kw):
    e = mods["engine"]
    a = dict(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"], NVMR=c["NVMR"],
             NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"], EMTEMP=c["EMTEMP"],
             LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"], TSURF=c["TSURF"],
             EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"], ISPACE=c["ISPACE"])
    a.update(kw)
    return e.Evaluation(**a)


def _oracle_forward_jacobian(mods, c, mode="thermal"):
    orc, nt = mods["orc"], mods["nt"]
    tab = c["tab"]
    k, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True, nthreads=nt)
    tau, dk = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT, nthreads=nt)
    tl, tp, dtl = orc.assemble_opacity(tau, dk, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"], c["dtaucon"],
                                       c["LAYINC"], c["SCALE"])
    if mode == "thermal":
        S, dS, dT = orc.thermal_paths(c["ISPACE"], tab["WAVE"], tl, dtl, c["NVMR"], c["NLAYIN"], c["EMTEMP"],
                                      c["LAYPRESS"], c["LAYINC"], c["TSURF"], c["EMISSIVITY"], c["xfac"], nthreads=nt)
    else:
        S, dS = orc.transmission(tp, dtl, c["xfac"])
        dT = None
    spec, dspec, dts = orc.g_integrate(S, dS, dT, tab["DELG"])
    npath = c["LAYINC"].shape[1]
    d2 = orc.map2pro(dspec, tab["NWAVE"], c["NVMR"], c["NDUST"], c["NPRO"], npath, c["NLAYIN"], c["LAYINC"], c["DTE"],
                     c["DAM"], c["DCO"], INCPAR=orc.included_params(c["xmap"]))
    return spec, dspec, orc.map2xvec(d2, c["xmap"]), dts


def test_forward_jacobian_config2_shape_end_to_end(mods):
    """HotPath.forward_jacobian (the call bench.py's e2e arm makes) on 96 wavenumbers of config 2 (9 600 cells):
    spectrum and state-vector Jacobian against the whole oracle chain, 1e-9 as BASELINE.json states it."""
    c = mods["syn"].make_fm_case(nwave=96, ng=20, ngas=6, nlay=100, npro=100, nx=60, nvmr=8, seed=7)
    tab = c["tab"]
    hp = mods["engine"].HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    M = mods["plan"].fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    spec, dx, _ = hp.forward_jacobian(_evaluation(mods, c), M)
    s_ref, _, x_ref, _ = _oracle_forward_jacobian(mods, c)
    assert relerr(cpu(spec), s_ref) < 1e-9
    got = cpu(dx)
    for ix in range(x_ref.shape[-1]):
        assert colerr(got[..., ix], x_ref[..., ix]) < 1e-9, ix
    big = np.abs(x_ref) > 1e-6 * np.abs(x_ref).max()
    assert relerr(got[big], x_ref[big]) < 1e-9
    hp.close()


def _limb_case(mods, nwave, nlay, npath, seed, nx=40):
    """Config 4: the config-2 atmosphere seen along `npath` limb paths (down to a tangent layer and up again,
    NLAYIN up to 2*NLAY), padded like Path_0 pads them."""
    c = mods["syn"].make_fm_case(nwave=nwave, ng=20, ngas=6, nlay=nlay, npro=nlay, nx=nx, nvmr=8, seed=seed)
    rng = np.random.default_rng(seed + 100)
    nlm = 2 * nlay
    layinc = np.zeros((nlm, npath), np.int32)
    scale = np.zeros((nlm, npath))
    emtemp = np.zeros((nlm, npath))
    nlayin = np.zeros(npath, np.int32)
    for p in range(npath):
        t = (p * (nlay - 1)) // npath              # tangent layer: path 0 crosses every layer twice (NLAYIN = 2 NLAY)
        seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay))
        n = len(seq)
        nlayin[p] = n
        layinc[:n, p] = seq
        scale[:n, p] = rng.uniform(1.0, 25.0, n)
        emtemp[:n, p] = c["temp"][seq]
    c.update(LAYINC=layinc, SCALE=scale, NLAYIN=nlayin, EMTEMP=emtemp, xfac=np.linspace(0.5, 2.0, nwave))
    # optically thin enough that the far side of the limb still shows
    c["amount"] = c["amount"] * 1e-3
    c["taucon"] = c["taucon"] * 1e-2
    return c


def test_thermal_limb_64_paths_of_200_layers(mods):
    """Config 4, thermal limb emission with gradients: 64 paths, NLAYIN up to 200 (7 path layers per lane of
    ans_thermal_paths_kernel), layer-space Jacobian and its projection against the oracle's literal O(N^2)
    recurrence."""
    ops = mods["ops"]
    c = _limb_case(mods, nwave=4, nlay=100, npath=64, seed=7)
    assert c["NLAYIN"].max() == 200
    tab = c["tab"]
    hp = mods["engine"].HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    ev = _evaluation(mods, c)
    spec, dspec, dts = hp.cirsrad(ev, return_grad=True)
    s_ref, d_ref, x_ref, t_ref = _oracle_forward_jacobian(mods, c)
    assert relerr(cpu(spec), s_ref) < 1e-9
    got = np.transpose(cpu(dspec), (0, 2, 3, 1))              # -> (NWAVE, NPAR, NLAYMAX, NPATH)
    for kpar in range(d_ref.shape[1]):
        assert colerr(got[:, kpar], d_ref[:, kpar]) < 1e-9, kpar
    for p in range(64):
        assert not np.any(got[:, :, c["NLAYIN"][p]:, p])
    assert relerr(cpu(dts), t_ref) < 1e-9
    M = mods["plan"].fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    _, dx, _ = hp.forward_jacobian(ev, M)
    gx = cpu(dx)
    for ix in range(x_ref.shape[-1]):
        assert colerr(gx[..., ix], x_ref[..., ix]) < 1e-9, ix
    hp.close()


def test_transmission_64_paths_of_200_layers(mods):
    """Config 4, solar occultation: the same 64 paths through ans_transmission_paths_kernel and the projection."""
    c = _limb_case(mods, nwave=6, nlay=100, npath=64, seed=9)
    c["amount"] = c["amount"] * 1e-2
    tab = c["tab"]
    hp = mods["engine"].HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    ev = _evaluation(mods, c, mode=mods["engine"].TRANSMISSION, EMTEMP=None, LAYPRESS=None, EMISSIVITY=None)
    M = mods["plan"].fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    spec, dx, _ = hp.forward_jacobian(ev, M)
    s_ref, _, x_ref, _ = _oracle_forward_jacobian(mods, c, mode="transmission")
    assert s_ref.min() < 0.7 * s_ref.max()              # the paths are neither all opaque nor all clear
    assert relerr(cpu(spec), s_ref) < 1e-9
    gx = cpu(dx)
    for ix in range(x_ref.shape[-1]):
        assert colerr(gx[..., ix], x_ref[..., ix]) < 1e-9, ix
    hp.close()


def test_projection_nx_1000(mods):
    """Config 5's widest state vector: map2pro + map2xvec with NX = 1000 on the config-2 layer count."""
    ops, plan, orc = mods["ops"], mods["plan"], mods["orc"]
    c = mods["syn"].make_fm_case(nwave=48, ng=4, ngas=2, nlay=100, npro=100, nx=1000, nvmr=8, seed=15)
    rng = np.random.default_rng(2)
    dspec_ref = rng.normal(size=(48, c["NPAR"], 100, 1)) * 10.0 ** rng.uniform(-6, 0, size=(48, c["NPAR"], 100, 1))
    M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    assert M.shape == (1, c["NPAR"] * 100, 1000)
    out = ops.jacobian_project(ops.to_dev(np.transpose(dspec_ref, (0, 3, 1, 2))), ops.to_dev(M))
    d2 = orc.map2pro(dspec_ref, 48, c["NVMR"], c["NDUST"], c["NPRO"], 1, c["NLAYIN"], c["LAYINC"], c["DTE"], c["DAM"],
                     c["DCO"], INCPAR=orc.included_params(c["xmap"]))
    ref = orc.map2xvec(d2, c["xmap"])
    got = cpu(out)
    assert got.shape == ref.shape == (48, 1, 1000)
    for ix in range(0, 1000, 7):
        assert colerr(got[..., ix], ref[..., ix]) < 1e-12, ix
    assert colerr(got, ref) < 1e-12


def test_lbl_ten_thousand_lines_many_tiles(mods):
    """Config 3's structure at a size the oracle finishes: 10^4 lines (40 shared-memory tiles of 256) spread over
    +-75 cm-1 around a 1024-point grid, two (p,T) points; lines in the core, in the wings and outside the window."""
    orc, syn = mods["orc"], mods["syn"]
    from archnemesis_dist_b200 import lbl
    wn = np.linspace(2000.0, 2002.046, 1024)
    lines = syn.make_line_list(10000, 2000.0, 2002.046, seed=0, pad=80.0)
    pts = [(150.0, 0.05, 1.2), (280.0, 2.0, 0.9)]
    mix = np.array([0.1, 0.9])
    out = lbl.lbl_absorption(wn, lines, pts, t_ref=296.0, p_ref=1.0, abundance=0.99, mass=28.0, mix=mix)
    for i, (t, p, q) in enumerate(pts):
        ref = orc.lbl_absorption(wn, lines, t, p, 296.0, 1.0, q, 0.99, 28.0, mix)
        assert relerr(cpu(out[i]), ref) < 1e-10, i


def _layer_space_against_path_space(mods, c4, mode, tol=1e-13):
    """The layer-space route (visits of a layer added in the radiance kernel, ONE projection matrix for all paths)
    against the path-space route (dspec[NWAVE,NPATH,NPAR,NLAYIN] x per-path M) on the same staged inputs."""
    ops, plan, engine = mods["ops"], mods["plan"], mods["engine"]
    tab = c4["tab"]
    nlay, npath, nwave = len(c4["press"]), int(c4["LAYINC"].shape[1]), tab["NWAVE"]
    nx = c4["xmap"].shape[0]
    if tab["K"].shape[1] == 1:      # one g-ordinate: a line-by-line table (ILBL = 2), K[NWAVE,NP,NT,NGAS]
        hp = engine.HotPath(np.ascontiguousarray(tab["K"][:, 0]), tab["PRESS"], tab["TEMP"], np.array([1.0]), tab["WAVE"])
    else:
        hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    M = plan.fold_projection(c4["xmap"], c4["LAYINC"], c4["NLAYIN"], c4["DTE"], c4["DAM"], c4["DCO"], c4["NVMR"], c4["NDUST"])
    Mlay = plan.fold_projection_layers(c4["xmap"], nlay, c4["DTE"], c4["DAM"], c4["DCO"], c4["NVMR"], c4["NDUST"])
    assert Mlay.shape == (1, c4["NPAR"] * nlay, nx) and M.shape[1] == c4["NPAR"] * int(c4["LAYINC"].shape[0])
    ev = _evaluation(mods, c4, mode=mode)
    s_path = hp.stage(ev, True, M)
    s_lay = hp.stage(ev, True, M, Mlay=Mlay)
    assert s_lay.layer_space and not getattr(s_path, "layer_space", False)
    a_spec, a_dx, a_dt = hp.run(s_path)
    b_spec, b_dx, b_dt = hp.run(s_lay)
    if mode == engine.TRANSMISSION:
        assert mods["torch"].equal(a_spec, b_spec)
    else:
        assert relerr(cpu(b_spec), cpu(a_spec)) < 1e-14 and relerr(cpu(b_dt), cpu(a_dt)) < 1e-14
    got, ref = cpu(b_dx), cpu(a_dx)
    for ix in range(ref.shape[-1]):
        assert colerr(got[..., ix], ref[..., ix]) < tol, ix
    # the layer-space array itself: the visits of a layer added
    tau, dk = hp.gas_opacity(s_path)
    args = (s_path.mode, tau, dk, s_path.gas_slot, s_path.taucia, s_path.taudust, s_path.tauray, s_path.dtaucon,
            s_path.layinc, s_path.scale, s_path.nlayin, s_path.emtemp, s_path.laypress, hp.wave_d, hp.delg_d,
            s_path.emissivity, s_path.xfac, None, None, None, None, s_path.ISPACE, s_path.TSURF, s_path.NVMR, s_path.NPAR, True)
    _, d_path, _ = ops.radiance(*args)
    _, d_lay, _ = ops.radiance(*args, layer_space=True)
    d_path, d_lay = cpu(d_path), cpu(d_lay)
    assert d_lay.shape == (nwave, npath, c4["NPAR"], nlay)
    summed = np.zeros_like(d_lay)
    for p in range(npath):
        n = int(c4["NLAYIN"][p])
        np.add.at(summed[:, p].transpose(2, 0, 1), c4["LAYINC"][:n, p], d_path[:, p, :, :n].transpose(2, 0, 1))
    for k in range(c4["NPAR"]):
        assert colerr(d_lay[:, :, k], summed[:, :, k]) < tol, k
    hp.close()
    return got


@pytest.mark.parametrize("mode_name", ["transmission", "thermal"])
def test_layer_space_gradients_of_limb_paths(mods, mode_name):
    """ANSB200_RAD_LAYER_SPACE on limb / occultation paths (every layer above the tangent is crossed twice), 16 paths x
    up to 200 positions: ans_transmission_paths_kernel and ans_thermal_layers_kernel (8-path tiles, tensor-core g-sums)."""
    import bench
    engine = mods["engine"]
    c = mods["syn"].make_fm_case(nwave=24, ng=20, ngas=6, nlay=100, npro=100, nx=60, nvmr=8, seed=7)
    c4 = bench.limb_case(c, 16)
    _layer_space_against_path_space(mods, c4, engine.TRANSMISSION if mode_name == "transmission" else engine.THERMAL)


@pytest.mark.parametrize("mode_name", ["thermal", "transmission"])
@pytest.mark.parametrize("ng,ngas,nlay,npath", [(18, 7, 37, 13), (16, 3, 64, 5), (20, 6, 100, 9), (4, 2, 21, 70),
                                                (10, 5, 50, 33), (1, 4, 30, 6)])
def test_layer_space_irregular_paths(mods, ng, ngas, nlay, npath, mode_name):
    """ans_thermal_layers_kernel and the layer-space ans_transmission_paths_kernel away from the timed shape: NG not a
    multiple of 4 and below 8 (the transmission kernel then keeps a warp's lanes on the visits, not on the g-ordinates),
    NGAS + 2 > 8 columns (two column tiles), numbers of paths that leave the last tile ragged or need a second group of
    64, NLAY not a multiple of 8, and paths of every kind side by side -- limb paths, nadir paths that end on the ground
    (surface term), a path that crosses some layers three times, one of a single layer -- against the path-space
    kernels."""
    engine = mods["engine"]
    rng = np.random.default_rng(ng * 100 + npath)
    c = mods["syn"].make_fm_case(nwave=12, ng=ng, ngas=ngas, nlay=nlay, npro=nlay, nx=30, nvmr=ngas + 2, seed=11)
    nlm = 2 * nlay + 9
    layinc = np.zeros((nlm, npath), np.int32)
    scale = np.zeros((nlm, npath))
    emtemp = np.zeros((nlm, npath))
    nlayin = np.zeros(npath, np.int32)
    for p in range(npath):
        kind = p % 4
        if kind == 0:      # limb, tangent somewhere in the atmosphere
            t = int(rng.integers(0, nlay - 1))
            seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay))
        elif kind == 1:    # nadir: top to bottom, ends on the ground
            seq = list(range(nlay - 1, -1, -1))
        elif kind == 2:    # crosses the upper layers three times (down, up, down again)
            t = nlay // 2
            seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay)) + list(range(nlay - 1, nlay - 5, -1))
        else:              # a single layer
            seq = [int(rng.integers(0, nlay))]
        seq = np.array(seq[:nlm])
        n = len(seq)
        nlayin[p] = n
        layinc[:n, p] = seq
        scale[:n, p] = 1.0 + rng.uniform(0.0, 3.0, n)
        emtemp[:n, p] = c["temp"][seq]
    c4 = dict(c, LAYINC=layinc, SCALE=scale, NLAYIN=nlayin, EMTEMP=emtemp)
    _layer_space_against_path_space(mods, c4, engine.THERMAL if mode_name == "thermal" else engine.TRANSMISSION)
