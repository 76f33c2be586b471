"""Seeded synthetic k-tables, atmospheres and line lists of the shapes BASELINE.json names.

The recipes are the ones SURVEY.md section 8d fixes for the benchmark configs.  Table values are
rounded through float32 exactly as a ``.kta`` round trip leaves them
(archnemesis/Spectroscopy_0.py:67, :2844-2850), and PRESS/TEMP/DELG/G_ORD are float32 like the
legacy reader leaves them (:2544-2559), so both the reference and this package see the same bits.
"""
import numpy as np


def gauss_legendre_01(ng):
    """g-ordinates / weights on [0,1] as calc_ktable builds them (Spectroscopy_0.py:3446-3448)."""
    x, w = np.polynomial.legendre.leggauss(ng)
    return (0.5 * (x + 1.0)).astype(np.float32), (0.5 * w).astype(np.float32)


def make_ktable(nwave, ng, npress, ntemp, ngas, seed, zero_fraction=0.0):
    """K[NWAVE,NG,NP,NT,NGAS] float64 (gas fastest), PRESS[NP] atm f32, TEMP[NT] K f32,
    G_ORD/DELG f32, WAVE f64."""
    rng = np.random.default_rng(seed)
    press = np.exp(np.linspace(-15.0, 2.0, npress)).astype(np.float32)
    temp = np.linspace(70.0, 300.0, ntemp).astype(np.float32)
    g_ord, del_g = gauss_legendre_01(ng)
    base = 10.0 ** rng.uniform(-27.0, -20.0, size=(nwave, 1, 1, 1, ngas))
    gshape = (10.0 ** np.linspace(-2.0, 2.0, ng)).reshape(1, ng, 1, 1, 1)
    pfac = (press.astype(np.float64) ** 0.3).reshape(1, 1, npress, 1, 1)
    tfac = ((temp.astype(np.float64) / 150.0) ** 1.5).reshape(1, 1, 1, ntemp, 1)
    K = base * gshape * pfac * tfac
    if zero_fraction > 0.0:
        # whole (wave, gas) columns without absorption, as real tables have outside bands
        dead = rng.uniform(size=(nwave, 1, 1, 1, ngas)) < zero_fraction
        K = np.where(dead, 0.0, K)
    # .kta storage is float32(k * 1e20); read_ktable divides that float32 array by the Python float
    # 1e20, which numpy evaluates in float32 (Spectroscopy_0.py:2844-2850), then widens: every table
    # value is exactly float32-representable
    K = ((K * 1.0e20).astype(np.float32) / np.float32(1.0e20)).astype(np.float64)
    wave = np.linspace(100.0, 100.0 + 0.25 * (nwave - 1), nwave)
    return dict(K=np.ascontiguousarray(K), PRESS=press, TEMP=temp, G_ORD=g_ord, DELG=del_g, WAVE=wave,
                NWAVE=nwave, NG=ng, NP=npress, NT=ntemp, NGAS=ngas)


def make_layers(nlay, ngas, seed, p_top=1.0e-6, p_bot=5.0):
    """Layer pressures (atm), temperatures (K) and absorber amounts (cm-2, = AMOUNT*1e-4) of config 2."""
    rng = np.random.default_rng(seed + 1)
    press = np.exp(np.linspace(np.log(p_bot), np.log(p_top), nlay))
    temp = 110.0 + 60.0 * np.abs(np.linspace(-1.0, 1.0, nlay))
    amount = 10.0 ** rng.uniform(18.0, 24.0, size=(ngas, nlay)) * 1.0e-4
    return press, temp, amount


def make_fm_case(nwave=64, ng=20, npress=20, ntemp=15, ngas=6, nlay=100, nvmr=8, ndust=0, npro=100,
                 nx=60, seed=7, tsurf=-1.0, zero_fraction=0.0):
    """Everything one nadir forward+Jacobian evaluation consumes, as plain arrays
    (the attribute subset of SpectroscopyX/LayerX/PathX/AtmosphereX/SurfaceX/Variables listed in
    SURVEY.md 8b).  Continuum terms and the layer->profile matrices are synthetic but of the
    reference's shapes and sparsity (DAM/DTE/DCO are banded, xmap rows are per-parameter)."""
    rng = np.random.default_rng(seed + 2)
    tab = make_ktable(nwave, ng, npress, ntemp, ngas, seed, zero_fraction)
    press, temp, amount = make_layers(nlay, ngas, seed)
    npar = nvmr + 2 + ndust
    gas_slot = np.sort(rng.choice(nvmr, size=ngas, replace=False)).astype(np.int32)
    taucon = 10.0 ** rng.uniform(-6.0, -2.0, size=(nwave, nlay))
    dtaucon = np.zeros((nwave, npar, nlay))
    dtaucon[:, :nvmr, :] = 10.0 ** rng.uniform(-30.0, -26.0, size=(nwave, nvmr, nlay))
    # nadir path through every layer, bottom-up list as AtmCalc builds it (top of list = top layer)
    layinc = np.arange(nlay - 1, -1, -1, dtype=np.int32).reshape(nlay, 1)
    scale = np.full((nlay, 1), 1.0 / np.cos(np.deg2rad(10.0)))
    emtemp = temp[layinc[:, 0]].reshape(nlay, 1).copy()
    # banded layer->profile matrices
    def banded():
        m = np.zeros((nlay, npro))
        centre = np.linspace(0, npro - 1, nlay)
        for l in range(nlay):
            c = int(round(centre[l]))
            for d in (-1, 0, 1):
                if 0 <= c + d < npro:
                    m[l, c + d] = rng.uniform(0.1, 1.0)
        return m
    dte, dam, dco = banded(), banded() * 1.0e22, banded() * 1.0e3
    xmap = np.zeros((nx, npar, npro))
    # temperature profile on the first block of state elements, then one scaling per gas
    nt_elem = min(nx - min(nx - 1, nvmr), npro)
    for i in range(nt_elem):
        lo = i * npro // nt_elem
        hi = (i + 1) * npro // nt_elem
        xmap[i, nvmr, lo:hi] = 1.0
    for i in range(nt_elem, nx):
        gas = (i - nt_elem) % nvmr
        xmap[i, gas, :] = rng.uniform(1e-8, 1e-4, size=npro)
    emissivity = np.zeros(nwave) if tsurf <= 0 else np.full(nwave, 0.9)
    return dict(tab=tab, press=press, temp=temp, amount=amount, gas_slot=gas_slot, NVMR=nvmr, NDUST=ndust,
                NPAR=npar, NPRO=npro, NX=nx, taucon=taucon, dtaucon=dtaucon, LAYINC=layinc, SCALE=scale,
                NLAYIN=np.array([nlay], np.int32), EMTEMP=emtemp, LAYPRESS=press * 101325.0, DTE=dte, DAM=dam,
                DCO=dco, xmap=xmap, TSURF=tsurf, EMISSIVITY=emissivity, ISPACE=0, xfac=np.ones(nwave),
                JSURF=-1)


def make_line_list(nlines, wn_lo, wn_hi, seed=0, n_amb=1, pad=75.0):
    """HITRAN-shaped synthetic line list.  Rows follow LineData_0.py:681."""
    rng = np.random.default_rng(seed)
    nu = np.sort(rng.uniform(wn_lo - pad, wn_hi + pad, nlines))
    sw = 10.0 ** rng.uniform(-28.0, -19.0, nlines)
    e_lower = rng.uniform(0.0, 3000.0, nlines)
    c2 = 2.99792458E10 * 6.62607015E-27 / 1.380649E-16
    stim_ref = 1.0 - np.exp(-c2 * nu / 296.0)
    br = np.zeros((3 * (1 + n_amb), nlines))
    br[0] = rng.uniform(0.05, 0.1, nlines)      # gamma_self
    br[1] = rng.uniform(0.5, 0.8, nlines)       # n_self
    br[2] = 0.0                                 # delta_self
    for a in range(n_amb):
        br[3 * (a + 1)] = rng.uniform(0.03, 0.09, nlines)
        br[3 * (a + 1) + 1] = rng.uniform(0.5, 0.8, nlines)
        br[3 * (a + 1) + 2] = rng.uniform(-0.01, 0.0, nlines)
    return dict(nu=nu, sw=sw, e_lower=e_lower, stim_ref=stim_ref, broadening=br)


def make_continuum(nwave, nlay, nvmr, ndust=0, seed=0, npairs=5, ntemp=12, temp=None):
    """A synthetic continuum plan of the reference's structure (continuum.py): `npairs` collision-induced pairs with
    cross-section tables on `ntemp` temperatures (one para-H2 plane, as in the reference's isotest.tab), one fixed
    spectrum, gas-giant Rayleigh scattering and `ndust` aerosols.  Returns (ContinuumTables, plan) with magnitudes like
    the config-2 synthetic `taucon` (1e-6 .. 1e-2) so that benchmark cases can swap one for the other."""
    from . import continuum as _cont
    rng = np.random.default_rng(seed + 11)
    nterm = npairs + 1
    wgrid = np.linspace(0.0, 1.0, nwave)
    kw = np.zeros((nterm, ntemp, nwave))
    for t in range(npairs):
        shape = np.exp(-0.5 * ((wgrid - rng.uniform(0.1, 0.9)) / rng.uniform(0.05, 0.4)) ** 2) + 0.05
        tdep = (1.0 + 0.3 * rng.uniform(-1, 1)) ** np.linspace(-1.0, 1.0, ntemp)
        kw[t] = 10.0 ** rng.uniform(-46.0, -44.0) * tdep[:, None] * shape[None, :]
    kw[npairs, 0] = 10.0 ** rng.uniform(-47.0, -46.0) * (1.0 + wgrid)
    nplanes = np.array([ntemp] * npairs + [1], dtype=np.int32)
    pairs = [(int(rng.integers(0, nvmr)), int(rng.integers(0, nvmr))) for _ in range(npairs)]
    ext = int(rng.integers(0, nvmr))
    terms = [("pair", a, b) for a, b in pairs] + [("co2", ext, ext)]
    tables = _cont.ContinuumTables(kw, nplanes, dict(terms=terms, NWAVE=nwave, npl=ntemp, has_cia=True))
    # per layer: temperatures on the table grid, mixing ratios, column densities
    tgrid = np.linspace(60.0, 320.0, ntemp)
    temp = np.asarray(temp if temp is not None else 110.0 + 60.0 * np.abs(np.linspace(-1.0, 1.0, nlay)), dtype=np.float64)
    itl = np.clip(np.searchsorted(tgrid, temp) - 1, 0, ntemp - 2)
    fhl = (temp - tgrid[itl]) / (tgrid[itl + 1] - tgrid[itl])
    pl = np.stack([itl, itl + 1, itl, itl + 1], axis=1).astype(np.int32)
    wt = np.stack([1.0 - fhl, fhl, np.full(nlay, 0.5), np.full(nlay, 0.5), 1.0 / (tgrid[itl + 1] - tgrid[itl])], axis=1)
    q = rng.dirichlet(np.ones(nvmr), size=nlay)                         # (NLAY, NVMR)
    totam = 10.0 ** rng.uniform(26.0, 29.0, nlay)                       # m-2
    xfac = (totam * 1.0e-4) ** 2 / (10.0 ** rng.uniform(5.5, 6.5, nlay))
    q1 = np.zeros((nterm, nlay))
    q2 = np.zeros((nterm, nlay))
    ca = np.zeros((nterm, nlay))
    cb = np.zeros((nterm, nlay))
    slots = np.full((nterm, 3), -1, dtype=np.int32)
    for t, (kind, a, b) in enumerate(terms):
        q1[t], q2[t] = q[:, a], q[:, b]
        if kind == "pair":
            slots[t] = (a, b, nvmr - 2)
            ca[t], cb[t] = q[:, b], q[:, a]
        else:
            slots[t] = (a, -1, -1)
            ca[t] = 2.0 * q[:, a]
    ur = (10.0 ** rng.uniform(-33.0, -32.0) * (1.0 + 3.0 * wgrid) ** 4)[None, :]
    ud = 10.0 ** rng.uniform(-13.0, -12.0, size=(ndust, 1)) * (1.0 + wgrid)[None, :]
    vd = 10.0 ** rng.uniform(6.0, 9.0, size=(ndust, nlay))
    plan = dict(NLAY=nlay, NVMR=nvmr, NDUST=ndust, NPAR=nvmr + 2 + ndust, NTERM=nterm, pl=pl, wt=wt, q1=q1, q2=q2,
                slots=slots, ca=ca, cb=cb, xfac=xfac, totam=totam, ur=ur, ud=ud, vr=totam[None, :].copy(),
                vrd=np.ones((1, nlay)), vd=vd, has_cia=True)
    return tables, plan
