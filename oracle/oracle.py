"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the correlated-k forward model + Jacobian hot path.

A restatement (numpy for the host-sized pieces, plain C in ``ansb200_oracle.c`` for the loops) of
the archNEMESIS v1.1.0 functions on the hot path; every function cites the reference lines it
follows.  Parity pin: compared with the *live* reference in the build container
(``tests/test_plan.py::test_oracle_matches_live_reference_functions``) and with the golden vectors under ``tests/golden/`` made by
``oracle/make_golden.py`` from the unmodified reference.  The reference's own tests pin none of
these functions in isolation.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py`` may import this module.  The product package ``archnemesis_dist_b200`` never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libansb200_oracle.so")
_SRC = os.path.join(_HERE, "ansb200_oracle.c")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile the C restatement with gcc (no -march, no FP contraction: numba does not fuse)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
        subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def max_threads():
    return int(lib().orc_max_threads())


NUMBA_ORDER, STABLE_ORDER = 0, 1


def set_sort_mode(mode):
    """Order of EQUAL sort keys in k_overlap[g]: NUMBA_ORDER (default) reproduces the permutation of
    numba's unstable quicksort, i.e. the reference bit for bit; STABLE_ORDER breaks ties by original
    index, which is what the CUDA kernels do.  The two differ only in how gradient rows of tied
    elements are split across bin edges (tau is unaffected)."""
    lib().orc_set_sort_mode(int(mode))


# ----------------------------------------------------------------------------------------------
# k-table interpolation
# ----------------------------------------------------------------------------------------------
def _bracket(grid, x):
    """Nearest-point bracket search with clamping, Spectroscopy_0.py:2336-2371 (calc_k) /
    :2181-2218 (calc_kg).  Returns (lo, hi, clamped_to) with clamped_to None or the grid value."""
    n = len(grid)
    i = int(np.abs(grid - x).argmin())
    clamp = None
    if grid[i] >= x:
        hi = i
        if i == 0:
            clamp = grid[0]
            lo, hi = 0, 1
        else:
            lo = i - 1
    else:
        lo = i
        if i == n - 1:
            clamp = grid[n - 1]
            hi, lo = n - 1, n - 2
        else:
            hi = i + 1
    return lo, hi, clamp


def kinterp_plan(PRESS, TEMP, press, temp, grad):
    """Per-layer interpolation plan with the reference's dtype behaviour (float32 PRESS/TEMP on
    the .kta path; numpy NEP-50 scalar promotion).  calc_k: Spectroscopy_0.py:2331-2389,
    calc_kg: :2176-2236.  The two differ in where the clamped log-pressure comes from."""
    n = len(press)
    ip_lo = np.zeros(n, np.int32)
    it_lo = np.zeros(n, np.int32)
    w4 = np.zeros((n, 4))
    omv = np.zeros(n)
    vv = np.zeros(n)
    dudt = np.zeros(n)
    for l in range(n):
        press1 = press[l]
        temp1 = temp[l]
        ipl, iph, pcl = _bracket(PRESS, press1)
        itl, ith, tcl = _bracket(TEMP, temp1)
        if grad:
            lpress = np.log(press1)            # :2181
            if pcl is not None:
                lpress = np.log(pcl)           # :2187 / :2195 (log of the table's own dtype)
        else:
            if pcl is not None:
                press1 = pcl                   # :2340 / :2348
            lpress = np.log(press1)            # :2373
        if tcl is not None:
            temp1 = tcl
        plo = np.log(PRESS[ipl])
        phi = np.log(PRESS[iph])
        tlo = TEMP[itl]
        thi = TEMP[ith]
        v = (lpress - plo) / (phi - plo)
        u = (temp1 - tlo) / (thi - tlo)
        ip_lo[l] = ipl
        it_lo[l] = itl
        w4[l] = [(1.0 - v) * (1.0 - u), v * (1.0 - u), v * u, (1.0 - v) * u]
        omv[l] = 1.0 - v
        vv[l] = v
        dudt[l] = 1. / (thi - tlo)
    return ip_lo, it_lo, w4, omv, vv, dudt


def calc_k(K, PRESS, TEMP, press, temp, want_grad=False, nthreads=1):
    """calc_k (Spectroscopy_0.py:2298-2437, WAVECALC=None) / calc_kg (:2147-2295).
    K[NWAVE,NG,NP,NT,NGAS] -> k[NWAVE,NG,NLAY,NGAS] (, dkdT)."""
    K = _d(K)
    NWAVE, NG, NP, NT, NGAS = K.shape
    ip_lo, it_lo, w4, omv, vv, dudt = kinterp_plan(PRESS, TEMP, press, temp, want_grad)
    NLAY = len(press)
    k = np.zeros((NWAVE, NG, NLAY, NGAS))
    dkdT = np.zeros((NWAVE, NG, NLAY, NGAS)) if want_grad else None
    w4 = _d(w4)
    lib().orc_kinterp(_p(K), NWAVE, NG, NP, NT, NGAS, NLAY, ip_lo.ctypes.data_as(_ip),
                      it_lo.ctypes.data_as(_ip), _p(w4), _p(omv), _p(vv), _p(dudt), int(want_grad),
                      _p(k), _p(dkdT), int(nthreads))
    return (k, dkdT) if want_grad else k


# ----------------------------------------------------------------------------------------------
# line-by-line tables: Spectroscopy_0.calc_klbl (:1768-1919) / calc_klblg (:1601-1765) and the LBL-table
# branch of calculate_gaseous_line_opacity (ForwardModel_0.py:3795-3815).  numpy restatement.
# ----------------------------------------------------------------------------------------------
def _klbl_bracket(T, t, clamp_low):
    it = int(np.searchsorted(T, t)) - 1
    if clamp_low and it < 0:          # calc_klbl clamps at the first node (:1838-1839); calc_klblg does not
        it = 0
    if it >= len(T) - 1:
        it = len(T) - 2
    return it


def calc_klbl(K, PRESS, TEMP, press, temp, want_grad=False):
    """K[NWAVE,NP,NT,NGAS] -> k[NWAVE,NLAY,NGAS] (, dkdT).  Python's negative index is kept where calc_klblg
    lets ``it`` reach -1 (a layer on the first temperature node)."""
    K = np.asarray(K, dtype=np.float64)
    lnP = np.log(np.asarray(PRESS))
    TEMP = np.asarray(TEMP)
    nw, _, _, ngas = K.shape
    nlay = len(press)
    k = np.zeros((nw, nlay, ngas))
    dkdT = np.zeros((nw, nlay, ngas)) if want_grad else None
    for l in range(nlay):
        pl = min(max(np.log(press[l]), np.min(lnP)), np.max(lnP))
        tl = min(max(temp[l], np.min(TEMP)), np.max(TEMP))
        ip = min(max(int(np.searchsorted(lnP, pl)) - 1, 0), len(lnP) - 2)
        v = (pl - lnP[ip]) / (lnP[ip + 1] - lnP[ip])
        Ta, Tb = (TEMP[ip], TEMP[ip + 1]) if TEMP.ndim == 2 else (TEMP, TEMP)
        ia, ib = _klbl_bracket(Ta, tl, not want_grad), _klbl_bracket(Tb, tl, not want_grad)
        ua, da = (tl - Ta[ia]) / (Ta[ia + 1] - Ta[ia]), 1. / (Ta[ia + 1] - Ta[ia])
        ub, db = (tl - Tb[ib]) / (Tb[ib + 1] - Tb[ib]), 1. / (Tb[ib + 1] - Tb[ib])
        c_lo1, c_lo2 = K[:, ip, ia, :], K[:, ip, ia + 1, :]
        c_hi1, c_hi2 = K[:, ip + 1, ib, :], K[:, ip + 1, ib + 1, :]
        pos = (c_lo1 > 0.0) & (c_lo2 > 0.0) & (c_hi1 > 0.0) & (c_hi2 > 0.0)
        neg = (c_lo1 <= 0.0) & (c_lo2 <= 0.0) & (c_hi1 <= 0.0) & (c_hi2 <= 0.0)
        for mask, f in ((pos, np.log), (neg, lambda a: a)):
            a1, a2, b1, b2 = f(c_lo1[mask]), f(c_lo2[mask]), f(c_hi1[mask]), f(c_hi2[mask])
            x = (1.0 - v) * (1.0 - ua) * a1 + v * (1.0 - ub) * b1 + v * ub * b2 + (1.0 - v) * ua * a2
            kv = np.exp(x) if f is np.log else x
            k[:, l, :][mask] = kv
            if want_grad:
                dx = -a1 * (1.0 - v) * da - b1 * v * db + b2 * v * db + a2 * (1.0 - v) * da
                dkdT[:, l, :][mask] = kv * dx if f is np.log else dx
    return (k, dkdT) if want_grad else k


def lbl_table_opacity(K, PRESS, TEMP, press, temp, amount, want_grad=False):
    """TAUGAS[NWAVE,1,NLAY] (, dk[NWAVE,1,NLAY,NGAS+1] = [k_i ..., dTAUGAS/dT]); amount[NGAS,NLAY] already
    carries the 1e-4 (VLOSDENS).  ForwardModel_0.py:3803-3815: per-gas products, np.sum over the gas axis,
    the temperature derivative as a running sum in gas order."""
    out = calc_klbl(K, PRESS, TEMP, press, temp, want_grad)
    k, dkdT = out if want_grad else (out, None)
    nw, nlay, ngas = k.shape
    per_gas = np.zeros((nw, 1, nlay, ngas))
    for i in range(ngas):
        per_gas[:, 0, :, i] = k[:, :, i] * amount[i]
    tau = np.sum(per_gas, 3)
    if not want_grad:
        return tau
    dk = np.zeros((nw, 1, nlay, ngas + 1))
    for i in range(ngas):
        dk[:, 0, :, i] = k[:, :, i]
        dk[:, 0, :, ngas] = dk[:, 0, :, ngas] + dkdT[:, :, i] * amount[i]
    return tau, dk


# ----------------------------------------------------------------------------------------------
# random overlap
# ----------------------------------------------------------------------------------------------
def overlap_tables(del_g):
    """weight[NG*NG] and g_ord[NG+1] in del_g's own dtype then widened: ForwardModel_0.py:6087
    (float32 product on the .kta path), :6141-6143 (cumsum in del_g's dtype, last forced to 1)."""
    del_g = np.asarray(del_g)
    ng = len(del_g)
    weight = np.zeros(ng * ng)
    il = 0
    for i in range(ng):
        for j in range(ng):
            weight[il] = del_g[i] * del_g[j]
            il += 1
    g_ord = np.zeros(ng + 1)
    g_ord[1:] = np.cumsum(del_g)
    g_ord[ng] = 1
    return weight, g_ord


def k_overlap(del_g, k, amount, dkdT=None, nthreads=1):
    """k_overlap (ForwardModel_0.py:6029-6115) or, with dkdT, k_overlapg (:5842-5957)."""
    k = _d(k)
    amount = _d(amount)
    NWAVE, NG, NLAY, NGAS = k.shape
    weight, g_ord = overlap_tables(del_g)
    want_grad = dkdT is not None
    if want_grad:
        dkdT = _d(dkdT)
    tau = np.zeros((NWAVE, NG, NLAY))
    dk = np.zeros((NWAVE, NG, NLAY, NGAS + 1)) if want_grad else None
    lib().orc_koverlap(_p(k), _p(dkdT), _p(amount), _p(weight), _p(g_ord), NWAVE, NG, NLAY, NGAS,
                       int(want_grad), _p(tau), _p(dk), int(nthreads))
    return (tau, dk) if want_grad else tau


# ----------------------------------------------------------------------------------------------
# opacity assembly (host-sized numpy): ForwardModel_0.py:3857-3877, :3989-4012
# ----------------------------------------------------------------------------------------------
def assemble_opacity(tau_gas, dk, gas_slot, NVMR, NPAR, taucon, dtaucon, LAYINC, SCALE):
    """tau_gas[NWAVE,NG,NLAY]; dk[NWAVE,NG,NLAY,NGAS+1] or None; gas_slot[NGAS] = index of each
    active gas in the atmosphere's VMR list (locate_gas); taucon[NWAVE,NLAY] = TAUCIA+TAUDUST+TAURAY;
    dtaucon[NWAVE,NPAR,NLAY].  Returns TAUTOT_LAYINC[NWAVE,NG,NLAYIN,NPATH], TAUTOT_PATH,
    dTAUTOT_LAYINC[NWAVE,NG,NPAR,NLAYIN,NPATH] (or None)."""
    TAUTOT = tau_gas + taucon[:, None, :]
    TAUTOT_LAYINC = TAUTOT[:, :, LAYINC] * SCALE
    TAUTOT_PATH = np.sum(TAUTOT_LAYINC, 2)
    dTAUTOT_LAYINC = None
    if dk is not None:
        NWAVE, NG, NLAY, ngp1 = dk.shape
        dTAUGAS = np.zeros((NWAVE, NG, NPAR, NLAY))
        for i, slot in enumerate(gas_slot):
            dTAUGAS[:, :, slot, :] = dk[:, :, :, i] * 1.0e-4
        dTAUGAS[:, :, NVMR, :] = dk[:, :, :, ngp1 - 1]
        dTAUTOT = dTAUGAS + dtaucon[:, None, ...]
        dTAUTOT_LAYINC = dTAUTOT[:, :, :, LAYINC] * SCALE
    return TAUTOT_LAYINC, TAUTOT_PATH, dTAUTOT_LAYINC


# ----------------------------------------------------------------------------------------------
# thermal emission, transmission, g-integration
# ----------------------------------------------------------------------------------------------
def thermal(ISPACE, WAVE, TAU, EMITOT, TEMP, PRESS, TSURF, EMISSIVITY, SOLFLUX, REFLECTANCE,
            SOL_ANG, EMISS_ANG, nthreads=1):
    """calc_thermal_emission_spectrum, ForwardModel_0.py:6287-6377.  TAU[NWAVE,NG,NLAYIN]."""
    TAU = _d(TAU)
    NWAVE, NG, NLAYIN = TAU.shape
    spec = np.zeros((NWAVE, NG))
    em = _d(EMITOT) if EMITOT is not None else None
    a = [_d(x) for x in (WAVE, TEMP, PRESS, EMISSIVITY, SOLFLUX, REFLECTANCE)]
    lib().orc_thermal(int(ISPACE), _p(a[0]), _p(TAU), _p(em), _p(a[1]), _p(a[2]),
                      ctypes.c_double(float(TSURF)), _p(a[3]), _p(a[4]), _p(a[5]),
                      ctypes.c_double(float(SOL_ANG)), ctypes.c_double(float(EMISS_ANG)),
                      NWAVE, NG, NLAYIN, _p(spec), int(nthreads))
    return spec


def thermalg(ISPACE, WAVE, TAU, dTAU, NVMR, TEMP, PRESS, TSURF, EMISSIVITY, nthreads=1):
    """calc_thermal_emission_spectrumg, ForwardModel_0.py:6380-6504 (literal O(N^2) recurrence).
    TAU[NWAVE,NG,NLAYIN], dTAU[NWAVE,NG,NPAR,NLAYIN]."""
    TAU = _d(TAU)
    dTAU = _d(dTAU)
    NWAVE, NG, NLAYIN = TAU.shape
    NPAR = dTAU.shape[2]
    spec = np.zeros((NWAVE, NG))
    dspec = np.zeros((NWAVE, NG, NPAR, NLAYIN))
    dts = np.zeros((NWAVE, NG))
    a = [_d(x) for x in (WAVE, TEMP, PRESS, EMISSIVITY)]
    lib().orc_thermalg(int(ISPACE), _p(a[0]), _p(TAU), _p(dTAU), int(NVMR), _p(a[1]), _p(a[2]),
                       ctypes.c_double(float(TSURF)), _p(a[3]), NWAVE, NG, NPAR, NLAYIN,
                       _p(spec), _p(dspec), _p(dts), int(nthreads))
    return spec, dspec, dts


def thermal_paths(ISPACE, WAVE, TAUTOT_LAYINC, dTAUTOT_LAYINC, NVMR, NLAYIN, EMTEMP, LAYPRESS, LAYINC,
                  TSURF, EMISSIVITY, xfac, SOLFLUX=None, REFLECTANCE=None, SOL_ANG=None,
                  EMISS_ANG=None, nthreads=1):
    """Path loop + unit scaling of calculate_thermal_emission_spectrum, ForwardModel_0.py:4216-4247."""
    NWAVE, NG, NLMAX, NPATH = TAUTOT_LAYINC.shape
    grad = dTAUTOT_LAYINC is not None
    SPEC = np.zeros((NWAVE, NG, NPATH))
    dSPEC = np.zeros((NWAVE, NG, dTAUTOT_LAYINC.shape[2], NLMAX, NPATH)) if grad else None
    dTS = np.zeros((NWAVE, NG, NPATH)) if grad else None
    for ip in range(NPATH):
        n = int(NLAYIN[ip])
        emtemp = EMTEMP[0:n, ip]
        empress = LAYPRESS[LAYINC[0:n, ip]]
        if grad:
            s, ds, dt = thermalg(ISPACE, WAVE, TAUTOT_LAYINC[:, :, 0:n, ip], dTAUTOT_LAYINC[:, :, :, 0:n, ip],
                                 NVMR, emtemp, empress, TSURF, EMISSIVITY, nthreads)
            SPEC[:, :, ip] = s
            dSPEC[:, :, :, 0:n, ip] = ds
            dTS[:, :, ip] = dt
        else:
            SPEC[:, :, ip] = thermal(ISPACE, WAVE, TAUTOT_LAYINC[:, :, 0:n, ip], None, emtemp, empress, TSURF,
                                     EMISSIVITY, SOLFLUX, REFLECTANCE, SOL_ANG[ip], EMISS_ANG[ip], nthreads)
        SPEC[:, :, ip] = (SPEC[:, :, ip].T * xfac).T
        if grad:
            dTS[:, :, ip] = (dTS[:, :, ip].T * xfac).T
            dSPEC[:, :, :, :, ip] = np.transpose(np.transpose(dSPEC[:, :, :, :, ip], axes=[1, 2, 3, 0]) * xfac,
                                                 axes=[3, 0, 1, 2])
    return SPEC, dSPEC, dTS


def transmission(TAUTOT_PATH, dTAUTOT_LAYINC, xfac=None):
    """calculate_transmission_spectrum, ForwardModel_0.py:4104-4129."""
    SPEC = np.exp(-TAUTOT_PATH)
    if xfac is not None:
        SPEC = SPEC * xfac[:, None, None]
    dSPEC = None
    if dTAUTOT_LAYINC is not None:
        dSPEC = np.transpose(-SPEC * np.transpose(dTAUTOT_LAYINC, axes=[2, 3, 0, 1, 4]), axes=[2, 3, 0, 1, 4])
    return SPEC, dSPEC


def g_integrate(SPEC, dSPEC, dTSURF, DELG):
    """CIRSrad g-integration, ForwardModel_0.py:4504-4508."""
    out = np.tensordot(SPEC, DELG, axes=([1], [0]))
    if dSPEC is None:
        return out
    d = np.nan_to_num(np.tensordot(dSPEC, DELG, axes=([1], [0])))
    dt = np.tensordot(dTSURF, DELG, axes=([1], [0])) if dTSURF is not None else None
    return out, d, dt


# ----------------------------------------------------------------------------------------------
# layer -> profile -> state vector: ForwardModel_0.py:5319-5424 and the incpar selection :699-707
# ----------------------------------------------------------------------------------------------
def map2pro(dSPECIN, NWAVE, NVMR, NDUST, NPRO, NPATH, NLAYIN, LAYINC, DTE, DAM, DCO, INCPAR=(-1,)):
    DAMx = DAM[LAYINC, :]
    DCOx = DCO[LAYINC, :]
    DTEx = DTE[LAYINC, :]
    out = np.zeros((NWAVE, NVMR + 2 + NDUST, NPRO, NPATH))
    if INCPAR[0] != -1:
        pars = list(INCPAR)
    else:
        pars = list(range(NVMR + 2 + NDUST))
    for ipath in range(NPATH):
        prev = None
        for p in pars:
            if p <= NVMR - 1:
                prev = np.tensordot(dSPECIN[:, p, :, ipath], DAMx[:, ipath, :], axes=(1, 0))
            elif p <= NVMR:
                prev = np.tensordot(dSPECIN[:, p, :, ipath], DTEx[:, ipath, :], axes=(1, 0))
            elif (p > NVMR) and (p <= NVMR + NDUST):
                prev = np.tensordot(dSPECIN[:, p, :, ipath], DCOx[:, ipath, :], axes=(1, 0))
            # p == NVMR+NDUST+1 (para-H2): the reference re-uses the previous product (:5378-5381)
            out[:, p, :, ipath] = prev[:, :]
    return out


def map2xvec(dSPECIN, xmap):
    return np.tensordot(dSPECIN, xmap, axes=([1, 2], [1, 2]))


def included_params(xmap):
    """ForwardModel_0.py:699-702."""
    return [i for i in range(xmap.shape[1]) if np.mean(xmap[:, i, :]) != 0.0]


# ----------------------------------------------------------------------------------------------
# Line-by-line absorption with the reference's Voigt (= SciPy voigt_profile)
# ----------------------------------------------------------------------------------------------
_SHAPE_T = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double)


def _voigt_scipy_cfunc():
    """C pointer to scipy.special.cython_special.voigt_profile(double,double,double) wrapped with
    the alpha_d -> sigma conversion of lineshape/voigt_impl/voigt_scipy.py:52."""
    from scipy.special import voigt_profile
    s2l2 = np.sqrt(2.0 * np.log(2.0))

    def f(d, ad, gl):
        return float(voigt_profile(d, ad / s2l2, gl))
    return _SHAPE_T(f)


def lbl_absorption(wn_grid, lines, t_calc, p_calc, t_ref, p_ref, q_ratio, abundance, mass, mix,
                   s_floor=0.0, wn_calc_window=25.0, wn_approx_window=75.0, shape="voigt"):
    """add_line_set_monochromatic_absorption, LineData_0.py:279-358.
    lines: dict with nu, sw, e_lower, stim_ref [N] and broadening [3*M, N]."""
    wn = _d(wn_grid)
    nu, sw, el, st = (_d(lines[k]) for k in ("nu", "sw", "e_lower", "stim_ref"))
    br = _d(lines["broadening"])
    mix = _d(mix)
    out = np.zeros(len(wn))
    if shape != "voigt":
        raise NotImplementedError(shape)
    cb = _voigt_scipy_cfunc()
    lib().orc_lbl_absorption(_p(wn), len(wn), cb, ctypes.c_double(t_calc), ctypes.c_double(t_ref),
                             ctypes.c_double(p_calc), ctypes.c_double(p_ref), ctypes.c_double(q_ratio),
                             ctypes.c_double(abundance), ctypes.c_double(mass), _p(mix), len(mix), _p(br),
                             _p(nu), _p(sw), _p(el), _p(st), len(nu), ctypes.c_double(s_floor),
                             ctypes.c_double(wn_calc_window), ctypes.c_double(wn_approx_window), _p(out))
    return out


# ----------------------------------------------------------------------------------------------
# instrument line shape: Measurement_0.conv / convg for k-tables (Measurement_0.py:2288-2692)
# ----------------------------------------------------------------------------------------------
def apply_conv(op, y):
    """Apply a conv operator (archnemesis_dist_b200.plan.conv_operator layout, host arrays) to y[NWAVE]
    (spectrum) or y[NWAVE,NCOL] (gradients) in the reference's arithmetic.  Mode 0: np.interp's slope form for
    a 1-D y, SciPy's two-weight form for a 2-D y (np.interp for both when the operator says np_interp_all:
    lblconv / lblconvg); mode 1: sequential sum(f1*y)/sum(f1)."""
    y = np.asarray(y, dtype=np.float64)
    one_d = y.ndim == 1
    y2 = y.reshape(len(y), -1)
    out = np.zeros((op["NCONV"], y2.shape[1]))
    rs, wi, wv = op["row_start"], op["widx"], op["wval"]
    for c in range(op["NCONV"]):
        if op["mode"] == 0 and (one_d or op.get("np_interp_all", False)) and not op.get("weighted_sum_only", False):
            j = op["np_lo"][c]
            x_lo, x_hi, x_new = op["xinfo"][c]
            if op["np_exact"][c]:
                out[c] = y2[j]
            else:
                slope = (y2[j + 1] - y2[j]) / (x_hi - x_lo)
                out[c] = slope * (x_new - x_lo) + y2[j]
        else:
            acc = np.zeros(y2.shape[1])
            for e in range(rs[c], rs[c + 1]):
                acc = acc + wv[e] * y2[wi[e]]
            out[c] = acc / op["norm"][c] if op["mode"] == 1 else acc
    return out[:, 0] if one_d else out


# ----------------------------------------------------------------------------------------------
# optimal-estimation linear algebra: OptimalEstimation_0.py:545-720 (numpy restatement)
# ----------------------------------------------------------------------------------------------
def oe_gain_matrix(KK, SA, SE):
    sa_kt = SA @ KK.T
    M = KK @ sa_kt + SE
    DD = np.linalg.solve(M.T, sa_kt.T).T
    return DD, DD @ KK


def oe_phiret(Y, YN, XN, XA, SE, SA):
    b, d = YN - Y, XN - XA
    if SE.shape == (1, 1):
        meas = float(np.dot(b, b) / float(SE[0, 0]))
    elif np.all(SE == np.diag(np.diagonal(SE))):
        meas = float(np.dot(b / np.diagonal(SE), b))
    else:
        meas = float(b.T @ np.linalg.solve(SE, b))
    apri = float(d.T @ np.linalg.solve(SA, d))
    return meas / len(b), meas + apri


def oe_next_xn(XA, XN, Y, YN, DD, AA):
    return XA + (DD @ (Y - YN)) - (AA @ (XA - XN))


def oe_serr(DD, AA, SA, SE, simple=False):
    a = DD * SE[0, 0] if simple else DD @ SE
    SM = a @ DD.T
    b = AA.copy()
    b[np.diag_indices_from(b)] -= 1.0
    SN = (b @ SA) @ b.T
    return SM, SN, SN + SM


# ----------------------------------------------------------------------------------------------
# continuum plan (archnemesis_dist_b200.continuum): numpy evaluation in the reference's order of operations
# (calc_tau_cia :4688-4776, calculate_layer_opacity :3938-3981, calc_tau_dust :4858-4863, calc_tau_rayleigh*)
# ----------------------------------------------------------------------------------------------
def continuum_eval(tables, plan, want_grad=True):
    """TAUCIA, TAUDUST, TAURAY [NWAVE,NLAY] (None where the reference has none) and dTAUCON[NWAVE,NPAR,NLAY] from a
    continuum plan."""
    kw, npl = tables.kw, tables.nplanes
    NW = tables.meta["NWAVE"]
    NLAY, NVMR, NDUST, NPAR = plan["NLAY"], plan["NVMR"], plan["NDUST"], plan["NPAR"]
    NS = NVMR + 2
    taucia = np.zeros((NW, NLAY)) if plan["has_cia"] else None
    dcia = np.zeros((NW, NLAY, NS))
    for l in range(NLAY):
        fhh_t, fhl_t, fhh_f, fhl_f, dfhldT = plan["wt"][l]
        sum1 = np.zeros(NW)
        for t in range(plan["NTERM"]):
            if npl[t] > 1:          # a cross-section table: four (para, T) planes of this layer
                a, b, c, d = (kw[t, p] for p in plan["pl"][l])
                ktlo = a * fhh_t + b * fhl_t
                kthi = c * fhh_t + d * fhl_t
                k = ktlo * fhh_f + kthi * fhl_f
                dk = (kthi - ktlo) * dfhldT
            else:                   # a fixed spectrum (co2cia, n2n2cia, n2h2cia)
                k, dk = kw[t, 0], None
            q1, q2 = plan["q1"][t, l], plan["q2"][t, l]
            sum1 = sum1 + k * q1 * q2
            sa, sb, st = plan["slots"][t]
            if sa >= 0:
                dcia[:, l, sa] = dcia[:, l, sa] + plan["ca"][t, l] * k
            if sb >= 0:
                dcia[:, l, sb] = dcia[:, l, sb] + plan["cb"][t, l] * k
            if st >= 0 and dk is not None:
                dcia[:, l, st] = dcia[:, l, st] + dk * q1 * q2
        if taucia is not None:
            taucia[:, l] = sum1 * plan["xfac"][l]
        dcia[:, l, :] = dcia[:, l, :] * plan["xfac"][l]
    NR = plan["ur"].shape[0]
    tauray = np.zeros((NW, NLAY))
    dray = np.zeros((NW, NLAY))
    for r in range(NR):
        tauray = tauray + plan["ur"][r][:, None] * plan["vr"][r][None, :]
        dray = dray + plan["ur"][r][:, None] * plan["vrd"][r][None, :]
    taud1 = np.zeros((NW, NLAY, NDUST))
    for i in range(NDUST):
        taud1[:, :, i] = plan["ud"][i][:, None] * plan["vd"][i][None, :]
    taud1 = np.clip(np.nan_to_num(taud1), 0, 1e20)
    taudust = np.sum(taud1, 2)
    if not want_grad:
        return taucia, taudust, tauray, None
    dtaucon = np.zeros((NW, NPAR, NLAY))
    if plan["has_cia"]:
        dtaucon[:, 0:NVMR, :] = dtaucon[:, 0:NVMR, :] + np.transpose(
            np.transpose(dcia[:, :, 0:NVMR], axes=(2, 0, 1)) / plan["totam"], axes=(1, 0, 2))
        dtaucon[:, NVMR, :] = dtaucon[:, NVMR, :] + dcia[:, :, NVMR]
    if NR > 0:
        for i in range(NVMR):
            dtaucon[:, i, :] = dtaucon[:, i, :] + dray
    for i in range(NDUST):
        dtaucon[:, NVMR + 1 + i, :] = dtaucon[:, NVMR + 1 + i, :] + plan["ud"][i][:, None]
    return taucia, taudust, tauray, dtaucon


# ----------------------------------------------------------------------------------------------
# k-distribution of one spectral bin: the tail of calc_ktable_chunk (Spectroscopy_0.py:3619-3660), numpy as written there
# ----------------------------------------------------------------------------------------------
def k_distribution(kabs, wavecalc, vbinmin, vbinmax, g_ord, ils=None):
    """kabs, wavecalc [ncalc]: the line-by-line spectrum of one (p, T) point; vbinmin / vbinmax [NBIN]; ils(ibin,
    wavesel) -> the instrument function at those grid points (None: ones).  Returns k[NBIN, NG]."""
    kabs, wavecalc = np.asarray(kabs, dtype=np.float64), np.asarray(wavecalc, dtype=np.float64)
    out = np.zeros((len(vbinmin), len(g_ord)))
    for ib in range(len(vbinmin)):
        mask = (wavecalc >= vbinmin[ib]) & (wavecalc <= vbinmax[ib])
        idx = np.argsort(kabs[mask])
        wavesel = wavecalc[mask]
        k_sorted = kabs[mask][idx]
        ils_sorted = np.ones_like(wavesel) if ils is None else np.asarray(ils(ib, wavesel[idx]), dtype=np.float64)
        delvarray = np.zeros_like(k_sorted) + (wavecalc[1] - wavecalc[0])
        g_sorted = np.cumsum(ils_sorted * delvarray) / np.sum(ils_sorted * delvarray)
        out[ib] = np.interp(g_ord, g_sorted, k_sorted)
    return out
