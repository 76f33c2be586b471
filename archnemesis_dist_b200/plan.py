"""Host-side plans: the tiny, dtype-sensitive pieces the reference computes per call, reproduced
with the reference's own numpy dtypes and handed to the kernels.

Nothing here touches the large arrays; it is O(NLAY), O(NG^2) or O(NPAR*NLAYIN*NPRO*NX) host work.
"""
import numpy as np


def _bracket(grid, x):
    """Nearest-node bracket with clamping exactly as archnemesis/Spectroscopy_0.py:2336-2371
    (calc_k) and :2182-2218 (calc_kg) do it.  Returns (lo, clamp) where clamp is the grid value the
    coordinate is clamped to (in the grid's dtype) or None."""
    n = len(grid)
    i = int(np.abs(grid - x).argmin())
    if grid[i] >= x:
        if i == 0:
            return 0, grid[0]
        return i - 1, None
    if i == n - 1:
        return n - 2, grid[n - 1]
    return i, None


def _bracket_all(grid, x):
    """_bracket for every layer at once: lo[n], clamped-to-first[n], clamped-to-last[n]."""
    n = len(grid)
    i = np.abs(grid[None, :] - x[:, None]).argmin(axis=1)
    ge = grid[i] >= x
    lo = np.where(ge, np.maximum(i - 1, 0), np.minimum(i, n - 2))
    return lo, ge & (i == 0), (~ge) & (i == n - 1)


def kinterp_plan(PRESS, TEMP, press, temp, grad):
    """Per-layer bracket indices and bilinear weights for calc_k (grad=False,
    Spectroscopy_0.py:2331-2389) or calc_kg (grad=True, :2176-2236).

    PRESS/TEMP are taken from the live Spectroscopy object (float32 after a ``.kta`` read,
    float64 after HDF5): numpy's scalar promotion then decides whether v, u and the four weight
    products are rounded to float32 before they multiply the float64 table -- that rounding is part
    of the reference's result, so it is reproduced here and the weights are widened afterwards.
    The reference works layer by layer on scalars; here the layers are grouped by which of the two
    coordinates is clamped to a grid edge (a clamped coordinate is a grid-dtype scalar, so v or u --
    and, if both are clamped, the weight products -- are evaluated in the grid's dtype) and every
    group is evaluated with arrays of exactly those dtypes: bit-identical to the scalar loop
    (tests/test_plan.py against the oracle's scalar restatement), ~10x less host time per evaluation.
    calc_k clamps the pressure and then takes the log, calc_kg takes the log first and replaces it by
    log(PRESS[edge]): both end up with the log of a grid-dtype scalar when clamped.
    """
    PRESS = np.asarray(PRESS)
    TEMP = np.asarray(TEMP)
    press = np.asarray(press, dtype=np.float64)
    temp = np.asarray(temp, dtype=np.float64)
    n = len(press)
    ipl, p_first, p_last = _bracket_all(PRESS, press)
    itl, t_first, t_last = _bracket_all(TEMP, temp)
    pcl, tcl = p_first | p_last, t_first | t_last
    plo, phi = np.log(PRESS[ipl]), np.log(PRESS[ipl + 1])            # grid dtype
    tlo, thi = TEMP[itl], TEMP[itl + 1]
    pedge = np.where(p_first, PRESS[0], PRESS[len(PRESS) - 1])        # grid dtype
    tedge = np.where(t_first, TEMP[0], TEMP[len(TEMP) - 1])
    w4 = np.zeros((n, 4), np.float64)
    omv = np.zeros(n, np.float64)
    vv = np.zeros(n, np.float64)
    for pc in (False, True):
        for tc in (False, True):
            m = (pcl == pc) & (tcl == tc)
            if not m.any():
                continue
            lp = np.log(pedge[m]) if pc else np.log(press[m])       # grid dtype if clamped, else float64
            t1 = tedge[m] if tc else temp[m]
            v = (lp - plo[m]) / (phi[m] - plo[m])
            u = (t1 - tlo[m]) / (thi[m] - tlo[m])
            w4[m, 0] = (1.0 - v) * (1.0 - u)
            w4[m, 1] = v * (1.0 - u)
            w4[m, 2] = v * u
            w4[m, 3] = (1.0 - v) * u
            omv[m] = 1.0 - v
            vv[m] = v
    dudt = (1. / (thi - tlo)).astype(np.float64)
    return dict(ip_lo=ipl.astype(np.int32), it_lo=itl.astype(np.int32), w4=w4, omv=omv, vv=vv, dudt=dudt)


def klbl_plan(PRESS, TEMP, press, temp, grad):
    """Per-layer corner planes and weights for calc_klbl (grad=False, Spectroscopy_0.py:1800-1850) or
    calc_klblg (grad=True, :1636-1684) on a line-by-line table K[NWAVE,NP,NT,NGAS].

    The reference works on scalars with the grids' own dtypes (``np.log(self.PRESS)`` stays float32 after a
    float32 read); every operation is repeated here on numpy scalars of those dtypes and widened at the end.
    TEMP may be 2-D (pressure-dependent temperature grids, NT < 0): the two pressure levels then bracket the
    temperature separately (it1/u1 on level ip, it2/u2 on level ip+1).  calc_klblg does not clamp ``it`` at
    zero, so a layer sitting exactly on (or clamped to) the first temperature node indexes TEMP[-1] and
    K[:, ip, -1]: the plan carries the wrapped plane numbers so that the device reads what the reference reads.
    corner[NLAY,4] = ip*NT+it1, ip*NT+it1+1, (ip+1)*NT+it2, (ip+1)*NT+it2+1."""
    LP = np.log(np.asarray(PRESS))
    TEMP = np.asarray(TEMP)
    press = np.asarray(press, dtype=np.float64)
    temp = np.asarray(temp, dtype=np.float64)
    if TEMP.ndim == 1:
        return _klbl_plan_grouped(LP, TEMP, press, temp, grad)
    return _klbl_plan_scalar(LP, TEMP, press, temp, grad)


def _klbl_plan_grouped(LP, TEMP, press, temp, grad):
    """klbl_plan for a single temperature grid, all layers at once.  A clamped coordinate is a grid-dtype scalar in
    the reference, so v or u -- and, if both are clamped, the weight products -- are evaluated in the grid's dtype:
    the layers are grouped by which coordinates are clamped and every group is evaluated with arrays of exactly
    those dtypes (same device as kinterp_plan; bit-identical to the scalar loop, tests/test_lbl_table.py)."""
    n, npg, ntg = len(press), len(LP), len(TEMP)
    pmin, pmax, tmin, tmax = np.min(LP), np.max(LP), np.min(TEMP), np.max(TEMP)
    lp = np.log(press)
    p_lo, p_hi = lp < pmin, lp > pmax
    t_lo, t_hi = temp < tmin, temp > tmax
    pcl, tcl = p_lo | p_hi, t_lo | t_hi
    pedge = np.where(p_lo, pmin, pmax)                       # grid dtype
    tedge = np.where(t_lo, tmin, tmax)
    corner = np.zeros((n, 4), np.int32)
    w4 = np.zeros((n, 4))
    omv, vv, du = np.zeros(n), np.zeros(n), np.zeros(n)
    for pc in (False, True):
        for tc in (False, True):
            m = (pcl == pc) & (tcl == tc)
            if not m.any():
                continue
            p_l = pedge[m] if pc else lp[m]
            t_l = tedge[m] if tc else temp[m]
            ip = np.clip(np.searchsorted(LP, p_l) - 1, 0, npg - 2)
            v = (p_l - LP[ip]) / (LP[ip + 1] - LP[ip])
            it = np.searchsorted(TEMP, t_l) - 1
            if not grad:
                it = np.maximum(it, 0)
            it = np.minimum(it, ntg - 2)                     # it = -1 survives with gradients: Python's index wrap
            u = (t_l - TEMP[it]) / (TEMP[it + 1] - TEMP[it])
            d = 1. / (TEMP[it + 1] - TEMP[it])
            lo, hi = it % ntg, (it + 1) % ntg
            corner[m] = np.stack([ip * ntg + lo, ip * ntg + hi, (ip + 1) * ntg + lo, (ip + 1) * ntg + hi], axis=1)
            w4[m, 0] = (1.0 - v) * (1.0 - u)
            w4[m, 1] = v * (1.0 - u)
            w4[m, 2] = v * u
            w4[m, 3] = (1.0 - v) * u
            omv[m], vv[m], du[m] = 1.0 - v, v, d
    return dict(corner=corner, w4=w4, omv=omv, vv=vv, du1dt=du, du2dt=du.copy())


def _klbl_plan_scalar(LP, TEMP, press, temp, grad):
    """klbl_plan layer by layer on numpy scalars, literally as the reference does it (any TEMP layout)."""
    n, npg, ntg = len(press), len(LP), TEMP.shape[-1]
    pmin, pmax, tmin, tmax = np.min(LP), np.max(LP), np.min(TEMP), np.max(TEMP)
    corner = np.zeros((n, 4), np.int32)
    w4 = np.zeros((n, 4))
    omv, vv, du1dt, du2dt = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)

    def bracket(T, t):
        it = int(np.searchsorted(T, t)) - 1
        if it < 0 and not grad:
            it = 0
        if it >= len(T) - 1:
            it = len(T) - 2
        return it, (t - T[it]) / (T[it + 1] - T[it]), 1. / (T[it + 1] - T[it])

    for i in range(n):
        p_l = np.log(press[i])
        if p_l < pmin:
            p_l = pmin
        if p_l > pmax:
            p_l = pmax
        t_l = temp[i]
        if t_l < tmin:
            t_l = tmin
        if t_l > tmax:
            t_l = tmax
        ip = min(max(int(np.searchsorted(LP, p_l)) - 1, 0), npg - 2)
        v = (p_l - LP[ip]) / (LP[ip + 1] - LP[ip])
        Tn, Tn2 = (TEMP[ip], TEMP[ip + 1]) if TEMP.ndim == 2 else (TEMP, TEMP)
        it1, u1, d1 = bracket(Tn, t_l)
        it2, u2, d2 = bracket(Tn2, t_l)
        corner[i] = (ip * ntg + it1 % ntg, ip * ntg + (it1 + 1) % ntg,
                     (ip + 1) * ntg + it2 % ntg, (ip + 1) * ntg + (it2 + 1) % ntg)
        w4[i] = ((1.0 - v) * (1.0 - u1), v * (1.0 - u2), v * u2, (1.0 - v) * u1)
        omv[i], vv[i], du1dt[i], du2dt[i] = 1.0 - v, v, d1, d2
    return dict(corner=corner, w4=w4, omv=omv, vv=vv, du1dt=du1dt, du2dt=du2dt)


def planes_touched(plan, NT):
    """Number of distinct (ip,it) table planes the plan references (U of SURVEY.md 8d)."""
    s = set()
    for ip, it in zip(plan["ip_lo"], plan["it_lo"]):
        for a in (0, 1):
            for b in (0, 1):
                s.add((int(ip) + a) * NT + int(it) + b)
    return len(s)


def overlap_tables(del_g):
    """weight[NG*NG], g_ord[NG+1] and the need for the literal sequential bin scan.

    ``del_g[i]*del_g[j]`` is evaluated in del_g's dtype (a float32 product on the ``.kta`` path,
    archnemesis/ForwardModel_0.py:6087) and ``np.cumsum(del_g)`` likewise (:6141-6143; numba keeps a
    float32 running sum), then both are widened.  If a single weight can span two bin edges the
    parallel bin search is not equivalent to the reference's loop and ``seq`` is True.
    """
    del_g = np.asarray(del_g)
    ng = len(del_g)
    weight = np.outer(del_g, del_g).astype(np.float64).reshape(-1)
    g_ord = np.zeros(ng + 1, np.float64)
    g_ord[1:] = np.cumsum(del_g)
    g_ord[ng] = 1.0
    widths = np.diff(g_ord)
    seq = bool(widths.min() <= 0.0 or weight.max() * (1.0 + 1e-9) >= widths.min())
    return weight, g_ord, seq


def included_params(xmap):
    """Parameters that map2pro processes: archnemesis/ForwardModel_0.py:699-702."""
    return [i for i in range(xmap.shape[1]) if np.mean(xmap[:, i, :]) != 0.0]


def fold_projection(xmap, LAYINC, NLAYIN, DTE, DAM, DCO, NVMR, NDUST):
    """M[NPATH, NPAR*NLAYMAX, NX] with  M[p, k*NLAYMAX+j, x] = sum_pro D_k[LAYINC[j,p],pro] xmap[x,k,pro].

    Folds map2pro (archnemesis/ForwardModel_0.py:5357-5381) and map2xvec (:5422), including the
    reference's quirks: only parameters whose xmap mean is non-zero are mapped (:699-702), layers
    beyond NLAYIN are zero (dSPECOUT is zero there), and the para-H2 slot re-uses the previous
    parameter's product (:5378-5381 leave dSPECOUT1 stale).
    """
    NX, NPAR, NPRO = xmap.shape
    NLM, NPATH = LAYINC.shape
    M = np.zeros((NPATH, NPAR, NLM, NX))
    incpar = included_params(xmap)
    for ipath in range(NPATH):
        n = int(NLAYIN[ipath])
        rows = LAYINC[0:n, ipath]
        prev = None   # (source parameter, matrix) of the last tensordot, as in the reference loop
        for p in incpar:
            if p <= NVMR - 1:
                prev = (p, DAM[rows, :])
            elif p <= NVMR:
                prev = (p, DTE[rows, :])
            elif (p > NVMR) and (p <= NVMR + NDUST):
                prev = (p, DCO[rows, :])
            if prev is None:
                raise UnboundLocalError("map2pro: para-H2 gradient requested before any other parameter "
                                        "(the reference fails the same way, ForwardModel_0.py:5381)")
            src, D = prev
            # contribution of profile parameter p to state element x, driven by layer gradients of `src`
            M[ipath, src, 0:n, :] += D @ xmap[:, p, :].T
    return np.ascontiguousarray(M.reshape(NPATH, NPAR * NLM, NX))


def fold_projection_layers(xmap, NLAY, DTE, DAM, DCO, NVMR, NDUST):
    """The layer-space form of fold_projection: M[1, NPAR*NLAY, NX] with M[0, k*NLAY+l, x] = sum_pro D_k[l,pro]
    xmap[x,k,pro] -- what every path's M is before its rows are gathered by LAYINC.  Used with gradients per layer
    (ansb200_radiance with ANSB200_RAD_LAYER_SPACE): sum_j dspec[k,j] M_path[(k,j)] = sum_l (sum of the visits of l) M[(k,l)]."""
    ident = np.arange(NLAY, dtype=np.int32).reshape(NLAY, 1)
    return fold_projection(xmap, ident, np.array([NLAY], dtype=np.int32), DTE, DAM, DCO, NVMR, NDUST)


SPARSE_LONG = 16          # columns with more entries are summed by a whole warp (csrc/project.cu PS_LONG)


def sparse_projection(M, max_density=0.15, max_fill=4):
    """A folded projection matrix M[P, E, NX] (fold_projection / fold_projection_layers) by columns, for
    ansb200_jacobian_project_sparse.  A state-vector element acts on the layers of one parameter, so a column of M is
    one run of consecutive rows: col_r0 / col_len / col_voff[P*NX] give the run of column x of path p (rows r0 ..
    r0+len-1, values vals[voff : voff+len], zeros inside the run included), long_cols / long_ptr[P+1] the columns
    longer than SPARSE_LONG.  Returns None when M is not like that (a column spread over distant rows, or M too dense
    for the gather to beat the tiled product)."""
    M = np.asarray(M, dtype=np.float64)
    P, E, NX = M.shape
    nz = M != 0.0
    if nz.mean() > max_density:
        return None
    any_nz = nz.any(axis=1)                                   # [P, NX]
    first = np.where(any_nz, nz.argmax(axis=1), 0)
    last = np.where(any_nz, E - 1 - nz[:, ::-1, :].argmax(axis=1), -1)
    length = (last - first + 1).astype(np.int64)
    count = nz.sum(axis=1)
    if np.any(length > np.maximum(SPARSE_LONG, max_fill * count)) or length.sum() >= 2 ** 31:
        return None
    voff = np.concatenate([[0], np.cumsum(length.reshape(-1))[:-1]])
    vals = np.empty(int(length.sum()))
    for p in range(P):
        for x in np.nonzero(any_nz[p])[0]:
            o = voff[p * NX + x]
            vals[o:o + length[p, x]] = M[p, first[p, x]:last[p, x] + 1, x]
    longs = [np.nonzero(length[p] > SPARSE_LONG)[0].astype(np.int32) for p in range(P)]
    long_ptr = np.concatenate([[0], np.cumsum([len(q) for q in longs])]).astype(np.int32)
    return dict(col_r0=first.astype(np.int32).reshape(-1), col_len=length.astype(np.int32).reshape(-1),
                col_voff=voff.astype(np.int32), vals=vals, long_cols=np.concatenate(longs), long_ptr=long_ptr,
                shape=(P, E, NX))


def tangent_mix(BASEH_TANHE, TANHE):
    """The pair of paths and the weights with which nemesisSOfmg / nemesisLfmg interpolate the path spectra to each
    measured tangent height (ForwardModel_0.py:1206-1228, :1464-1486), as arrays lo, hi (-1: path lo alone), wlo, whi:
    SPECMOD[:, i] = SPECOUT[:, lo] * wlo + SPECOUT[:, hi] * whi.  Same expressions as the reference, evaluated once on
    the host; the reference's indices can leave the path range (a tangent height below the lowest path), in which case
    it raises IndexError or wraps around -- here that is a ValueError."""
    base = np.asarray(BASEH_TANHE, dtype=np.float64)
    tan = np.asarray(TANHE, dtype=np.float64).reshape(-1)
    n, npath = len(tan), len(base)
    lo, hi = np.zeros(n, np.int32), np.full(n, -1, np.int32)
    wlo, whi = np.ones(n), np.zeros(n)
    for i in range(n):
        ibase = int(np.argmin(np.abs(base - tan[i])))
        if base[ibase] <= tan[i]:
            il, ih = ibase, ibase + 1
        else:
            il, ih = ibase - 1, ibase
        if il < 0:
            raise ValueError("tangent height %g km below the lowest path (%g km)" % (tan[i], base.min()))
        lo[i] = il
        if ih <= npath - 1:
            fhl = (tan[i] - base[il]) / (base[ih] - base[il])
            fhh = (base[ih] - tan[i]) / (base[ih] - base[il])
            hi[i], wlo[i], whi[i] = ih, 1. - fhl, 1. - fhh
    return dict(lo=lo, hi=hi, wlo=wlo, whi=whi)


# ------------------------------------------------------------------------------------------------
# Instrument line shape as a sparse operator on the wavenumber axis (Measurement_0.conv / convg,
# archnemesis/Measurement_0.py:2288-2465 / :2467-2692), for the two modes convg supports with k-tables.
# ------------------------------------------------------------------------------------------------
CONV_INTERP, CONV_FILTER = 0, 1


def conv_operator(Wave, VCONV, FWHM, NFIL=None, VFIL=None, AFIL=None):
    """The convolution of one geometry as a sparse operator on the wavenumber axis.

    FWHM == 0 (:2632-2640): ``scipy.interpolate.interp1d(Wave, y, axis=0)(VCONV)``.  SciPy (1.18) evaluates a
    2-D y (the gradients) with ``idx = searchsorted(x, x_new).clip(1, n-1)`` and
    ``(x_new-x_lo)/(x_hi-x_lo) * y_hi + (x_hi-x_new)/(x_hi-x_lo) * y_lo``: each row holds (hi, lo) with those
    two weights, summed in that order.  A 1-D float64 y (the spectrum) is delegated to ``np.interp``, whose
    kernel is ``slope = (y[j+1]-y[j])/(x[j+1]-x[j]); slope*(x_new-x[j]) + y[j]`` with x[j] <= x_new < x[j+1] and
    y[j] itself on a knot: ``np_lo`` = j, ``np_exact`` marks the knots, ``xinfo`` = (x[j], x[j+1], x_new).
    FWHM < 0 (:2642-2690): filter-weighted mean ``sum(f1*y)/sum(f1)`` over the calculation points between the
    last one below VFIL[0] and the first one above VFIL[NFIL-1], ``f1 = np.interp(Wave[i], VFIL, AFIL)``, only
    where f1 > 0, accumulated in ascending order; ``norm`` is that sum of f1.
    FWHM > 0 with k-tables raises in the reference's convg (:2617-2619) and here."""
    Wave = np.asarray(Wave, dtype=np.float64)
    VCONV = np.asarray(VCONV, dtype=np.float64)
    nconv = len(VCONV)
    if FWHM > 0.0:
        raise ValueError("conv_operator: FWHM>0 is not available for k-table Jacobians (Measurement_0.convg raises)")
    if FWHM == 0.0:
        if VCONV.min() < Wave[0] or VCONV.max() > Wave[-1]:
            raise ValueError("A value in x_new is outside the interpolation range.")     # interp1d bounds_error
        n = len(Wave)
        idx = np.searchsorted(Wave, VCONV).clip(1, n - 1)
        x_lo, x_hi = Wave[idx - 1], Wave[idx]
        widx = np.stack([idx, idx - 1], axis=1).astype(np.int32).reshape(-1)
        wval = np.stack([(VCONV - x_lo) / (x_hi - x_lo), (x_hi - VCONV) / (x_hi - x_lo)], axis=1).reshape(-1)
        # np.interp bracket: x[j] <= x_new < x[j+1]; the last knot and exact hits return y[j]
        j = np.searchsorted(Wave, VCONV, side="right") - 1
        exact = (Wave[j.clip(0, n - 1)] == VCONV) | (j >= n - 1)
        j = j.clip(0, n - 2)
        j = np.where(exact & (VCONV == Wave[n - 1]), n - 1, j)
        jn = np.minimum(j + 1, n - 1)
        xinfo = np.stack([Wave[j], Wave[jn], VCONV], axis=1)
        return dict(mode=CONV_INTERP, row_start=(2 * np.arange(nconv + 1)).astype(np.int32), widx=widx,
                    wval=np.ascontiguousarray(wval), norm=np.ones(nconv), np_lo=j.astype(np.int32),
                    np_exact=exact.astype(np.int32), xinfo=np.ascontiguousarray(xinfo), NCONV=nconv)
    rows, vals, start, norm = [], [], [0], []
    for ic in range(nconv):
        nf = int(NFIL[ic])
        xp = np.asarray(VFIL[0:nf, ic], dtype=np.float64)
        yp = np.asarray(AFIL[0:nf, ic], dtype=np.float64)
        lo = np.where(Wave < xp[0])[0]
        hi = np.where(Wave > xp[nf - 1])[0]
        i0, i1 = int(lo[len(lo) - 1]), int(hi[0])
        tot = 0.0
        for i in range(i0, i1 + 1):
            f1 = float(np.interp(Wave[i], xp, yp))
            if f1 > 0.0:
                rows.append(i)
                vals.append(f1)
                tot = tot + f1
        start.append(len(rows))
        norm.append(tot)
    return dict(mode=CONV_FILTER, row_start=np.asarray(start, np.int32), widx=np.asarray(rows, np.int32),
                wval=np.asarray(vals, np.float64), norm=np.asarray(norm, np.float64),
                np_lo=np.zeros(nconv, np.int32), np_exact=np.zeros(nconv, np.int32), xinfo=np.zeros((nconv, 3)),
                NCONV=nconv)


def filter_integral_operator(Wave, NCONV, NFIL, VFIL, AFIL):
    """Measurement_0.integrate_filter / integrate_filterg (archnemesis/Measurement_0.py:4079-4250; IFORM =
    integrated radiance): ``np.trapz(y[i] * np.interp(Wave[i], VFIL, AFIL), Wave[i])`` over the calculation points
    inside the filter, as rows of a weighted sum (no normalisation): the coefficient of y_i is
    ``a_i * ((x_{i+1} - x_i) + (x_i - x_{i-1})) / 2`` with one-sided ends.  ``Wave`` is the Doppler-corrected grid
    the reference passes (``correct_doppler_shift``).  Same sum as numba's trapz up to the order of the
    additions (a few ulp)."""
    Wave = np.asarray(Wave, dtype=np.float64)
    rows, vals, start = [], [], [0]
    for ic in range(int(NCONV)):
        nf = int(NFIL[ic])
        xp = np.asarray(VFIL[0:nf, ic], dtype=np.float64)
        yp = np.asarray(AFIL[0:nf, ic], dtype=np.float64)
        idx = np.where((Wave >= xp[0]) & (Wave <= xp[nf - 1]))[0]
        if len(idx) >= 2:
            x = Wave[idx]
            a = np.interp(x, xp, yp)
            d = np.diff(x)
            c = np.zeros(len(idx))
            c[:-1] += d / 2.0
            c[1:] += d / 2.0
            rows.extend(idx.tolist())
            vals.extend((a * c).tolist())
        start.append(len(rows))
    n = int(NCONV)
    return dict(mode=CONV_INTERP, row_start=np.asarray(start, np.int32), widx=np.asarray(rows, np.int32),
                wval=np.asarray(vals, np.float64), norm=np.ones(n), np_lo=np.zeros(n, np.int32),
                np_exact=np.zeros(n, np.int32), xinfo=np.zeros((n, 3)), NCONV=n, weighted_sum_only=True)


ILS_SQUARE, ILS_TRIANGULAR, ILS_GAUSSIAN, ILS_HAMMING, ILS_HANNING = 0, 1, 2, 3, 4


def lbl_conv_operator(Wave, VCONV, FWHM, ISHAPE=ILS_GAUSSIAN, NFIL=None, VFIL=None, AFIL=None, grad=True):
    """Measurement_0.lblconv / lblconvg (archnemesis/Measurement_0.py:2125-2284 and the kernels :3335-4076) of one
    geometry as a sparse operator on the (Doppler-corrected) calculation grid ``Wave``.

    FWHM > 0: the analytic line shape ISHAPE of width FWHM around every convolution point: all calculation points
      with v1 <= Wave <= v2 contribute ``f1`` where f1 > 0, in ascending order, and the sum is divided by sum(f1).
      Square: v = vcen -+ FWHM/2 (v2 = v1 + FWHM), f1 = 1.  Triangular: +-FWHM, 1 - |dv|/FWHM.  Gaussian: sig =
      0.5 FWHM / sqrt(ln 2), +-3 sig, exp(-(dv/sig)^2).  Hamming: the reference's apodised sinc,
      a = 0.907/FWHM; the gradient kernels window it at +-FWHM (:3864-3866) while the spectrum-only kernel sets
      v1 = v2 = vcen - 1.1 FWHM (:3389-3391) -- ``grad`` selects which one is reproduced.  Hanning has no weight in
      the reference (f1 stays 0): the division by zero is raised here as ZeroDivisionError, like numba does.
      ``grad`` also selects the arithmetic of the weights: the gradient kernels are numba-compiled (libm exp / sin
      through ``math``, squares as products), the spectrum-only ``lblconv`` is plain Python on numpy scalars (its
      @jit is commented out) and returns NaN instead of raising where no point carries weight.
    FWHM < 0: the tabulated filter (VFIL, AFIL) of every convolution point, f1 = np.interp(Wave[i], VFIL, AFIL) for
      VFIL[0] <= Wave[i] <= VFIL[NFIL-1].  (The k-table convg takes one extra point either side; lblconv does not.)
    FWHM == 0: np.interp(VCONV, Wave, y) for the spectrum AND every gradient column (``np_interp_all``)."""
    import math
    Wave = np.asarray(Wave, dtype=np.float64)
    VCONV = np.asarray(VCONV, dtype=np.float64)
    nconv, n = len(VCONV), len(Wave)
    if FWHM == 0.0:
        # np.interp clamps outside the grid: the end values, flagged as exact hits
        j = np.searchsorted(Wave, VCONV, side="right") - 1
        below, above = VCONV < Wave[0], VCONV >= Wave[n - 1]
        exact = below | above | (Wave[j.clip(0, n - 1)] == VCONV)
        j = np.where(above, n - 1, j.clip(0, n - 2))
        jn = np.minimum(j + 1, n - 1)
        xinfo = np.stack([Wave[j], Wave[jn], VCONV], axis=1)
        return dict(mode=CONV_INTERP, row_start=np.zeros(nconv + 1, np.int32), widx=np.zeros(1, np.int32),
                    wval=np.zeros(1), norm=np.ones(nconv), np_lo=j.astype(np.int32), np_exact=exact.astype(np.int32),
                    xinfo=np.ascontiguousarray(xinfo), NCONV=nconv, np_interp_all=True)
    rows, vals, start, norm = [], [], [0], []
    for ic in range(nconv):
        vcen = VCONV[ic]
        if FWHM > 0.0:
            yfwhm = FWHM
            if ISHAPE == ILS_SQUARE:
                v1 = vcen - 0.5 * yfwhm
                v2 = v1 + yfwhm
            elif ISHAPE == ILS_TRIANGULAR:
                v1, v2 = vcen - yfwhm, vcen + yfwhm
            elif ISHAPE == ILS_GAUSSIAN:
                sig = 0.5 * yfwhm / np.sqrt(np.log(2.0))
                v1, v2 = vcen - 3. * sig, vcen + 3. * sig
            elif ISHAPE == ILS_HAMMING:
                v1, v2 = (vcen - yfwhm, vcen + yfwhm) if grad else (vcen - 1.1 * yfwhm, vcen - 1.1 * yfwhm)
            else:
                v1, v2 = vcen - 3. * yfwhm, vcen + 3. * yfwhm
        else:
            nf = int(NFIL[ic])
            xp = np.asarray(VFIL[0:nf, ic], dtype=np.float64)
            yp = np.asarray(AFIL[0:nf, ic], dtype=np.float64)
            v1, v2 = xp[0], xp[nf - 1]
        idx = np.where((Wave >= v1) & (Wave <= v2))[0]
        x = Wave[idx]
        if FWHM < 0.0:
            f = np.interp(x, xp, yp)
        elif ISHAPE == ILS_SQUARE:
            f = np.ones(len(idx))
        elif ISHAPE == ILS_TRIANGULAR:
            f = 1.0 - np.abs(x - vcen) / yfwhm
        elif ISHAPE == ILS_GAUSSIAN:
            if grad:
                # the compiled kernels: libm exp, and x**2.0 is a plain product (LLVM folds pow(x, 2) to x*x)
                f = np.array([math.exp(-(((xi - vcen) / sig) * ((xi - vcen) / sig))) for xi in x.tolist()],
                             dtype=np.float64)
            else:
                # the spectrum-only lblconv runs as plain Python (its @jit is commented out, :3334): numpy scalars
                f = np.array([np.exp(-((xi - vcen) / sig) ** 2.0) for xi in x], dtype=np.float64)
        elif ISHAPE == ILS_HAMMING:
            a = 0.907 / yfwhm
            f = np.zeros(len(idx))
            for q, xi in enumerate(x.tolist() if grad else x):
                k = xi - vcen
                if k == 0.0:
                    f[q] = a * 1.08
                elif grad:
                    num = a * (1.08 - (0.64 * (a * a) * (k * k))) * math.sin(2 * np.pi * a * k)
                    den = (1 - 4 * (a * a) * (k * k)) * (2 * np.pi * a * k)
                    f[q] = num / den
                else:
                    num = a * (1.08 - (0.64 * a**2 * k**2)) * np.sin(2 * np.pi * a * k)
                    den = (1 - 4 * a**2 * k**2) * (2 * np.pi * a * k)
                    f[q] = num / den
        else:
            f = np.zeros(len(idx))
        keep = f > 0.0
        tot = 0.0
        for fv in f[keep].tolist():
            tot = tot + fv
        if tot == 0.0 and (grad or FWHM < 0.0):
            # numba raises on the scalar 0/0; the plain-Python lblconv returns NaN (norm 0 and no entries: the kernel's 0/0)
            raise ZeroDivisionError("lbl_conv_operator: no positive line-shape weight at convolution point %d" % ic)
        rows.extend(idx[keep].tolist())
        vals.extend(f[keep].tolist())
        start.append(len(rows))
        norm.append(tot)
    return dict(mode=CONV_FILTER, row_start=np.asarray(start, np.int32), widx=np.asarray(rows, np.int32),
                wval=np.asarray(vals, np.float64), norm=np.asarray(norm, np.float64),
                np_lo=np.zeros(nconv, np.int32), np_exact=np.zeros(nconv, np.int32), xinfo=np.zeros((nconv, 3)),
                NCONV=nconv)
