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


def kinterp_plan(PRESS, TEMP, press, temp, grad):
    """Per-layer bracket indices and bilinear weights for calc_k (grad=False,
    Spectroscopy_0.py:2331-2389) or calc_kg (grad=True, :2176-2236).

    PRESS/TEMP are taken from the live Spectroscopy object (float32 after a ``.kta`` read,
    float64 after HDF5): numpy's scalar promotion then decides whether v, u and the four weight
    products are rounded to float32 before they multiply the float64 table -- that rounding is part
    of the reference's result, so it is reproduced here and the weights are widened afterwards.
    """
    PRESS = np.asarray(PRESS)
    TEMP = np.asarray(TEMP)
    press = np.asarray(press)
    temp = np.asarray(temp)
    n = len(press)
    ip_lo = np.zeros(n, np.int32)
    it_lo = np.zeros(n, np.int32)
    w4 = np.zeros((n, 4), np.float64)
    omv = np.zeros(n, np.float64)
    vv = np.zeros(n, np.float64)
    dudt = np.zeros(n, np.float64)
    for l in range(n):
        press1 = press[l]
        temp1 = temp[l]
        ipl, pclamp = _bracket(PRESS, press1)
        itl, tclamp = _bracket(TEMP, temp1)
        if grad:
            # calc_kg takes the log first and replaces it by log(PRESS[edge]) when clamped
            lpress = np.log(press1) if pclamp is None else np.log(pclamp)
        else:
            # calc_k clamps the pressure, then takes the log
            lpress = np.log(press1 if pclamp is None else pclamp)
        if tclamp is not None:
            temp1 = tclamp
        plo = np.log(PRESS[ipl])
        phi = np.log(PRESS[ipl + 1])
        tlo = TEMP[itl]
        thi = TEMP[itl + 1]
        v = (lpress - plo) / (phi - plo)
        u = (temp1 - tlo) / (thi - tlo)
        ip_lo[l] = ipl
        it_lo[l] = itl
        w4[l, 0] = (1.0 - v) * (1.0 - u)
        w4[l, 1] = v * (1.0 - u)
        w4[l, 2] = v * u
        w4[l, 3] = (1.0 - v) * u
        omv[l] = 1.0 - v
        vv[l] = v
        dudt[l] = 1. / (thi - tlo)
    return dict(ip_lo=ip_lo, it_lo=it_lo, w4=w4, omv=omv, vv=vv, dudt=dudt)


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
