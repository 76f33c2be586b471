"""Continuum opacities (collision-induced absorption, Rayleigh scattering, aerosol extinction) as a device plan.

The reference builds TAUCIA / TAURAY / TAUDUST [NWAVE,NLAY] and their gradients as dense host arrays on every
evaluation (calc_tau_cia, archnemesis/ForwardModel_0.py:4516-4788; calc_tau_rayleigh*, :4869-4937, :5524-5835;
calc_tau_dust, :4790-4867) and calculate_layer_opacity folds them into dTAUCON[NWAVE,NPAR,NLAY] (:3938-3981).  With
the hot path on the device those arrays were the whole per-evaluation host->device traffic (36 MB at the config-2
size, 386 MB on a 10^5-wavenumber line-by-line grid).  All of them are products of something per wavenumber and
something per layer:

* CIA: for every pair of gases a cross-section table K_CIA[pair, para, T, wavenumber]; a layer picks four
  (para, T) planes and five interpolation weights, and the pair's opacity is  k(w; planes, weights) * q1 * q2 * XFAC
  (:4618-4750).  The table is interpolated ONCE onto the calculation wavenumbers (SciPy's linear interp1d, the
  reference's own call) and kept resident: `kw[term, plane, wave]`.  The fixed spectra of co2cia / n2n2cia / n2h2cia
  (:4752-4771) are further terms with a single plane.
* Rayleigh: cross-section(w) * TOTAM(l)  (gas-giant and CO2 formulae), or sum over gases of
  cross-section_gas(w) * column_gas(l) -- `ur[r, wave] * vr[r, layer]`.
* Aerosols: kext_i(w) * 1e-4 * CONT(l, i), clipped like the reference (:3960) -- `ud[i, wave] * vd[i, layer]`.

`build_plan` evaluates the reference's own per-layer logic (bracket search, mixing ratios, pair look-up, INORMAL
switch, the NVMR-2 slot the reference writes the CIA temperature derivative to, Rayleigh added to every gas slot)
on the host -- O(NLAY * NPAIR) work -- and returns the arrays `ansb200_continuum` (csrc/continuum.cu) turns into the
dense device arrays the radiance kernels read.  Per evaluation a few hundred KB cross PCIe instead of tens of MB.
"""
import numpy as np

MAX_SLOTS = 24          # csrc/continuum.cu: gradient slots (NVMR + 2) held per thread

_WAVENUMBER, _WAVELENGTH = 0, 1


class ContinuumTables:
    """The state-independent, large part: the CIA cross sections on the calculation wavenumbers.  Built once per
    (CIA table, wavenumber grid, gas list) and kept by the caller (forward_model caches it beside the resident k-table);
    the Rayleigh and aerosol spectra (a few vectors) travel with every plan."""

    def __init__(self, kw, nplanes, meta):
        self.kw = np.ascontiguousarray(kw, dtype=np.float64)              # [NTERM, NPL, NWAVE]
        self.nplanes = np.ascontiguousarray(nplanes, dtype=np.int32)      # [NTERM]: NPL for table terms, 1 for fixed spectra
        self.meta = meta
        self.device = None          # engine-side resident copy

    @property
    def nbytes(self):
        return self.kw.nbytes


def _interp_rows(x, y2d, xq):
    """scipy.interpolate.interp1d(x, row)(xq) for every row of y2d -- the reference's call (:4709-4712)."""
    import scipy.interpolate
    out = np.empty((y2d.shape[0], len(xq)))
    for r in range(y2d.shape[0]):
        out[r] = scipy.interpolate.interp1d(x, y2d[r])(xq)
    return out


def cia_pairs(CIA, Atmosphere):
    """(term, igas1, igas2, active) of every CIA pair both of whose gases are in the atmosphere (:4690-4703, :4716-4735)."""
    ID, ISO = np.asarray(Atmosphere.ID), np.asarray(Atmosphere.ISO)
    inormald = CIA.locate_INORMAL_pairs()
    out = []
    for ipair in range(CIA.NPAIR):
        g1 = np.where(ID == CIA.IPAIRG1[ipair])[0]
        g2 = np.where(ID == CIA.IPAIRG2[ipair])[0]
        if len(g1) > 1:
            g1 = np.where((ID == CIA.IPAIRG1[ipair]) & (ISO == 1))[0]
        if len(g2) > 1:
            g2 = np.where((ID == CIA.IPAIRG2[ipair]) & (ISO == 1))[0]
        if len(g1) == 1 and len(g2) == 1:
            active = (CIA.INORMALT[ipair] == CIA.INORMAL) if inormald[ipair] else True
            out.append((ipair, int(g1[0]), int(g2[0]), bool(active)))
    return out


def special_gases(Atmosphere, gas_enum):
    """ico2, ih2, in2 as calc_tau_cia finds them (:4563-4583)."""
    ico2 = ih2 = in2 = -1
    for i in range(Atmosphere.NVMR):
        gid, iso = Atmosphere.ID[i], Atmosphere.ISO[i]
        if gid == gas_enum.H2 and iso in (0, 1):
            ih2 = i
        if gid == gas_enum.N2:
            in2 = i
        if gid == gas_enum.CO2 and iso in (0, 1):
            ico2 = i
    return ico2, ih2, in2


def build_tables(ISPACE, WAVEC, CIA, Atmosphere, ext_cia, gas_enum):
    """State-independent tables.  `ext_cia` = (co2cia, n2n2cia, n2h2cia) of archnemesis.CIA_0."""
    WAVEC = np.asarray(WAVEC, dtype=np.float64)
    NW = len(WAVEC)
    ispace = int(ISPACE)
    if ispace == _WAVENUMBER:
        WAVEN, isort = WAVEC, None
    else:
        WAVEN = 1.0e4 / WAVEC
        isort = np.argsort(WAVEN)
        WAVEN = WAVEN[isort]
    terms = []       # (kind, data...) in the reference's order of accumulation
    kws = []
    nplanes = []
    npl = 1
    if CIA is not None:
        in_range = (CIA.WAVEN.min() <= WAVEN.min()) & (CIA.WAVEN.max() >= WAVEN.max())
        kc = np.asarray(CIA.K_CIA, dtype=np.float64)                     # [NPAIR, NPARA', NT, NWAVEN]
        npl = kc.shape[1] * kc.shape[2]
        if in_range:
            for (ipair, g1, g2, active) in cia_pairs(CIA, Atmosphere):
                if not active:
                    continue
                kws.append(_interp_rows(CIA.WAVEN, kc[ipair].reshape(npl, -1), WAVEN))
                nplanes.append(npl)
                terms.append(("pair", g1, g2))
        ico2, ih2, in2 = special_gases(Atmosphere, gas_enum)
        co2cia, n2n2cia, n2h2cia = ext_cia
        for name, fn, ga, gb in (("co2", co2cia, ico2, ico2), ("n2n2", n2n2cia, in2, in2), ("n2h2", n2h2cia, in2, ih2)):
            if ga != -1 and gb != -1:
                k = np.zeros((npl, NW))
                k[0] = fn(WAVEN)
                kws.append(k)
                nplanes.append(1)
                terms.append((name, ga, gb))
    kw = np.stack(kws) if kws else np.zeros((0, npl, NW))
    if isort is not None and kw.size:
        kw = kw[:, :, isort]        # the reference's un-sort (:4778-4780: result[isort]), applied to the tables
    return ContinuumTables(kw, nplanes, dict(terms=terms, NWAVE=NW, npl=npl, has_cia=CIA is not None))


def cia_layer_weights(CIA, temp, frac):
    """Per layer: the four (para, T) planes and (fhh_temp, fhl_temp, fhh_frac, fhl_frac, dfhldT), :4618-4686."""
    NLAY = len(temp)
    NT = CIA.NT
    T = np.asarray(CIA.TEMP)
    F = np.asarray(CIA.FRAC)
    pl = np.zeros((NLAY, 4), dtype=np.int32)
    wt = np.zeros((NLAY, 5))
    for ilay in range(NLAY):
        temp1 = temp[ilay]
        it = np.argmin(np.abs(T - temp1))
        if T[it] >= temp1:
            ithi = it
            if it == 0:
                temp1 = T[it]
                itl = 0
                ithi = 1
            else:
                itl = it - 1
        elif T[it] < temp1:
            itl = it
            if it == NT - 1:
                temp1 = T[it]
                ithi = NT - 1
                itl = NT - 2
            else:
                ithi = it + 1
        else:               # NaN temperature: the reference raises UnboundLocalError here
            raise ValueError("calc_tau_cia: layer temperature is not a number")
        frac1 = frac[ilay]
        ip = np.argmin(np.abs(F - frac1))
        if F[ip] >= frac1:
            iphi = ip
            if ip == 0:
                frac1 = F[ip]
                ipl = 0
                iphi = 1
            else:
                ipl = ip - 1
        elif F[ip] < frac1:
            ipl = ip
            if ip == CIA.NPARA - 1:
                iphi = CIA.NPARA - 1
                ipl = CIA.NPARA - 2
            else:
                iphi = ip + 1
        else:
            raise ValueError("calc_tau_cia: para-H2 fraction is not a number")
        if CIA.NPARA == 0:
            ipl = 0
            iphi = 0
        fhl_t = (temp1 - T[itl]) / (T[ithi] - T[itl])
        fhh_t = (T[ithi] - temp1) / (T[ithi] - T[itl])
        dfhldT = 1.0 / (T[ithi] - T[itl])
        if len(F) > 1:
            fhl_f = (frac1 - F[ipl]) / (F[iphi] - F[ipl])
            fhh_f = (F[iphi] - frac1) / (F[iphi] - F[ipl])
        else:
            fhl_f = 0.5
            fhh_f = 0.5
        pl[ilay] = (ipl * NT + itl, ipl * NT + ithi, iphi * NT + itl, iphi * NT + ithi)
        wt[ilay] = (fhh_t, fhl_t, fhh_f, fhl_f, dfhldT)
    return pl, wt


def build_plan(tables, CIA, Atmosphere, Layer, NDUST, rayleigh, dust_spectra, sq_cm_to_sq_m=1.0e-4):
    """The per-evaluation part.  `rayleigh` = (ur[NR,NWAVE], vr[NR,NLAY], vrd[NR,NLAY]) or None: TAURAY = sum_r ur*vr,
    dTAURAY = sum_r ur*vrd.  `dust_spectra` = ud[NDUST,NWAVE] = kext * 1e-4 as calc_tau_dust interpolates it (or None);
    the layer factor is Layer.CONT."""
    NLAY, NVMR = int(Layer.NLAY), int(Atmosphere.NVMR)
    NS = NVMR + 2
    if NS > MAX_SLOTS:
        return None
    terms = tables.meta["terms"]
    NTERM = len(terms)
    q = np.transpose(np.asarray(Layer.PP).T / np.asarray(Layer.PRESS))           # (NLAY, NVMR), :4559
    TOTAM_cm = np.asarray(Layer.TOTAM) * sq_cm_to_sq_m
    XLEN = np.asarray(Layer.DELH) * 1.0e2
    XFAC = TOTAM_cm ** 2. / XLEN
    if tables.meta["has_cia"]:
        pl, wt = cia_layer_weights(CIA, np.asarray(Layer.TEMP), np.asarray(Layer.FRAC))
    else:
        pl, wt = np.zeros((NLAY, 4), dtype=np.int32), np.zeros((NLAY, 5))
    q1 = np.zeros((NTERM, NLAY))
    q2 = np.zeros((NTERM, NLAY))
    slots = np.full((NTERM, 3), -1, dtype=np.int32)      # slot of d/dq1-type, d/dq2-type and dk/dT terms
    ca = np.zeros((NTERM, NLAY))                          # coefficient of k in slot A
    cb = np.zeros((NTERM, NLAY))                          # coefficient of k in slot B
    for t, (kind, ga, gb) in enumerate(terms):
        q1[t], q2[t] = q[:, ga], q[:, gb]
        if kind == "pair":
            slots[t] = (ga, gb, NVMR - 2)                 # (:4741: the T derivative goes to slot NVMR-2, sic)
            ca[t], cb[t] = q[:, gb], q[:, ga]
        elif kind in ("co2", "n2n2"):
            slots[t] = (ga, -1, -1)
            ca[t] = 2. * q[:, ga]
        else:                                             # n2h2: ga = N2, gb = H2 (:4766-4770)
            slots[t] = (gb, ga, -1)
            ca[t], cb[t] = q[:, ga], q[:, gb]
    NW = tables.meta["NWAVE"]
    ur, vr, vrd = rayleigh if rayleigh is not None else (np.zeros((0, NW)), np.zeros((0, NLAY)), np.zeros((0, NLAY)))
    ud = np.atleast_2d(dust_spectra) if (dust_spectra is not None and NDUST > 0) else np.zeros((0, NW))
    vd = np.ascontiguousarray(np.asarray(Layer.CONT)[:, :NDUST].T, dtype=np.float64) if NDUST > 0 else np.zeros((0, NLAY))
    return dict(NLAY=NLAY, NVMR=NVMR, NDUST=int(NDUST), NPAR=NVMR + 2 + int(NDUST), NTERM=NTERM,
                pl=pl, wt=wt, q1=q1, q2=q2, slots=slots, ca=ca, cb=cb, xfac=np.ascontiguousarray(XFAC),
                totam=np.ascontiguousarray(np.asarray(Layer.TOTAM, dtype=np.float64)),
                ur=np.ascontiguousarray(np.atleast_2d(ur), dtype=np.float64), ud=np.ascontiguousarray(ud, dtype=np.float64),
                vr=np.ascontiguousarray(np.atleast_2d(vr), dtype=np.float64),
                vrd=np.ascontiguousarray(np.atleast_2d(vrd), dtype=np.float64), vd=vd,
                has_cia=bool(tables.meta["has_cia"]))


def plan_bytes(plan):
    return int(sum(v.nbytes for v in plan.values() if isinstance(v, np.ndarray)))
