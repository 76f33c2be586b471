"""Python face of the drop-in: the hot methods of archNEMESIS' ``ForwardModel_0`` re-routed to the
device through ``engine.HotPath``.

``B200HotPathMixin`` overrides the methods on the hot path and keeps their signatures and return
shapes:

    CIRSrad(return_grad=False)                       archnemesis/ForwardModel_0.py:4376-4511
    calculate_gaseous_line_opacity(return_grad)      :3781-3891   (k-tables and in-memory line-by-line tables)
    nemesisfmg()                                     :593-779     (CIRSrad -> map2pro -> map2xvec fused, JSURF column,
                                                     WGEOM and the instrument line shape on the device when the
                                                     geometry allows: b200_forward_jacobian[_conv])
    nemesisSOfmg() / nemesisLfmg()                   :983-1243 / :1372-1518  (all tangent paths in one evaluation,
                                                     tangent-height interpolation and the IGEOM='All' line shape on the
                                                     device: b200_tangent_fmg; the AOTF branch stays the reference's)
    jacobian_nemesis()                               :2184-2361   (numerical columns batched: StateBatch)
    map2pro / map2xvec (module functions)            :5319-5424   (rebound by install(): CIRSrad's gradient stays on
                                                     the device as a DeviceGradient and the projection is fused for
                                                     every other driver body -- process_IAV :2039-2059, the AOTF
                                                     branch, user code -- without overriding them)

``calculate_layer_opacity`` (:3905-4016), ``calculate_thermal_emission_spectrum`` and ``calculate_transmission_spectrum``
are NOT overridden: their work happens inside ``ansb200_radiance`` and is never materialised, so the device ``CIRSrad``
does not set the diagnostics the reference leaves behind on the way (``LayerX.TAUGAS`` :3925, ``LayerX.TAUTOT`` :3997,
``LayerX.TAUCIA`` :3901 stay as they were); ``calculate_gaseous_line_opacity`` still returns TAUGAS / dTAUGAS for
callers that want them.

The mix-in reads only the attributes the reference methods read (``SpectroscopyX``, ``LayerX``, ``PathX``,
``AtmosphereX``, ``SurfaceX``, ``MeasurementX``, ``ScatterX``, ``StellarX``, ``CIAX``, ``Variables``).  The continuum
terms (CIA, Rayleigh, aerosols) are planned on the host from those objects (continuum.py follows
``calc_tau_cia`` / ``calc_tau_rayleigh*`` / ``calc_tau_dust``) and evaluated on the device; ``ArrayForwardModel``, which
has no reference objects to plan from, takes them as dense arrays.  Path types that are out of scope (scattering,
absorption, emissions, run-time line-by-line inside CIRSrad, tables read on line from HDF5, shapes beyond the native
limits of include/ansb200.h) are delegated to the reference implementation further up the MRO, unchanged; the supported
path has no CPU fallback.

``install()`` builds ``ForwardModel_B200(B200HotPathMixin, archnemesis.ForwardModel_0)`` and rebinds the three names
callers resolve at call time, so ``coreretOE``, ``coreretNS`` and ``retrieval_nemesis`` pick it up
without edits; it also installs ``OE_B200`` (oe.py), the memoised ``read_tables`` with the vectorised table readers
(table_io.py), the run-time line-by-line functions (linedata.py) and ``calc_ktable_chunk`` (ktable.py).
``ArrayForwardModel`` is the same mix-in over plain namespaces for hosts where the reference package is not
installed.
"""
import sys
import types

import numpy as np

from . import _lib
from . import engine as _engine
from . import plan as _plan

# enum values of the reference (archnemesis/enum/*.py) as plain ints so that nothing here imports it
_K_TABLES = 0                 # SpectralCalculationModeEnum.K_TABLES
_LBL_TABLES = 2               # SpectralCalculationModeEnum.LINE_BY_LINE_TABLES
# PathCalcEnum is an IntFlag built with auto() (enum/path_calc_enum.py:3-25): bit = 1 << position
_THERMAL_EMISSION = 1 << 6
_MULTIPLE_SCATTERING = 1 << 8
_SINGLE_SCATTERING_PLANE_PARALLEL = 1 << 10
_ABSORBTION = 1 << 12
_NON_TRANSMISSION_BITS = _ABSORBTION | _THERMAL_EMISSION | _MULTIPLE_SCATTERING | _SINGLE_SCATTERING_PLANE_PARALLEL
_IFORM_INTEGRATED_RADIANCE = 6
_IFORM_FLUXRATIO = 1          # SpectraUnitEnum.FluxRatio
_IFORM_ATM_TRANSMISSION = 4   # SpectraUnitEnum.Atmospheric_transmission
_ATM_TO_PASCAL = 101325.0
_SQ_CM_TO_SQ_METER = 1.0e-4


class _TableCache:
    """Device-resident tables keyed on the identity of the host K array (plus shape): the
    reference re-reads and re-allocates its tables on every nemesisfm[g] call
    (ForwardModel_0.py:652-654); `install()` memoises read_tables (make_cached_read_tables) so the same
    host array -- and with it the device copy -- is reused when the files, wavemin and wavemax repeat."""

    def __init__(self, capacity=4):
        import threading
        self.capacity = capacity
        self.items = []   # (key, host K ref, HotPath), least recently used first
        self.lock = threading.Lock()       # (the forward models of a numerical Jacobian run in threads)

    def get(self, spec, factory):
        K = spec.K
        key = (id(K), K.shape, str(np.asarray(spec.PRESS).dtype), str(np.asarray(spec.DELG).dtype))
        with self.lock:
            for n, (k, ref, hp) in enumerate(self.items):
                if k == key and ref is K:
                    self.items.append(self.items.pop(n))      # a hit refreshes the entry
                    return hp
            try:
                hp = factory()
            except MemoryError:
                # device memory: drop the other resident tables and try once more
                while self.items:
                    self.items.pop(0)[2].close()
                hp = factory()
            self.items.append((key, K, hp))
            while len(self.items) > self.capacity:
                _, _, old = self.items.pop(0)
                old.close()
            return hp


_TABLES = _TableCache()


def make_cached_read_tables(ref_read_tables, capacity=4):
    """Spectroscopy_0.read_tables (archnemesis/Spectroscopy_0.py:1448-1528) memoised on what it depends on.

    Every nemesisfm[g] / nemesisSOfm[g] / nemesisLfm[g] call deep-copies the Spectroscopy object and re-reads the
    ``.kta`` / ``.lta`` files of every active gas from disk (ForwardModel_0.py:652-654; 7-8 s per call on the real
    Jupiter tables).  The outcome is a pure function of the table files, the header grid and the
    requested range, so it is kept: a repeat call installs the SAME (read-only) WAVE and K arrays on the new copy,
    and since the device table cache (_TableCache) is keyed on the identity of K, the resident device table is
    reused as well -- no disk read, no host re-assembly, no re-upload.  The key holds the files' size and mtime, so
    an edited table is read again.  Run-time line-by-line, on-line HDF5 tables and objects without LOCATION go
    straight to the reference method."""
    import os
    import threading
    cache = []      # [(key, WAVE, K)], least recently used first
    lock = threading.Lock()

    def stamp(path):
        for cand in (path, path + ".kta", path + ".lta", path + ".h5"):
            try:
                st = os.stat(cand)
                return (cand, st.st_size, st.st_mtime_ns)
            except OSError:
                continue
        return (path, None, None)

    def read_tables(self, wavemin=0., wavemax=1.0e10, wavedelta=1.0):
        ilbl = int(self.ILBL)
        if ilbl not in (_K_TABLES, _LBL_TABLES) or self.LOCATION is None or getattr(self, "ONLINE", False):
            return ref_read_tables(self, wavemin=wavemin, wavemax=wavemax, wavedelta=wavedelta)
        if self.WAVE is None:
            self.read_header()
        W = np.asarray(self.WAVE)
        key = (ilbl, tuple(stamp(str(f)) for f in self.LOCATION), float(wavemin), float(wavemax),
               len(W), float(W[0]) if len(W) else 0.0, float(W[-1]) if len(W) else 0.0,
               int(self.NG), int(self.NP), int(self.NT), int(self.NGAS))
        if any(st[1] is None for st in key[1]):
            # a table file that cannot be stat'ed (resolved some other way by the reader): nothing to tell a stale
            # entry by, so do not memoise
            return ref_read_tables(self, wavemin=wavemin, wavemax=wavemax, wavedelta=wavedelta)
        with lock:      # (one reader at a time: concurrent forward models of a numerical Jacobian share the result)
            for n, (k, wave, K) in enumerate(cache):
                if k == key:
                    cache.append(cache.pop(n))
                    self.WAVE, self.NWAVE, self.K = wave, len(wave), K
                    read_tables.hits += 1
                    return None
            out = ref_read_tables(self, wavemin=wavemin, wavemax=wavemax, wavedelta=wavedelta)
            if isinstance(self.K, np.ndarray) and isinstance(self.WAVE, np.ndarray):
                self.K.setflags(write=False)          # shared between evaluations from now on
                self.WAVE.setflags(write=False)
                cache.append((key, self.WAVE, self.K))
                del cache[:-capacity]
            return out

    read_tables.hits = 0
    read_tables.cache = cache
    read_tables.__doc__ = (ref_read_tables.__doc__ or "") + "\n(archnemesis_dist_b200: memoised, see make_cached_read_tables)"
    return read_tables


class DeviceGradient:
    """What CIRSrad(return_grad=True) returns in place of dSPECOUT[NWAVE,NPAR,NLAYIN,NPATH] once install() is
    active: the layer-space Jacobian stays on the device.  The reference's drivers (nemesisfmg, nemesisSOfmg,
    nemesisLfmg, process_IAV) hand it straight to the module functions ``map2pro`` and ``map2xvec``
    (ForwardModel_0.py:694-714, :1188-1206, :1452-1470, :2039-2059); install() rebinds those two names to
    wrappers that recognise this object and run the fused projection on the device, so the 4-D array (64 MB
    per path at config 2) never crosses PCIe.  Any other use -- indexing, numpy functions -- materialises it
    through ``__array__`` in the reference's layout, so code that looks at dSPECOUT directly still works."""

    def __init__(self, hotpath, dspec_dev, shape):
        self._hp, self._dev, self.shape = hotpath, dspec_dev, tuple(shape)
        self.ndim, self.dtype = 4, np.dtype(np.float64)
        self._host = None

    def materialise(self):
        if self._host is None:
            self._host = np.ascontiguousarray(np.transpose(self._hp.to_host(self._dev), (0, 2, 3, 1)))
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.materialise()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        return self.materialise()[idx]

    def __len__(self):
        return self.shape[0]


class ProfileGradient:
    """map2pro applied lazily to a DeviceGradient: remembers the layer->profile matrices so that map2xvec can
    fold them with xmap into one projection (plan.fold_projection) and run it on the device."""

    def __init__(self, grad, NVMR, NDUST, NPRO, NPATH, NLAYIN, LAYINC, DTE, DAM, DCO, INCPAR, fallback):
        self.grad, self.args, self.INCPAR, self._fallback = grad, (NVMR, NDUST, NPRO, NPATH, NLAYIN, LAYINC, DTE, DAM, DCO), INCPAR, fallback
        self.shape = (grad.shape[0], NVMR + 2 + NDUST, NPRO, NPATH)
        self.ndim, self.dtype = 4, np.dtype(np.float64)
        self._host = None

    def materialise(self):
        if self._host is None:
            NVMR, NDUST, NPRO, NPATH, NLAYIN, LAYINC, DTE, DAM, DCO = self.args
            self._host = self._fallback(self.grad.materialise(), self.grad.shape[0], NVMR, NDUST, NPRO, NPATH, NLAYIN,
                                        LAYINC, DTE, DAM, DCO, INCPAR=self.INCPAR)
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.materialise()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        return self.materialise()[idx]

    def __len__(self):
        return self.shape[0]


def make_fused_map_functions(ref_map2pro, ref_map2xvec):
    """Wrappers for the module-level map2pro / map2xvec of archnemesis.ForwardModel_0 (:5319-5424): identical to
    the reference for ordinary arrays; for a DeviceGradient they defer and then project on the device."""

    def map2pro(dSPECIN, NWAVE, NVMR, NDUST, NPRO, NPATH, NLAYIN, LAYINC, DTE, DAM, DCO, INCPAR=[-1]):
        if isinstance(dSPECIN, DeviceGradient):
            return ProfileGradient(dSPECIN, NVMR, NDUST, NPRO, NPATH, NLAYIN, LAYINC, DTE, DAM, DCO, INCPAR, ref_map2pro)
        return ref_map2pro(dSPECIN, NWAVE, NVMR, NDUST, NPRO, NPATH, NLAYIN, LAYINC, DTE, DAM, DCO, INCPAR=INCPAR)

    def map2xvec(dSPECIN, NWAVE, NVMR, NDUST, NPRO, NPATH, NX, xmap):
        if isinstance(dSPECIN, ProfileGradient):
            pg = dSPECIN
            nvmr, ndust, npro, npath, NLAYIN, LAYINC, DTE, DAM, DCO = pg.args
            inc = list(pg.INCPAR)
            # the fused matrix applies the reference's own inclusion rule (parameters whose xmap mean is non-zero,
            # :699-702); any other INCPAR goes through the reference functions on the materialised array
            if inc == _plan.included_params(xmap) and pg._host is None and pg.grad._host is None:
                M = _plan.fold_projection(xmap, LAYINC, NLAYIN, DTE, DAM, DCO, nvmr, ndust)
                hp = pg.grad._hp
                return hp.to_host(hp.project(pg.grad._dev, M))
            dSPECIN = pg.materialise()
        return ref_map2xvec(np.asarray(dSPECIN), NWAVE, NVMR, NDUST, NPRO, NPATH, NX, xmap)

    map2pro.__doc__ = (ref_map2pro.__doc__ or "") + "\n(archnemesis_dist_b200: defers for device-resident gradients)"
    map2xvec.__doc__ = (ref_map2xvec.__doc__ or "") + "\n(archnemesis_dist_b200: fused projection on the device)"
    return map2pro, map2xvec


def _table_on_device(sp):
    """k-tables [NWAVE,NG,NP,NT,NGAS] and in-memory line-by-line tables [NWAVE,NP,NT,NGAS] run on the device;
    run-time line-by-line and tables read on line from HDF5 (K is None) stay on the reference."""
    if sp.NGAS < 1 or sp.K is None:
        return False
    ilbl = int(sp.ILBL)
    # shapes beyond the native limits (include/ansb200.h: NG*NG sort keys per warp, gradient columns, gas sum of the
    # LBL kernel) stay with the reference instead of raising from inside CIRSrad
    if ilbl == _K_TABLES and np.ndim(sp.K) == 5:
        return int(sp.NG) <= _lib.MAX_NG and int(sp.NGAS) <= _lib.MAX_NGAS
    if ilbl == _LBL_TABLES and np.ndim(sp.K) == 4:
        return int(sp.NGAS) <= _lib.MAX_LBL_NGAS
    return False


class StateBatch:
    """Rendezvous of the forward models of a numerical Jacobian (jacobian_nemesis, ForwardModel_0.py:2184-2361).

    The reference runs one forward model per perturbed state-vector element in joblib worker processes.  Here every
    forward model runs the reference's own driver (nemesisfm / nemesisSOfm / ...) in a THREAD of this process; its
    CIRSrad call hands the host-side inputs of the radiative transfer to this object and waits.  When every live
    thread waits, all pending evaluations are merged -- the states become one more dimension of the launch
    (engine.combine_evaluations) -- and run in one gas-opacity + one radiance launch per group of compatible
    evaluations; the threads then carry on with their own results.  Forward models with several CIRSrad calls
    (geometries, averaging points) simply meet again."""

    def __init__(self, nthreads):
        import threading
        self.cv = threading.Condition()
        self.active = nthreads
        self.pending = []
        self.results = {}
        self.error = None
        self.rounds = 0
        self.launch_groups = 0
        self.evaluations = 0

    def submit(self, hp, ev):
        """Called from CIRSrad(return_grad=False) of a worker thread: returns SPECOUT[NWAVE,NPATH] (numpy)."""
        with self.cv:
            ticket = object()
            self.pending.append((ticket, hp, ev))
            if len(self.pending) >= self.active:
                self._run_locked()
            else:
                while ticket not in self.results and self.error is None:
                    self.cv.wait()
            if ticket not in self.results:
                raise RuntimeError("batched forward models: a companion evaluation failed") from self.error
            return self.results.pop(ticket)

    def done(self):
        """A worker thread has finished (or died): it will not submit again."""
        with self.cv:
            self.active -= 1
            if self.pending and len(self.pending) >= self.active:
                self._run_locked()

    def _run_locked(self):
        pending, self.pending = self.pending, []
        try:
            groups = {}
            for ticket, hp, ev in pending:
                groups.setdefault((id(hp), _engine.batch_key(ev)), []).append((ticket, hp, ev))
            for items in groups.values():
                hp = items[0][1]
                if len(items) == 1:
                    out = hp.to_host(hp.cirsrad(items[0][2], False))
                    self.results[items[0][0]] = np.array(out)
                else:
                    ev, cols = _engine.combine_evaluations([it[2] for it in items])
                    out = hp.to_host(hp.cirsrad(ev, False))
                    for (ticket, _, _), (c0, n) in zip(items, cols):
                        self.results[ticket] = np.array(out[:, c0:c0 + n])
                self.launch_groups += 1
                self.evaluations += len(items)
            self.rounds += 1
        except BaseException as e:       # noqa: BLE001  (handed to every waiting thread)
            self.error = e
        self.cv.notify_all()


class B200HotPathMixin:
    """Overrides of the hot methods of ForwardModel_0.  `b200_engine` may be replaced (tests inject an
    oracle-backed engine to exercise the host logic without a GPU)."""

    b200_engine = _engine

    # -- scope test -----------------------------------------------------------------------------
    def _b200_mode(self):
        """THERMAL / TRANSMISSION if this path type runs on the device, else None."""
        sp = self.SpectroscopyX
        if not _table_on_device(sp):
            return None
        if getattr(self, "EmissionsX", None) is not None:
            return None
        imod = np.unique(self.PathX.IMOD)
        if imod.size != 1:
            return None
        imod = int(imod[0])
        if not (imod & _NON_TRANSMISSION_BITS):
            return _engine.TRANSMISSION
        if not (imod & _ABSORBTION) and (imod & _THERMAL_EMISSION):     # dispatch order of CIRSrad :4478-4489
            return _engine.THERMAL
        return None

    # -- host-side inputs -----------------------------------------------------------------------
    def _b200_hotpath(self):
        sp = self.SpectroscopyX
        eng = self.b200_engine
        return _TABLES.get(sp, lambda: eng.HotPath(sp.K, sp.PRESS, sp.TEMP, sp.DELG, sp.WAVE))

    def _b200_continuum(self, return_grad):
        """TAUCIA, TAUDUST, TAURAY [NWAVE,NLAY] and dTAUCON [NWAVE,NPAR,NLAY] with the reference's own
        host routines and the assembly of calculate_layer_opacity (ForwardModel_0.py:3938-3981),
        including its quirks."""
        NWAVE, NLAY = self.SpectroscopyX.NWAVE, self.LayerX.NLAY
        NVMR, NDUST = self.AtmosphereX.NVMR, self.ScatterX.NDUST
        dTAUCON = np.zeros((NWAVE, NVMR + 2 + NDUST, NLAY)) if return_grad else None
        touched = False      # no contribution at all: hand None to the engine instead of NWAVE*NPAR*NLAY zeros
        TAUCIA, dTAUCIA = self.calculate_vertical_cia_opacity(return_grad)
        if return_grad and dTAUCIA is not None:
            touched = True
            dTAUCON[:, 0:NVMR, :] = dTAUCON[:, 0:NVMR, :] + np.transpose(
                np.transpose(dTAUCIA[:, :, 0:NVMR], axes=(2, 0, 1)) / (self.LayerX.TOTAM.T), axes=(1, 0, 2))
            dTAUCON[:, NVMR, :] = dTAUCON[:, NVMR, :] + dTAUCIA[:, :, NVMR]
        TAURAY, dTAURAY = self.calc_tau_rayleigh(MakePlot=False)
        self.LayerX.TAURAY = TAURAY
        if return_grad and (dTAURAY is not None):
            touched = touched or NVMR > 0
            for i in range(NVMR):
                dTAUCON[:, i, :] = dTAUCON[:, i, :] + dTAURAY[:, :]
        TAUDUST1, TAUCLSCAT, dTAUDUST1, dTAUCLSCAT = self.calc_tau_dust()
        TAUDUST1 = np.clip(np.nan_to_num(TAUDUST1), 0, 1e20)
        TAUDUST = np.sum(TAUDUST1, 2)
        self.LayerX.TAUDUST = TAUDUST
        self.LayerX.TAUSCAT = np.sum(TAUCLSCAT, 2)
        self.LayerX.TAUCLSCAT = TAUCLSCAT
        if return_grad:
            touched = touched or NDUST > 0
            for i in range(NDUST):
                dTAUCON[:, NVMR + 1 + i, :] = dTAUCON[:, NVMR + 1 + i, :] + dTAUDUST1[:, :, i]
        if hasattr(self, "CIAX") and self.CIAX is None:
            TAUCIA = None                 # calculate_vertical_cia_opacity returned zeros (:3894-3896)
        return TAUCIA, TAUDUST, TAURAY, (dTAUCON if touched else None)

    # -- continuum terms as a device plan (continuum.py) ------------------------------------------------
    b200_device_continuum = True       # False: always hand the dense host arrays of _b200_continuum to the engine

    def _b200_rayleigh_factors(self):
        """(ur[NR,NWAVE], vr[NR,NLAY], vrd[NR,NLAY]) with TAURAY = sum ur*vr, dTAURAY = sum ur*vrd, or None when this
        Rayleigh mode is not a product of a spectrum and a layer quantity (JOVIAN_AIR: per-layer composition)."""
        fm = sys.modules["archnemesis.ForwardModel_0"]
        iray = int(self.ScatterX.IRAY)
        WAVEC = self.SpectroscopyX.WAVE
        TOTAM = np.asarray(self.LayerX.TOTAM, dtype=np.float64)
        if iray == 0:
            return np.zeros((0, len(WAVEC))), np.zeros((0, len(TOTAM))), np.zeros((0, len(TOTAM)))
        ispace = fm.WaveUnitEnum(self.MeasurementX.ISPACE)
        one = np.ones(1)
        if iray == int(fm.RayleighScatteringModeEnum.GAS_GIANT_ATM):
            t, d = fm.calc_tau_rayleighj(ispace, WAVEC, one)       # cross section * 1.0: the spectrum itself (:5586-5594)
        elif iray == int(fm.RayleighScatteringModeEnum.C02_DOMINATED_ATM):
            t, d = fm.calc_tau_rayleighv2(ispace, WAVEC, one)
        else:
            return None
        if not np.array_equal(t[:, 0], d[:, 0]):
            return None
        return t[:, 0][None, :].copy(), TOTAM[None, :].copy(), np.ones((1, len(TOTAM)))

    def _b200_dust_spectra(self):
        """ud[NDUST,NWAVE] = kext * 1e-4 exactly as calc_tau_dust interpolates it (:4832-4863): the reference routine
        itself on a one-layer column of unit density.  Applies the same renormalisation of LayerX.CONT (:4834-4835)."""
        nd = int(self.ScatterX.NDUST)
        NW = self.SpectroscopyX.NWAVE
        if nd == 0:
            return np.zeros((0, NW))
        lay = self.LayerX
        for i in range(nd):
            if i in self.AtmosphereX.DUST_RENORMALISATION.keys():
                lay.CONT[:, i] = lay.CONT[:, i] / lay.CONT[:, i].sum() * 1e4 * self.AtmosphereX.DUST_RENORMALISATION[i]
        unit = types.SimpleNamespace(NLAY=1, CONT=np.ones((1, nd)))
        saved = self.AtmosphereX.DUST_RENORMALISATION
        self.AtmosphereX.DUST_RENORMALISATION = {}
        try:
            _, TAUCLSCAT, dT, dS = self.calc_tau_dust(Layer=unit)
        finally:
            self.AtmosphereX.DUST_RENORMALISATION = saved
        # (side products the reference keeps on LayerX: scattering opacities, as _b200_continuum sets them)
        ksca = np.ascontiguousarray(dS[:, 0, :].T)
        self.LayerX.TAUCLSCAT = ksca.T[:, None, :] * np.asarray(lay.CONT)[None, :, :nd]
        self.LayerX.TAUSCAT = np.sum(self.LayerX.TAUCLSCAT, 2)
        return np.ascontiguousarray(dT[:, 0, :].T)

    def _b200_continuum_plan(self, hp):
        """(tables, plan) for ansb200_continuum, or None when some term is outside what the plan expresses (then the
        dense arrays of _b200_continuum are sent instead)."""
        from . import continuum as _cont
        if not self.b200_device_continuum or not hasattr(hp, "continuum_tables"):
            return None
        atm, lay = self.AtmosphereX, self.LayerX
        if atm.NVMR + 2 > _cont.MAX_SLOTS:
            return None
        ray = self._b200_rayleigh_factors()
        if ray is None:
            return None
        cia = getattr(self, "CIAX", None)
        fm = sys.modules["archnemesis.ForwardModel_0"]
        WAVEC = self.SpectroscopyX.WAVE

        def make_tables():
            cia_mod = sys.modules.get("archnemesis.CIA_0")
            if cia_mod is None:
                import importlib
                cia_mod = importlib.import_module("archnemesis.CIA_0")
            gas_enum = sys.modules["archnemesis"].enum.GasEnum
            return _cont.build_tables(int(self.MeasurementX.ISPACE), WAVEC, cia, atm,
                                      (cia_mod.co2cia, cia_mod.n2n2cia, cia_mod.n2h2cia), gas_enum)
        if cia is None:
            key = ("nocia", len(WAVEC))
        else:
            kc = np.ascontiguousarray(cia.K_CIA)
            key = (hash(kc.tobytes()), kc.shape, hash(np.ascontiguousarray(cia.WAVEN).tobytes()),
                   hash(np.ascontiguousarray(WAVEC).tobytes()), tuple(int(x) for x in atm.ID), tuple(int(x) for x in atm.ISO),
                   int(cia.INORMAL), tuple(int(x) for x in cia.INORMALT), tuple(int(x) for x in cia.IPAIRG1),
                   tuple(int(x) for x in cia.IPAIRG2), int(self.MeasurementX.ISPACE))
        tables = hp.continuum_tables(key, make_tables)
        ud = self._b200_dust_spectra()
        plan = _cont.build_plan(tables, cia, atm, lay, int(self.ScatterX.NDUST), ray, ud, _SQ_CM_TO_SQ_METER)
        if plan is None:
            return None
        del fm
        return tables, plan

    def _b200_surface_terms(self, mode):
        """xfac, EMISSIVITY, SOLFLUX, REFLECTANCE as calculate_thermal_emission_spectrum
        (ForwardModel_0.py:4157-4213) / calculate_transmission_spectrum (:4113-4119) prepare them."""
        import scipy.interpolate
        WAVE = self.SpectroscopyX.WAVE
        NWAVE = self.SpectroscopyX.NWAVE
        xfac = np.ones(NWAVE)
        if mode == _engine.TRANSMISSION:
            if int(self.MeasurementX.IFORM) == _IFORM_ATM_TRANSMISSION:
                self.StellarX.calc_solar_flux()
                xfac = scipy.interpolate.interp1d(self.StellarX.WAVE, self.StellarX.SOLFLUX)(WAVE)
            return xfac, None, None, None
        if int(self.MeasurementX.IFORM) == _IFORM_FLUXRATIO:
            xfac *= np.pi * 4. * np.pi * ((self.AtmosphereX.RADIUS) * 1.0e2) ** 2.
            self.StellarX.calc_solar_flux()
            solflux = scipy.interpolate.interp1d(self.StellarX.WAVE, self.StellarX.SOLFLUX)(WAVE)
            xfac = xfac / solflux
        if self.SurfaceX.TSURF > 0.0:
            EMISSIVITY = scipy.interpolate.interp1d(self.SurfaceX.VEM, self.SurfaceX.EMISSIVITY)(WAVE)
        else:
            EMISSIVITY = np.zeros(NWAVE)
        SOLFLUX = np.zeros(NWAVE)
        REFLECTANCE = np.zeros(NWAVE)
        st = self.StellarX
        if (st is not None and st.SOLEXIST is True and self.SurfaceX.GASGIANT is False
                and int(self.SurfaceX.LOWBC) != 0):      # LowerBoundaryConditionEnum.THERMAL == 0
            st.calc_solar_flux()
            SOLFLUX = scipy.interpolate.interp1d(st.WAVE, st.SOLFLUX)(WAVE)
            # the reference leaves REFLECTANCE at zero (the BRDF line is commented out, :4208-4209)
        return xfac, EMISSIVITY, SOLFLUX, REFLECTANCE

    def _b200_evaluation(self, mode, return_grad):
        sp, lay, path, atm = self.SpectroscopyX, self.LayerX, self.PathX, self.AtmosphereX
        gas_slot = np.array([atm.locate_gas(sp.ID[i], sp.ISO[i]) for i in range(sp.NGAS)], dtype=np.int32)
        amount = np.zeros((sp.NGAS, lay.NLAY))
        for i in range(sp.NGAS):
            amount[i, :] = lay.AMOUNT[:, gas_slot[i]] * _SQ_CM_TO_SQ_METER      # :3861
        cplan = self._b200_continuum_plan(self._b200_hotpath())
        if cplan is None:
            TAUCIA, TAUDUST, TAURAY, dTAUCON = self._b200_continuum(return_grad)
        else:
            TAUCIA = TAUDUST = TAURAY = dTAUCON = None       # made on the device from the plan
        xfac, EMISSIVITY, SOLFLUX, REFLECTANCE = self._b200_surface_terms(mode)
        NPAR = atm.NVMR + 2 + self.ScatterX.NDUST
        return _engine.Evaluation(
            press_atm=lay.PRESS / _ATM_TO_PASCAL, temp=lay.TEMP, amount=amount, gas_slot=gas_slot, NVMR=atm.NVMR,
            NPAR=NPAR, LAYINC=path.LAYINC, SCALE=path.SCALE, NLAYIN=path.NLAYIN, EMTEMP=path.EMTEMP,
            LAYPRESS=lay.PRESS, taucia=TAUCIA, taudust=TAUDUST, tauray=TAURAY, dtaucon=dTAUCON, mode=mode,
            ISPACE=int(self.MeasurementX.ISPACE), TSURF=float(self.SurfaceX.TSURF), EMISSIVITY=EMISSIVITY, xfac=xfac,
            SOLFLUX=SOLFLUX, REFLECTANCE=REFLECTANCE,
            SOL_ANG=None if return_grad else np.asarray(path.SOL_ANG, dtype=np.float64),
            EMISS_ANG=None if return_grad else np.asarray(path.EMISS_ANG, dtype=np.float64), continuum=cplan)

    # -- overridden reference methods ------------------------------------------------------------
    def calculate_gaseous_line_opacity(self, return_grad=False):
        """K_TABLES (:3850-3877) and LINE_BY_LINE_TABLES (:3795-3815) branches of
        ForwardModel_0.calculate_gaseous_line_opacity on the device.
        Returns numpy TAUGAS[NWAVE,NG,NLAY] and dTAUGAS[NWAVE,NG,NPAR,NLAY] like the reference."""
        sp = self.SpectroscopyX
        if not _table_on_device(sp):
            return super().calculate_gaseous_line_opacity(return_grad)
        atm, lay = self.AtmosphereX, self.LayerX
        hp = self._b200_hotpath()
        gas_slot = [atm.locate_gas(sp.ID[i], sp.ISO[i]) for i in range(sp.NGAS)]
        amount = np.stack([lay.AMOUNT[:, g] * _SQ_CM_TO_SQ_METER for g in gas_slot])
        ev = _engine.Evaluation(press_atm=lay.PRESS / _ATM_TO_PASCAL, temp=lay.TEMP, amount=amount,
                                gas_slot=np.asarray(gas_slot, np.int32), NVMR=atm.NVMR, NPAR=0,
                                LAYINC=np.zeros((1, 1), np.int32), SCALE=np.zeros((1, 1)), NLAYIN=np.ones(1, np.int32))
        out = hp.gas_opacity(hp.stage(ev, return_grad))
        if not return_grad:
            return hp.to_host(out), None
        tau, dk = hp.to_host(out[0]), hp.to_host(out[1])
        dTAUGAS = np.zeros([sp.NWAVE, sp.NG, atm.NVMR + 2 + self.ScatterX.NDUST, lay.NLAY])
        for i, g in enumerate(gas_slot):
            dTAUGAS[:, :, g, :] = dk[:, :, :, i] * _SQ_CM_TO_SQ_METER
        dTAUGAS[:, :, atm.NVMR, :] = dk[:, :, :, sp.NGAS]
        return tau, dTAUGAS

    def CIRSrad(self, return_grad=False):
        """ForwardModel_0.CIRSrad (:4376-4511) for thermal-emission and pure-transmission paths.
        Returns SPECOUT[NWAVE,NPATH] or (SPECOUT, dSPECOUT[NWAVE,NPAR,NLAYIN,NPATH], dTSURF[NWAVE,NPATH])."""
        mode = self._b200_mode()
        if mode is None:
            return super().CIRSrad(return_grad)
        hp = self._b200_hotpath()
        ev = self._b200_evaluation(mode, return_grad)
        batch = getattr(self, "_b200_batch", None)
        if batch is not None and not return_grad:
            return batch.submit(hp, ev)           # numerical Jacobian: evaluated together with the other states
        out = hp.cirsrad(ev, return_grad)
        if not return_grad:
            return hp.to_host(out)
        spec, dspec, dtsurf = out
        if _INSTALLED.get("lazy_gradients") and hasattr(hp, "project"):
            # (NWAVE, NPAR, NLAYIN, NPATH) like the reference, but still on the device: see DeviceGradient
            nw, npath, npar, nlm = dspec.shape
            return hp.to_host(spec), DeviceGradient(hp, dspec, (nw, npar, nlm, npath)), hp.to_host(dtsurf)
        return hp.to_host(spec), np.ascontiguousarray(np.transpose(hp.to_host(dspec), (0, 2, 3, 1))), hp.to_host(dtsurf)

    def b200_forward_jacobian(self, xmap):
        """Fused replacement of the CIRSrad -> map2pro -> map2xvec triple of nemesisfmg (:694-718):
        returns SPEC1[NWAVE,NPATH], dSPEC1[NWAVE,NPATH,NX] with the JSURF column filled."""
        mode = self._b200_mode()
        atm, lay, path = self.AtmosphereX, self.LayerX, self.PathX
        if mode is None:
            ref = sys.modules["archnemesis.ForwardModel_0"]   # out-of-scope path types stay on the reference
            SPEC1, dSPEC3, dTSURF = super().CIRSrad(return_grad=True)
            incpar = _plan.included_params(xmap)
            if len(incpar) > 0:
                dSPEC2 = ref.map2pro(dSPEC3, self.SpectroscopyX.NWAVE, atm.NVMR, atm.NDUST, atm.NP, path.NPATH,
                                     path.NLAYIN, path.LAYINC, lay.DTE, lay.DAM, lay.DCO, INCPAR=incpar)
            else:
                dSPEC2 = np.zeros((self.SpectroscopyX.NWAVE, atm.NVMR + 2 + atm.NDUST, atm.NP, path.NPATH))
            dSPEC1 = ref.map2xvec(dSPEC2, self.SpectroscopyX.NWAVE, atm.NVMR, atm.NDUST, atm.NP, path.NPATH,
                                  self.Variables.NX, xmap)
        else:
            hp = self._b200_hotpath()
            ev = self._b200_evaluation(mode, True)
            M = _plan.fold_projection(xmap, path.LAYINC, path.NLAYIN, lay.DTE, lay.DAM, lay.DCO, atm.NVMR, atm.NDUST)
            if int(path.NPATH) >= 4 and getattr(hp, "layer_space", False):
                # many paths (limb / occultation geometries evaluated together): gradients per layer on the device,
                # one projection matrix for all paths
                Mlay = _plan.fold_projection_layers(xmap, lay.NLAY, lay.DTE, lay.DAM, lay.DCO, atm.NVMR, atm.NDUST)
                spec, dx, dtsurf = hp.forward_jacobian(ev, M, Mlay=Mlay)
            else:
                spec, dx, dtsurf = hp.forward_jacobian(ev, M)
            SPEC1, dSPEC1, dTSURF = hp.to_host(spec), hp.to_host(dx), hp.to_host(dtsurf)
        if self.Variables.JSURF >= 0:
            dSPEC1[:, 0, self.Variables.JSURF] = dTSURF[:, 0]      # :717-718
        return SPEC1, dSPEC1

    def b200_device_conv_ok(self, IGEOM):
        """Can the tail of nemesisfmg for this geometry (JSURF column, WGEOM weight, Measurement_0.convg or
        integrate_filterg) run on the device?  One averaging point, no telluric transmission, k-tables with
        FWHM <= 0 (the only modes the reference's convg supports with k-tables; FWHM < 0 for integrated
        radiances), and a path type on the device."""
        M = self.Measurement
        if not (self._b200_mode() is not None and int(M.NAV[IGEOM]) == 1 and self.Telluric is None and
                self.PathX.NPATH == 1):
            return False
        if int(M.NCONV[IGEOM]) > _lib.MAX_NCONV:       # grid limit of ans_convolve_kernel: convolve on the host
            return False
        if int(M.IFORM) == _IFORM_INTEGRATED_RADIANCE:
            return float(M.FWHM) < 0.0            # integrate_filterg raises for FWHM >= 0 (Measurement_0.py:2772)
        if int(self.Spectroscopy.ILBL) == _LBL_TABLES:
            return True                           # lblconvg: analytic shapes (FWHM > 0), filters (< 0), np.interp (== 0)
        return float(M.FWHM) <= 0.0

    def b200_forward_jacobian_conv(self, xmap, IGEOM, wgeom):
        """b200_forward_jacobian + the JSURF column + WGEOM + Measurement_0.convg (:716-768) with only
        [NCONV, 1+NX] coming back from the device.  Returns SPECONV1[NCONV], dSPECONV1[NCONV,NX]."""
        atm, lay, path, M = self.AtmosphereX, self.LayerX, self.PathX, self.Measurement
        hp = self._b200_hotpath()
        ev = self._b200_evaluation(self._b200_mode(), True)
        Mx = _plan.fold_projection(xmap, path.LAYINC, path.NLAYIN, lay.DTE, lay.DAM, lay.DCO, atm.NVMR, atm.NDUST)
        cop = self._b200_conv_operator(hp, IGEOM)
        out = hp.to_host(hp.forward_jacobian_conv(ev, Mx, cop, int(self.Variables.JSURF), float(wgeom)))
        return out[:, 0], out[:, 1:]

    def _b200_conv_operator(self, hp, IGEOM, integrated=None):
        """The line-shape operator of geometry IGEOM on the device, rebuilt only when the calculation grid or the
        measurement's line-shape description changes (the analytic shapes cost NCONV x window-length libm calls).
        integrated: filter integral instead of a line shape (default: IFORM says so; the occultation driver never does)."""
        M, WAVE = self.Measurement, np.asarray(self.SpectroscopyX.WAVE, dtype=np.float64)
        n = int(M.NCONV[IGEOM])
        integ = int(M.IFORM) == _IFORM_INTEGRATED_RADIANCE if integrated is None else bool(integrated)
        lbl = int(self.Spectroscopy.ILBL) == _LBL_TABLES
        vconv = np.asarray(M.VCONV[0:n, IGEOM], dtype=np.float64)
        key = [IGEOM, integ, lbl, float(M.FWHM), int(getattr(M, "ISHAPE", 0) or 0), float(getattr(M, "V_DOPPLER", 0.0) or 0.0),
               WAVE.tobytes(), vconv.tobytes()]
        if float(M.FWHM) < 0.0:
            key += [np.asarray(M.NFIL).tobytes(), np.asarray(M.VFIL).tobytes(), np.asarray(M.AFIL).tobytes()]
        key = hash(tuple(key))
        cache = self.__dict__.setdefault("_b200_conv_cache", {})
        hit = cache.get(IGEOM)
        if hit is not None and hit[0] == key and hit[1] is hp:
            return hit[2]
        if integ:
            # integrate_filterg works on the Doppler-corrected grid (Measurement_0.py:2768)
            op = _plan.filter_integral_operator(M.correct_doppler_shift(WAVE), n, M.NFIL, M.VFIL, M.AFIL)
        elif lbl:
            # lblconvg, also on the Doppler-corrected grid (Measurement_0.py:2222)
            op = _plan.lbl_conv_operator(M.correct_doppler_shift(WAVE), vconv, float(M.FWHM), int(M.ISHAPE), M.NFIL, M.VFIL,
                                         M.AFIL, grad=True)
        else:
            op = _plan.conv_operator(WAVE, vconv, float(M.FWHM), M.NFIL, M.VFIL, M.AFIL)
        cop = hp.conv_operator(op)
        cache[IGEOM] = (key, hp, cop)
        return cop


    # ---- limb / occultation drivers: every tangent height in one evaluation ---------------------------------------
    def b200_tangent_conv_ok(self, integrated):
        """Can the tail of nemesisSOfmg / nemesisLfmg (tangent-height interpolation + line shape with IGEOM='All') run on
        the device?  The combinations the reference itself supports for IGEOM='All': k-tables with FWHM == 0
        (Measurement_0.convg :2506-2528 raises otherwise), line-by-line tables with an analytic shape (FWHM > 0, not
        Hamming: lblconvg_ngeom's Hamming window is empty, :3752-3754) or filter functions (FWHM < 0), and the filter
        integral of nemesisLfmg; every geometry on geometry 0's grid."""
        M = self.MeasurementX
        if self._b200_mode() is None or M.NORDERS_AOTF is not None:
            return False
        n0 = int(M.NCONV[0])
        if n0 > _lib.MAX_NCONV or not np.all(np.asarray(M.NCONV) == n0) or int(np.asarray(M.VCONV).shape[0]) < n0:
            return False
        fwhm = float(M.FWHM)
        if integrated:
            return fwhm < 0.0
        if int(self.SpectroscopyX.ILBL) == _LBL_TABLES:
            return fwhm < 0.0 or (fwhm > 0.0 and int(getattr(M, "ISHAPE", 0) or 0) != _plan.ILS_HAMMING)
        return fwhm == 0.0

    def b200_tangent_fmg(self, calc_path, filter_integral_allowed):
        """Body shared by the nemesisSOfmg / nemesisLfmg overrides (ForwardModel_0.py:1160-1243, :1409-1518; the AOTF
        branch of the occultation driver stays with the reference).  The set-up calls are the reference's; CIRSrad ->
        map2pro -> map2xvec -> tangent-height interpolation -> line shape run on the device and only
        SPECMOD[NWAVE,NGEOM] and [NCONV,NGEOM,1+NX] come back (the reference moves dSPECOUT[NWAVE,NPATH,NX]).
        Returns None when the case is not one the device tail covers (the caller then runs the reference's body)."""
        from copy import deepcopy
        self.Variables1 = deepcopy(self.Variables)
        self.MeasurementX = deepcopy(self.Measurement)
        self.AtmosphereX = deepcopy(self.Atmosphere)
        self.ScatterX = deepcopy(self.Scatter)
        self.StellarX = deepcopy(self.Stellar)
        self.SurfaceX = deepcopy(self.Surface)
        self.LayerX = deepcopy(self.Layer)
        self.SpectroscopyX = deepcopy(self.Spectroscopy)
        self.CIAX = deepcopy(self.CIA)
        self.check_gas_spec_atm()
        self.check_wave_range_consistency()
        if self.MeasurementX.NORDERS_AOTF is not None:
            return None
        self.Measurement.build_ils(IGEOM=0)
        wmin, wmax = self.Measurement.calc_wave_range(apply_doppler=True, IGEOM=None)
        if self.SpectroscopyX.NGAS > 0:
            self.SpectroscopyX.read_tables(wavemin=wmin, wavemax=wmax)
        self.adjust_hydrostat = False
        xmap = self.subprofretg()
        calc_path()
        MX, path, lay, atm = self.MeasurementX, self.PathX, self.LayerX, self.AtmosphereX
        integrated = bool(filter_integral_allowed and int(MX.IFORM) == _IFORM_INTEGRATED_RADIANCE)
        if not self.b200_tangent_conv_ok(integrated):
            return None
        BASEH_TANHE = np.zeros(path.NPATH)
        for i in range(path.NPATH):
            BASEH_TANHE[i] = lay.BASEH[path.LAYINC[int(path.NLAYIN[i] / 2), i]] / 1.0e3
        try:
            mix = _plan.tangent_mix(BASEH_TANHE, MX.TANHE[:, 0] if np.ndim(MX.TANHE) == 2 else MX.TANHE)
        except ValueError:
            return None
        hp = self._b200_hotpath()
        ev = self._b200_evaluation(self._b200_mode(), True)
        M = _plan.fold_projection(xmap, path.LAYINC, path.NLAYIN, lay.DTE, lay.DAM, lay.DCO, atm.NVMR, atm.NDUST)
        Mlay = None
        if int(path.NPATH) >= 4 and getattr(hp, "layer_space", False):
            Mlay = _plan.fold_projection_layers(xmap, lay.NLAY, lay.DTE, lay.DAM, lay.DCO, atm.NVMR, atm.NDUST)
        cop = self._b200_conv_operator(hp, 0, integrated=integrated)
        specmod, out = hp.forward_jacobian_mix_conv(ev, M, mix, cop, Mlay=Mlay)
        SPECMOD, out = hp.to_host(specmod), hp.to_host(out)
        SPECONV, dSPECONV = np.ascontiguousarray(out[:, :, 0]), np.ascontiguousarray(out[:, :, 1:])
        if not integrated:
            dSPECONV = self.subspeconv(self.SpectroscopyX.WAVE, SPECMOD, dSPECONV)
        return self.subspecret(SPECONV, dSPECONV)


class ArrayForwardModel(B200HotPathMixin):
    b200_device_continuum = False      # the continuum terms arrive as arrays, there is no reference object to plan from

    """The mix-in over plain namespaces: for hosts without the reference package.  `objects` maps
    the names SpectroscopyX, LayerX, PathX, AtmosphereX, SurfaceX, MeasurementX, ScatterX,
    StellarX, Variables to objects carrying the attributes listed in SURVEY.md 8b; the continuum
    terms the reference computes on the host are supplied as arrays."""

    def __init__(self, objects, TAUCIA=None, dTAUCIA=None, TAURAY=None, dTAURAY=None, TAUDUST1=None,
                 TAUCLSCAT=None, dTAUDUST1=None, dTAUCLSCAT=None):
        for k, v in objects.items():
            setattr(self, k, v)
        self.EmissionsX = objects.get("EmissionsX")
        self._cont = dict(TAUCIA=TAUCIA, dTAUCIA=dTAUCIA, TAURAY=TAURAY, dTAURAY=dTAURAY, TAUDUST1=TAUDUST1,
                          TAUCLSCAT=TAUCLSCAT, dTAUDUST1=dTAUDUST1, dTAUCLSCAT=dTAUCLSCAT)

    def _zeros(self, *extra):
        return np.zeros((self.SpectroscopyX.NWAVE, self.LayerX.NLAY) + extra)

    def calculate_vertical_cia_opacity(self, return_grad=False):
        c = self._cont
        return (c["TAUCIA"] if c["TAUCIA"] is not None else self._zeros()), (c["dTAUCIA"] if return_grad else None)

    def calc_tau_rayleigh(self, MakePlot=False):
        c = self._cont
        return (c["TAURAY"] if c["TAURAY"] is not None else self._zeros()), c["dTAURAY"]

    def calc_tau_dust(self):
        c = self._cont
        nd = self.ScatterX.NDUST
        z = self._zeros(nd)
        return (c["TAUDUST1"] if c["TAUDUST1"] is not None else z, c["TAUCLSCAT"] if c["TAUCLSCAT"] is not None else z,
                c["dTAUDUST1"] if c["dTAUDUST1"] is not None else z, c["dTAUCLSCAT"] if c["dTAUCLSCAT"] is not None else z)


def namespace(**kw):
    return types.SimpleNamespace(**kw)


# ------------------------------------------------------------------------------------------------
# drop-in installation
# ------------------------------------------------------------------------------------------------
_INSTALLED = {}


def make_forward_model_class(reference_cls):
    """ForwardModel_B200: the reference class with the hot methods re-routed, plus nemesisfmg with the
    fused Jacobian (the body mirrors archnemesis/ForwardModel_0.py:593-779 step for step)."""

    class ForwardModel_B200(B200HotPathMixin, reference_cls):
        __doc__ = (reference_cls.__doc__ or "") + "\n\n(B200 hot path: archnemesis_dist_b200)"

        def nemesisfmg(self):
            from copy import deepcopy
            import os
            M, S = self.Measurement, self.Spectroscopy
            if self.Atmosphere.NLOCATIONS > 1 or self.Surface.NLOCATIONS > 1:
                raise ValueError('error in nemesisfm :: archNEMESIS has not been setup for dealing with multiple '
                                 'locations yet')
            self.check_gas_spec_atm()
            self.check_wave_range_consistency()
            SPECONV = np.zeros(M.MEAS.shape)
            dSPECONV = np.zeros((M.NCONV.max(), M.NGEOM, self.Variables.NX))
            for IGEOM in range(M.NGEOM):
                M.build_ils(IGEOM=IGEOM)
                wmin, wmax = M.calc_wave_range(apply_doppler=True, IGEOM=IGEOM)
                self.SpectroscopyX = deepcopy(S)
                if self.SpectroscopyX.NGAS > 0:
                    self.SpectroscopyX.read_tables(wavemin=wmin, wavemax=wmax)
                NW = self.SpectroscopyX.NWAVE
                SPEC = np.zeros(NW)
                dSPEC = np.zeros((NW, self.Variables.NX))
                conv_done = False
                for IAV in range(M.NAV[IGEOM]):
                    self.select_Measurement(IGEOM, IAV)
                    self.AtmosphereX = deepcopy(self.Atmosphere)
                    self.ScatterX = deepcopy(self.Scatter)
                    self.StellarX = deepcopy(self.Stellar)
                    self.SurfaceX = deepcopy(self.Surface)
                    self.LayerX = deepcopy(self.Layer)
                    self.CIAX = deepcopy(self.CIA)
                    self.TelluricX = deepcopy(self.Telluric)
                    MX = self.MeasurementX
                    if MX.EMISS_ANG[0, 0] >= 0.0:
                        self.ScatterX.SOL_ANG = MX.SOL_ANG[0, 0]
                        self.ScatterX.EMISS_ANG = MX.EMISS_ANG[0, 0]
                        self.ScatterX.AZI_ANG = MX.AZI_ANG[0, 0]
                    else:
                        self.ScatterX.SOL_ANG = MX.TANHE[0, 0]
                        self.ScatterX.EMISS_ANG = MX.EMISS_ANG[0, 0]
                    xmap = self.subprofretg()
                    self.LayerX.DUST_UNITS_FLAG = self.AtmosphereX.DUST_UNITS_FLAG
                    self.calc_pathg()
                    if self.b200_device_conv_ok(IGEOM):
                        # spectrum, Jacobian and the instrument line shape on the device: [NCONV, 1+NX] comes back
                        S1, dS1 = self.b200_forward_jacobian_conv(xmap, IGEOM, M.WGEOM[IGEOM, IAV])
                        conv_done = True
                        break
                    SPEC1, dSPEC1 = self.b200_forward_jacobian(xmap)
                    if M.NAV[IGEOM] >= 1:
                        SPEC[:] = SPEC[:] + M.WGEOM[IGEOM, IAV] * SPEC1[:, 0]
                        dSPEC[:, :] = dSPEC[:, :] + M.WGEOM[IGEOM, IAV] * dSPEC1[:, 0, :]
                    else:
                        SPEC[:] = SPEC1[:, 0]
                        dSPEC[:, :] = dSPEC1[:, 0, :]
                if conv_done:
                    n = M.NCONV[IGEOM]
                    SPECONV[0:n, IGEOM] = S1[0:n]
                    dSPECONV[0:n, IGEOM, :] = dS1[0:n, :]
                    continue
                if self.TelluricX is not None:
                    tmin, tmax = M.calc_wave_range(apply_doppler=False, IGEOM=IGEOM)
                    self.TelluricX.Spectroscopy.read_tables(wavemin=tmin, wavemax=tmax)
                    WT, TT = self.TelluricX.calc_transmission()
                    wavecorr = self.MeasurementX.correct_doppler_shift(self.SpectroscopyX.WAVE)
                    T = np.interp(wavecorr, WT, TT)
                    SPEC *= T
                    dSPEC[:, :] = (dSPEC[:, :].T * T).T
                if int(M.IFORM) == _IFORM_INTEGRATED_RADIANCE:
                    S1, dS1 = M.integrate_filterg(self.SpectroscopyX.WAVE, SPEC, dSPEC, IGEOM=IGEOM)
                elif int(S.ILBL) == _K_TABLES:
                    fwh = self.runname if os.path.exists(self.runname + '.fwh') else ''
                    S1, dS1 = M.convg(self.SpectroscopyX.WAVE, SPEC, dSPEC, IGEOM=IGEOM, FWHMEXIST=fwh)
                else:
                    S1, dS1 = M.lblconvg(self.SpectroscopyX.WAVE, SPEC, dSPEC, IGEOM=IGEOM)
                n = M.NCONV[IGEOM]
                SPECONV[0:n, IGEOM] = S1[0:n]
                dSPECONV[0:n, IGEOM, :] = dS1[0:n, :]
            return self.subspecret(SPECONV, dSPECONV)

        def nemesisSOfmg(self):
            out = self.b200_tangent_fmg(self.calc_pathg_SO, filter_integral_allowed=False)
            return out if out is not None else super().nemesisSOfmg()

        def nemesisLfmg(self):
            out = self.b200_tangent_fmg(self.calc_pathg_L, filter_integral_allowed=True)
            return out if out is not None else super().nemesisLfmg()

        b200_max_states = 64          # forward models per rendezvous (threads alive at once)

        def _b200_batchable(self):
            """Do the forward models of this retrieval reach the device (k-tables or line-by-line tables, no
            scattering)?  Otherwise the reference's joblib workers are the better tool."""
            try:
                return int(self.Spectroscopy.ILBL) in (_K_TABLES, _LBL_TABLES) and int(self.Scatter.ISCAT) == 0 and \
                    self.Spectroscopy.NGAS >= 1
            except Exception:      # noqa: BLE001
                return False

        def jacobian_nemesis(self, NCores=1, nemesisSO=False, nemesisL=False, nemesisC=False, nemesisdisc=False,
                             nemesisPT=False, analytical_gradient=True):
            """ForwardModel_0.jacobian_nemesis (archnemesis/ForwardModel_0.py:2184-2361) with the forward models of
            the numerical columns evaluated together on the device instead of in `NCores` joblib processes (NCores is
            ignored on that route).  Same outputs: YN[NY], KK[NY,NX].  (Every forward model works on its own copy of the
            state vector, like the reference's worker processes; with NCores = 1 the reference's joblib runs in-process
            and leaves Variables.XN at the last perturbed state, which rescales its last numerical column -- not
            reproduced.)"""
            if not self._b200_batchable():
                return super().jacobian_nemesis(NCores=NCores, nemesisSO=nemesisSO, nemesisL=nemesisL, nemesisC=nemesisC,
                                                nemesisdisc=nemesisdisc, nemesisPT=nemesisPT,
                                                analytical_gradient=analytical_gradient)
            import copy
            import threading
            V, Mm = self.Variables, self.Measurement
            V.calc_DSTEP()
            xnx = np.zeros([V.NX, V.NX + 1], dtype=float)
            xnx[:, 0] = V.XN
            xnx[:, 1:] = np.repeat(V.XN[:, None], V.NX, axis=1) + np.diag(V.DSTEP)
            zeros_mask = xnx[:, 1:] == 0
            xnx[:, 1:][zeros_mask] = 0.05
            if analytical_gradient is False:
                V.NUM[:] = 1
            ian1 = np.where(V.NUM == 0)[0]
            iYN = 0
            KK = np.zeros([Mm.NY, V.NX])
            if len(ian1) > 0:
                method = self.select_nemesis_fm(nemesisSO=nemesisSO, nemesisL=nemesisL, nemesisC=nemesisC,
                                                nemesisdisc=nemesisdisc, nemesisPT=nemesisPT, analytical_gradient=True)
                SPECMOD, dSPECMOD = method()
                YN = np.zeros(Mm.NY)
                ik = 0
                for igeom in range(Mm.NGEOM):
                    n = Mm.NCONV[igeom]
                    YN[ik:ik + n] = SPECMOD[0:n, igeom]
                    KK[ik:ik + n, :] = dSPECMOD[0:n, igeom, :]
                    ik += n
                iYN = 1
            inum = np.where((V.NUM == 1) & (V.FIX == 0))[0]
            if iYN == 0:
                nfm = len(inum) + 1
                ixrun = np.zeros(nfm, dtype='int32')
                ixrun[1:nfm] = inum[:] + 1
            else:
                nfm = len(inum)
                ixrun = np.zeros(nfm, dtype='int32')
                ixrun[0:nfm] = inum[:] + 1
            YNtot = np.zeros((Mm.NY, nfm))
            self.b200_batch_stats = dict(forward_models=nfm, rounds=0, launch_groups=0, evaluations=0)
            for lo in range(0, nfm, self.b200_max_states):
                hi = min(nfm, lo + self.b200_max_states)
                batch = StateBatch(hi - lo)
                cols, errors = {}, []

                def worker(ifm):
                    try:
                        # (what a joblib worker gets by pickling: its own forward model; the base objects are only
                        # read -- nemesisfm copies them into the *X attributes before changing anything)
                        fm = copy.copy(self)
                        fm.Variables = copy.deepcopy(self.Variables)
                        fm.Measurement = copy.deepcopy(self.Measurement)
                        fm._b200_batch = batch
                        out = fm.execute_fm((ifm, nfm, xnx, ixrun, nemesisSO, nemesisL, nemesisC, nemesisdisc, nemesisPT,
                                             np.zeros((Mm.NY, nfm)), 1))
                        cols[ifm] = out[:, ifm]
                    except BaseException as e:      # noqa: BLE001
                        errors.append(e)
                    finally:
                        batch.done()

                threads = [threading.Thread(target=worker, args=(ifm,), name="ansb200-fm-%d" % ifm) for ifm in range(lo, hi)]
                for t in threads:
                    t.start()
                for t in threads:
                    t.join()
                if errors:
                    raise errors[0]
                for ifm in range(lo, hi):
                    YNtot[:, ifm] = cols[ifm]
                for k in ("rounds", "launch_groups", "evaluations"):
                    self.b200_batch_stats[k] += getattr(batch, k)
            if iYN == 0 and nfm > 0:
                YN = np.zeros(Mm.NY)
                YN[:] = YNtot[0:Mm.NY, 0]
            for i in range(len(inum)):
                ifm = i + 1 if iYN == 0 else i
                xn1 = V.XN[inum[i]] * 1.05
                if xn1 == 0.0:
                    xn1 = 0.05
                if V.FIX[inum[i]] == 0:
                    KK[:, inum[i]] = (YNtot[:, ifm] - YN) / (xn1 - V.XN[inum[i]])
            return YN, KK

    ForwardModel_B200.__name__ = "ForwardModel_B200"
    return ForwardModel_B200


def install(archnemesis=None):
    """Rebind the names through which callers reach ForwardModel_0:
    ``archnemesis.ForwardModel_0`` (the class, star-imported at archnemesis/__init__.py:24),
    ``archnemesis.ForwardModel_0`` inside the submodule, and ``archnemesis.OptimalEstimation_0``'s
    module-level import (OptimalEstimation_0.py:24).  Returns the new class; ``uninstall()`` undoes it."""
    if archnemesis is None:
        import archnemesis  # noqa: F811
    mod = sys.modules["archnemesis.ForwardModel_0"]
    ref_cls = _INSTALLED.get("reference") or mod.ForwardModel_0
    cls = make_forward_model_class(ref_cls)
    _INSTALLED.update(reference=ref_cls, cls=cls)
    mod.ForwardModel_0 = cls
    archnemesis.ForwardModel_0 = cls
    oe = sys.modules.get("archnemesis.OptimalEstimation_0")
    if oe is not None and hasattr(oe, "ForwardModel_0"):
        oe.ForwardModel_0 = cls
    # the optimal-estimation algebra under coreretOE: coreretOE builds its solver object with
    # `from archnemesis import OptimalEstimation_0` (OptimalEstimation_0.py:1254, :1263), i.e. through the package
    # attribute, which -- like ForwardModel_0 -- is the class
    if oe is not None and hasattr(oe, "OptimalEstimation_0"):
        from . import oe as _oe_mod
        if "oe_reference" not in _INSTALLED:
            _INSTALLED["oe_reference"] = oe.OptimalEstimation_0
        oe_cls = _oe_mod.make_oe_class(_INSTALLED["oe_reference"])
        _INSTALLED["oe_cls"] = oe_cls
        oe.OptimalEstimation_0 = oe_cls
        archnemesis.OptimalEstimation_0 = oe_cls
    # the two module functions every driver calls by bare name after CIRSrad: with them
    # rebound, nemesisSOfmg / nemesisLfmg / process_IAV get the fused device projection unchanged
    if "map2pro" not in _INSTALLED:
        _INSTALLED.update(map2pro=mod.map2pro, map2xvec=mod.map2xvec)
    mod.map2pro, mod.map2xvec = make_fused_map_functions(_INSTALLED["map2pro"], _INSTALLED["map2xvec"])
    _INSTALLED["lazy_gradients"] = True
    # table residency across evaluations: read_tables memoised on the Spectroscopy class
    spec_mod = sys.modules.get("archnemesis.Spectroscopy_0")
    if spec_mod is not None and "read_tables" not in _INSTALLED:
        _INSTALLED["read_tables"] = spec_mod.Spectroscopy_0.read_tables
        spec_mod.Spectroscopy_0.read_tables = make_cached_read_tables(_INSTALLED["read_tables"])
    # run-time line-by-line cross sections (ILBL = 1: calc_klbl[g]_online, calc_lbltable): the pair loop of
    # add_line_set_monochromatic_absorption on the device, line lists resident (linedata.py)
    from . import linedata as _linedata
    _linedata.install_lbl()
    # vectorised .kta / .lta readers under read_tables (the first touch of a table set costs ~16 s of per-record
    # Python in the reference at the config-2 size, table_io.py)
    from . import table_io as _table_io
    _table_io.install_readers()
    # k-table generation (calc_ktable): the per-bin sort / quantile tail of calc_ktable_chunk on the device (ktable.py)
    from . import ktable as _ktable
    _ktable.install_ktable()
    return cls


def uninstall(archnemesis=None):
    if "reference" not in _INSTALLED:
        return
    if archnemesis is None:
        import archnemesis  # noqa: F811
    ref_cls = _INSTALLED.pop("reference")
    _INSTALLED.pop("cls", None)
    _INSTALLED.pop("lazy_gradients", None)
    if "map2pro" in _INSTALLED:
        sys.modules["archnemesis.ForwardModel_0"].map2pro = _INSTALLED.pop("map2pro")
        sys.modules["archnemesis.ForwardModel_0"].map2xvec = _INSTALLED.pop("map2xvec")
    if "read_tables" in _INSTALLED:
        sys.modules["archnemesis.Spectroscopy_0"].Spectroscopy_0.read_tables = _INSTALLED.pop("read_tables")
    from . import linedata as _linedata
    _linedata.uninstall_lbl()
    from . import table_io as _table_io
    _table_io.uninstall_readers()
    from . import ktable as _ktable
    _ktable.uninstall_ktable()
    if "oe_reference" in _INSTALLED:
        oe_ref = _INSTALLED.pop("oe_reference")
        _INSTALLED.pop("oe_cls", None)
        sys.modules["archnemesis.OptimalEstimation_0"].OptimalEstimation_0 = oe_ref
        archnemesis.OptimalEstimation_0 = oe_ref
    sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0 = ref_cls
    archnemesis.ForwardModel_0 = ref_cls
    oe = sys.modules.get("archnemesis.OptimalEstimation_0")
    if oe is not None and hasattr(oe, "ForwardModel_0"):
        oe.ForwardModel_0 = ref_cls
