"""The hot path as one object: resident k-table + per-evaluation inputs -> spectrum and Jacobian.

``HotPath`` is the public entry point of this package for callers that hold plain arrays (the
benchmark, the GPU tests, the ``ForwardModel_0`` mix-in of ``forward_model.py``).  It strings the
C-ABI calls of libansb200.so together on torch's current CUDA stream:

    gas_opacity (k-interp + random overlap)  ->  radiance (+ layer Jacobian)  ->  jacobian_project

Host arrays are staged through pinned memory; results come back as numpy arrays.  There is no CPU
path: constructing a ``HotPath`` without a CUDA device raises.
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import ops as _ops
from . import plan as _plan

THERMAL, TRANSMISSION = 0, 1
_CONT_LOCK = __import__("threading").Lock()
_PINNED_OUT = {}                  # (shape, dtype) -> pinned host buffer of HotPath.to_host
_PINNED_OUT_MAX = 64 << 20


@dataclass
class Evaluation:
    """Host-side inputs of one CIRSrad evaluation (one atmosphere state, NPATH paths).

    Names follow the reference objects they are read from:
      press_atm, temp          LayerX.PRESS/101325, LayerX.TEMP            [NLAY]
      amount                   LayerX.AMOUNT[:,IGAS]*1e-4 per active gas   [NGAS,NLAY]  (cm-2)
      gas_slot                 AtmosphereX.locate_gas(ID,ISO) per active gas [NGAS]
      taucia/taudust/tauray    continuum opacities                         [NWAVE,NLAY] or None
      dtaucon                  d(continuum)/d(parameter)                   [NWAVE,NPAR,NLAY] or None
      LAYINC, SCALE, EMTEMP    PathX                                       [NLAYMAX,NPATH]
      NLAYIN                   PathX.NLAYIN                                [NPATH]
      LAYPRESS                 LayerX.PRESS (Pa)                           [NLAY]
    """
    press_atm: np.ndarray
    temp: np.ndarray
    amount: np.ndarray
    gas_slot: np.ndarray
    NVMR: int
    NPAR: int
    LAYINC: np.ndarray
    SCALE: np.ndarray
    NLAYIN: np.ndarray
    EMTEMP: Optional[np.ndarray] = None
    LAYPRESS: Optional[np.ndarray] = None
    taucia: Optional[np.ndarray] = None
    taudust: Optional[np.ndarray] = None
    tauray: Optional[np.ndarray] = None
    dtaucon: Optional[np.ndarray] = None
    mode: int = THERMAL
    ISPACE: int = 0
    TSURF: float = -1.0
    EMISSIVITY: Optional[np.ndarray] = None
    xfac: Optional[np.ndarray] = None
    SOLFLUX: Optional[np.ndarray] = None
    REFLECTANCE: Optional[np.ndarray] = None
    SOL_ANG: Optional[np.ndarray] = None
    EMISS_ANG: Optional[np.ndarray] = None
    # (continuum.ContinuumTables, plan dict): taucia / taudust / tauray / dtaucon are then made on the device from it
    # (ansb200_continuum) and the four dense arrays above stay None
    continuum: Optional[tuple] = None
    h2d_bytes: int = field(default=0, init=False)


def batch_key(ev: Evaluation):
    """Evaluations with equal keys can be merged by combine_evaluations (everything that is per evaluation, not per
    layer or per path, in the C ABI: path type, surface temperature, parameter layout, per-wavenumber surface terms)."""
    def h(a):
        return None if a is None else hash(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    cont = None
    if ev.continuum is not None:
        tables, plan = ev.continuum
        cont = (id(tables), h(plan["ur"]), h(plan["ud"]), plan["slots"].tobytes(), int(plan["NDUST"]), bool(plan["has_cia"]))
    return (int(ev.mode), int(ev.ISPACE), float(ev.TSURF), int(ev.NVMR), int(ev.NPAR), tuple(int(g) for g in ev.gas_slot),
            h(ev.EMISSIVITY), h(ev.xfac), h(ev.SOLFLUX), h(ev.REFLECTANCE), ev.EMTEMP is None, ev.SOL_ANG is None, cont)


def combine_evaluations(evs):
    """Several atmosphere states as ONE evaluation (no gradients): the states' layers are laid side by side on the layer
    axis (the k-interp plan, the amounts and the continuum terms are per layer; every (wavenumber, layer) cell of the gas
    opacity is independent) and their paths side by side on the path axis, each path's layer list pointing into its own
    state's block.  This is how the numerical-Jacobian columns of jacobian_nemesis (ForwardModel_0.py:2184-2361: one
    forward model per perturbed state-vector element, joblib workers in the reference) become an extra grid dimension
    of one launch.  Returns (combined Evaluation, [(first path, number of paths)] per state)."""
    if len({batch_key(e) for e in evs}) != 1:
        raise ValueError("combine_evaluations: evaluations differ in a per-evaluation quantity (see batch_key)")
    e0 = evs[0]
    nlay = [len(e.press_atm) for e in evs]
    off = np.concatenate([[0], np.cumsum(nlay)]).astype(np.int64)
    nlm = max(e.LAYINC.shape[0] for e in evs)
    npath = [e.LAYINC.shape[1] for e in evs]
    nw = None
    for e in evs:
        for a in (e.taucia, e.taudust, e.tauray):
            if a is not None:
                nw = a.shape[0]

    def per_layer(name, axis):
        parts = [getattr(e, name) for e in evs]
        if all(p is None for p in parts):
            return None
        full = []
        for e, p in zip(evs, parts):
            if p is None:
                p = np.zeros((nw, len(e.press_atm)))
            full.append(np.asarray(p, dtype=np.float64))
        return np.ascontiguousarray(np.concatenate(full, axis=axis))

    def per_path(name, dtype, shift=False, fill=0):
        parts = [getattr(e, name) for e in evs]
        if all(p is None for p in parts):
            return None
        out = np.full((nlm, sum(npath)), fill, dtype=dtype)
        c = 0
        for i, (e, p) in enumerate(zip(evs, parts)):
            p = np.asarray(p)
            blk = p.astype(dtype) + (off[i] if shift else 0)
            out[:p.shape[0], c:c + p.shape[1]] = blk
            c += p.shape[1]
        return out

    def cat1(name, dtype):
        parts = [getattr(e, name) for e in evs]
        if all(p is None for p in parts):
            return None
        return np.concatenate([np.atleast_1d(np.asarray(p, dtype=dtype)) for p in parts])

    cont = None
    if e0.continuum is not None:
        # continuum plans: everything per layer goes side by side like the layers themselves; the spectra and the
        # resident tables are shared (batch_key)
        tables, p0 = e0.continuum
        plans = [e.continuum[1] for e in evs]
        merged = dict(p0)
        merged["NLAY"] = int(sum(nlay))
        for k in ("pl", "wt"):
            merged[k] = np.ascontiguousarray(np.concatenate([p[k] for p in plans], axis=0))
        for k in ("q1", "q2", "ca", "cb", "vr", "vrd", "vd"):
            merged[k] = np.ascontiguousarray(np.concatenate([p[k] for p in plans], axis=1))
        for k in ("xfac", "totam"):
            merged[k] = np.concatenate([p[k] for p in plans])
        cont = (tables, merged)
    ev = Evaluation(press_atm=np.concatenate([np.asarray(e.press_atm, dtype=np.float64) for e in evs]),
                    temp=np.concatenate([np.asarray(e.temp, dtype=np.float64) for e in evs]),
                    amount=np.ascontiguousarray(np.concatenate([np.asarray(e.amount, dtype=np.float64) for e in evs], axis=1)),
                    gas_slot=e0.gas_slot, NVMR=e0.NVMR, NPAR=e0.NPAR,
                    LAYINC=per_path("LAYINC", np.int32, shift=True), SCALE=per_path("SCALE", np.float64),
                    NLAYIN=cat1("NLAYIN", np.int32), EMTEMP=per_path("EMTEMP", np.float64),
                    LAYPRESS=cat1("LAYPRESS", np.float64), taucia=per_layer("taucia", 1), taudust=per_layer("taudust", 1),
                    tauray=per_layer("tauray", 1), dtaucon=None, mode=e0.mode, ISPACE=e0.ISPACE, TSURF=e0.TSURF,
                    EMISSIVITY=e0.EMISSIVITY, xfac=e0.xfac, SOLFLUX=e0.SOLFLUX, REFLECTANCE=e0.REFLECTANCE,
                    SOL_ANG=cat1("SOL_ANG", np.float64), EMISS_ANG=cat1("EMISS_ANG", np.float64), continuum=cont)
    first = np.concatenate([[0], np.cumsum(npath)])
    return ev, [(int(first[i]), int(npath[i])) for i in range(len(evs))]


class _Stager:
    """Pinned host staging + async H2D on the current stream; counts the bytes it moves."""
    CHUNK_BYTES = 16 << 20

    def __init__(self):
        self.bytes = 0
        self._pinned = {}

    def __call__(self, name, a, dtype=torch.float64):
        if a is None:
            return None
        npdt = np.float64 if dtype == torch.float64 else np.int32
        a = np.ascontiguousarray(a, dtype=npdt)
        slot = self._pinned.get(name)
        if slot is None or slot[0].shape != a.shape or slot[0].dtype != dtype:
            slot = [torch.empty(a.shape, dtype=dtype, pin_memory=True), None]
            self._pinned[name] = slot
        buf, ev = slot
        if ev is not None:
            ev.synchronize()          # the previous async copy out of this pinned buffer must be done
        self.bytes += a.nbytes
        if a.nbytes < self.CHUNK_BYTES:
            buf.numpy()[...] = a
            dev = buf.to("cuda", non_blocking=True)
        else:
            # large inputs (line-by-line grids: continuum terms of 1e5 wavenumbers): the copy into pinned memory
            # is split over torch's host threads and pipelined against the DMA chunk by chunk
            src, pin = torch.from_numpy(a if a.flags.writeable else a.copy()).view(-1), buf.view(-1)
            dev = torch.empty(a.shape, dtype=dtype, device="cuda")
            dflat = dev.view(-1)
            step = self.CHUNK_BYTES // a.itemsize
            for i in range(0, src.numel(), step):
                pin[i:i + step].copy_(src[i:i + step])
                dflat[i:i + step].copy_(pin[i:i + step], non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        return dev


class HotPath:
    """Resident table + device operators for the correlated-k forward model and its Jacobian.

    K      [NWAVE,NG,NP,NT,NGAS] float64 (SpectroscopyX.K);  PRESS (atm) / TEMP (K) / DELG / WAVE are
           taken with the dtypes the live Spectroscopy object holds (float32 after a .kta read).
           A 4-D K[NWAVE,NP,NT,NGAS] is a line-by-line table (ILBL = LINE_BY_LINE_TABLES, NG = 1, DELG = [1]):
           the gas opacity then follows calc_klbl[g] and the plain sum over gases of the LBL branch of
           calculate_gaseous_line_opacity (ForwardModel_0.py:3795-3815) instead of calc_k[g] + k_overlap[g];
           everything downstream is shared.
    """

    layer_space = True      # forward_jacobian accepts Mlay (gradients per layer for decks with many paths)

    def __init__(self, K, PRESS, TEMP, DELG, WAVE, ops=_ops, table_storage="f64"):
        self.ops = ops
        self.lbl_table = len(K.shape) == 4
        if self.lbl_table:
            K = K.reshape(K.shape[0], 1, K.shape[1], K.shape[2], K.shape[3])
            if len(np.asarray(DELG)) != 1:
                raise ValueError("a line-by-line table has one g-ordinate (DELG = [1.0])")
        self.table = ops.Table(K, table_storage) if table_storage != "f64" else ops.Table(K)
        self.NWAVE, self.NG, self.NP, self.NT, self.NGAS = (int(x) for x in K.shape)
        self.PRESS = np.asarray(PRESS)
        self.TEMP = np.asarray(TEMP)
        self.DELG = np.asarray(DELG)
        self.WAVE = np.asarray(WAVE, dtype=np.float64)
        self.otab = ops.OverlapTables(self.DELG)
        self.wave_d = ops.to_dev(self.WAVE)
        self.delg_d = ops.to_dev(self.DELG.astype(np.float64))
        self._stage = _Stager()
        self._copy_stream = None
        self.launches = 0

    # -- host -> device ---------------------------------------------------------------------------
    def stage_opacity(self, ev: Evaluation, return_grad):
        """Inputs of the gas-opacity kernel only (k-interp plan + amounts: a few KB), packed into ONE pinned
        buffer and one host->device copy so that the kernel can be launched after a single enqueue."""
        st = self._stage
        st.bytes = 0
        if self.lbl_table:
            return self._stage_lbl_opacity(ev, return_grad)
        hp = _plan.kinterp_plan(self.PRESS, self.TEMP, ev.press_atm, ev.temp, return_grad)
        n = len(ev.press_atm)
        amount = np.ascontiguousarray(ev.amount, dtype=np.float64)
        ints = np.concatenate([hp["ip_lo"], hp["it_lo"]]).astype(np.int32)
        if ints.size & 1:
            ints = np.concatenate([ints, np.zeros(1, np.int32)])
        parts = [hp["w4"].reshape(-1), hp["omv"], hp["vv"], hp["dudt"], amount.reshape(-1), ints.view(np.float64)]
        packed = st("opacity_inputs", np.concatenate(parts))
        off = 0

        def take(count):
            nonlocal off
            v = packed[off:off + count]
            off += count
            return v
        dev = dict(w4=take(4 * n).view(n, 4), omv=take(n), vv=take(n), dudt=take(n))
        s = Staged()
        s.amount = take(amount.size).view(amount.shape)
        iv = take(ints.size // 2).view(torch.int32)
        dev["ip_lo"], dev["it_lo"] = iv[0:n], iv[n:2 * n]
        s.grad = return_grad
        s.plan_host = hp
        s.dplan = _DevPlan(dev, n)
        s.M = None
        return s

    def _stage_lbl_opacity(self, ev: Evaluation, return_grad):
        """stage_opacity for a line-by-line table: plan.klbl_plan + amounts in one pinned buffer."""
        st = self._stage
        hp = _plan.klbl_plan(self.PRESS, self.TEMP, ev.press_atm, ev.temp, return_grad)
        n = len(ev.press_atm)
        amount = np.ascontiguousarray(ev.amount, dtype=np.float64)
        parts = [hp["w4"].reshape(-1), hp["omv"], hp["vv"], hp["du1dt"], hp["du2dt"], amount.reshape(-1),
                 hp["corner"].reshape(-1).view(np.float64)]
        packed = st("opacity_inputs", np.concatenate(parts))
        off = 0

        def take(count):
            nonlocal off
            v = packed[off:off + count]
            off += count
            return v
        s = Staged()
        d = _LblDevPlan()
        d.NLAY = n
        d.w4, d.omv, d.vv, d.du1dt, d.du2dt = take(4 * n).view(n, 4), take(n), take(n), take(n), take(n)
        s.amount = take(amount.size).view(amount.shape)
        d.corner = take(2 * n).view(torch.int32)
        s.grad, s.plan_host, s.dplan, s.M = return_grad, hp, d, None
        return s

    def stage_radiance(self, s, ev: Evaluation, M=None, Mlay=None):
        """Inputs of the radiance / projection kernels (continuum terms, path, surface, M).  Mlay: the layer-space
        projection matrix (plan.fold_projection_layers); when given and the shape allows (transmission over several
        paths) the gradients are produced per layer and projected with that one matrix instead of the per-path M."""
        st = self._stage
        i32 = torch.int32
        s.gas_slot = st("gas_slot", ev.gas_slot, i32)
        if ev.continuum is not None:
            s.taucia, s.taudust, s.tauray, s.dtaucon = self._continuum_from_plan(ev.continuum, s.grad)
        else:
            s.taucia, s.taudust, s.tauray = st("taucia", ev.taucia), st("taudust", ev.taudust), st("tauray", ev.tauray)
            s.dtaucon = st("dtaucon", ev.dtaucon) if s.grad else None
        s.layinc, s.scale, s.nlayin = st("layinc", ev.LAYINC, i32), st("scale", ev.SCALE), st("nlayin", ev.NLAYIN, i32)
        s.emtemp, s.laypress = st("emtemp", ev.EMTEMP), st("laypress", ev.LAYPRESS)
        s.emissivity, s.xfac = st("emissivity", ev.EMISSIVITY), st("xfac", ev.xfac)
        s.solflux, s.reflectance = st("solflux", ev.SOLFLUX), st("reflectance", ev.REFLECTANCE)
        s.sol_ang, s.emiss_ang = st("sol_ang", ev.SOL_ANG), st("emiss_ang", ev.EMISS_ANG)
        s.mode, s.ISPACE, s.TSURF, s.NVMR, s.NPAR = ev.mode, ev.ISPACE, ev.TSURF, ev.NVMR, ev.NPAR
        s.layer_space = bool(
            Mlay is not None and s.grad and hasattr(self.ops, "radiance_layer_space_ok") and
            self.ops.radiance_layer_space_ok(ev.mode, self.NG, len(ev.press_atm), self.NGAS, ev.NPAR,
                                             ev.LAYINC.shape[1], ev.LAYINC.shape[0], True, s.dtaucon is not None))
        Mh = Mlay if s.layer_space else M
        # The projection matrix is a few per cent non-zeros (a state element acts on the layers of one parameter): while
        # it is small enough for the host to look at for nothing, it goes to the device as a sparse operator by columns
        # (plan.sparse_projection, ~15 KB instead of 0.5 MB) and the projection is a gather per dspec row; otherwise the
        # dense matrix, with the list of its non-empty 16-row chunks when that is cheap to find.
        s.M, s.M_chunks, s.M_sparse = None, None, None
        if Mh is not None:
            Mh = np.asarray(Mh)
            small = Mh.size <= (1 << 19) and hasattr(self.ops, "SparseProjection")
            sp = _plan.sparse_projection(Mh) if small and Mh.shape[1] <= self.ops.SPARSE_MAX_E else None
            if sp is not None:
                s.M_sparse = self.ops.SparseProjection(sp, put=lambda name, a, dt: st(name, a, dt))
            else:
                s.M = st("Mlay" if s.layer_space else "M", Mh)
                if small and hasattr(self.ops, "PROJECT_CHUNK"):
                    c, E = self.ops.PROJECT_CHUNK, Mh.shape[1]
                    rows = np.zeros(((E + c - 1) // c) * c, dtype=bool)
                    rows[:E] = np.any(Mh != 0.0, axis=(0, 2))
                    s.M_chunks = st("Mchunks", np.nonzero(rows.reshape(-1, c).any(axis=1))[0].astype(np.int32), i32)
        ev.h2d_bytes = st.bytes
        return s

    # -- continuum terms from a plan (continuum.py) ------------------------------------------------------
    def continuum_tables(self, key, factory):
        """The resident CIA cross sections for `key` (built by `factory` on a miss; two entries are kept)."""
        with _CONT_LOCK:          # (the forward models of a numerical Jacobian run in threads)
            cache = self.__dict__.setdefault("_cont_tables", [])
            for n, (k, t) in enumerate(cache):
                if k == key:
                    cache.append(cache.pop(n))
                    return t
            t = factory()
            t.device = (self.ops.to_dev(t.kw) if t.kw.size else None,
                        self.ops.to_dev(t.nplanes, torch.int32) if t.nplanes.size else None)
            cache.append((key, t))
            del cache[:-2]
            return t

    def _continuum_from_plan(self, cont, want_grad):
        """Upload the per-layer coefficients and the few spectra of a continuum plan (one pinned buffer, one copy) and
        make taucia / taudust / tauray / dtaucon on the device (ansb200_continuum)."""
        tables, plan = cont
        if tables.device is None:
            tables.device = (self.ops.to_dev(tables.kw) if tables.kw.size else None,
                             self.ops.to_dev(tables.nplanes, torch.int32) if tables.nplanes.size else None)
        fkeys = ("wt", "q1", "q2", "ca", "cb", "xfac", "totam", "ur", "vr", "vrd", "ud", "vd")
        ikeys = ("pl", "slots")
        ints = np.concatenate([np.ascontiguousarray(plan[k], dtype=np.int32).reshape(-1) for k in ikeys])
        if ints.size & 1:
            ints = np.concatenate([ints, np.zeros(1, np.int32)])
        packed = self._stage("continuum_plan", np.concatenate(
            [np.ascontiguousarray(plan[k], dtype=np.float64).reshape(-1) for k in fkeys] + [ints.view(np.float64)]))
        dev = {k: plan[k] for k in ("NLAY", "NVMR", "NDUST", "NTERM", "has_cia")}
        off = 0
        for k in fkeys:
            a = plan[k]
            dev[k] = packed[off:off + a.size].view(a.shape) if a.size else torch.empty(a.shape, dtype=torch.float64, device="cuda")
            off += a.size
        iv = packed[off:].view(torch.int32)
        o2 = 0
        for k in ikeys:
            a = plan[k]
            dev[k] = iv[o2:o2 + a.size].view(a.shape) if a.size else torch.empty(a.shape, dtype=torch.int32, device="cuda")
            o2 += a.size
        out = self.ops.continuum(tables.device[0], tables.device[1], dev, self.NWAVE, want_grad)
        self.launches += 1
        return out

    def stage(self, ev: Evaluation, return_grad, M=None, Mlay=None):
        """Copy one evaluation's inputs (pinned staging, async on the current stream) and build the
        k-interp plan.  Returns a Staged object; ev.h2d_bytes records the bytes moved."""
        return self.stage_radiance(self.stage_opacity(ev, return_grad), ev, M, Mlay)

    # -- device stages ---------------------------------------------------------------------------
    def gas_opacity(self, s, timers=None):
        """calc_k[g] + k_overlap[g] fused on the device -> tau[NWAVE,NG,NLAY] (, dk[...,NGAS+1])."""
        if timers is not None:
            timers[0].record()
        if self.lbl_table:
            out = self.ops.lbl_table_opacity(self.table, s.dplan, s.amount, s.grad)
        else:
            out = self.ops.gas_opacity(self.table, s.dplan, s.amount, self.otab, s.grad)
        if timers is not None:
            timers[1].record()
        if self.lbl_table:
            self.launches += 1
        else:
            self.launches += self.ops.overlap_kernel_launches(self.otab.NG, self.table.shape[4], self.otab.seq)
        return out

    def run(self, s, timers=None):
        """Kernels only, inputs already resident: gas_opacity -> radiance (-> jacobian_project if s.M).
        Device equivalent of ForwardModel_0.CIRSrad (archnemesis/ForwardModel_0.py:4376-4511) for the
        thermal-emission and pure-transmission path types, followed by map2pro/map2xvec (:5319-5424)."""
        return self.finish(s, self.gas_opacity(s, timers))

    def finish(self, s, go):
        tau, dk = go if s.grad else (go, None)
        out = self.ops.radiance(s.mode, tau, dk, s.gas_slot, s.taucia, s.taudust, s.tauray, s.dtaucon, s.layinc, s.scale,
                                s.nlayin, s.emtemp, s.laypress, self.wave_d, self.delg_d, s.emissivity, s.xfac,
                                s.solflux, s.reflectance, s.sol_ang, s.emiss_ang, s.ISPACE, s.TSURF, s.NVMR, s.NPAR,
                                s.grad, **({"layer_space": True} if getattr(s, "layer_space", False) else {}))
        self.launches += 1
        if not s.grad:
            return out
        spec, dspec, dtsurf = out
        if getattr(s, "M_sparse", None) is not None:
            dx = self.ops.jacobian_project_sparse(dspec, s.M_sparse, shared=bool(getattr(s, "layer_space", False)))
            self.launches += 1
            return spec, dx, dtsurf
        if s.M is None:
            return spec, dspec, dtsurf
        kw = {"shared": True} if getattr(s, "layer_space", False) else {}
        if getattr(s, "M_chunks", None) is not None:
            kw["chunks"] = s.M_chunks
        dx = self.ops.jacobian_project(dspec, s.M, **kw)
        self.launches += 1
        return spec, dx, dtsurf

    # -- host-facing calls -------------------------------------------------------------------------
    def _evaluate(self, ev: Evaluation, return_grad, M=None, Mlay=None):
        """stage -> kernels with the bulk of the host->device traffic (continuum terms, 36 MB at config 2)
        hidden behind the gas-opacity kernel: that kernel needs only the plan and the amounts, so it is
        launched first and the remaining inputs are staged on a side stream while it runs."""
        s = self.stage_opacity(ev, return_grad)
        go = self.gas_opacity(s)
        main = torch.cuda.current_stream()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        with torch.cuda.stream(self._copy_stream):
            self.stage_radiance(s, ev, M, Mlay)
            done = torch.cuda.Event()
            done.record()
        for t in vars(s).values():
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(main)
        main.wait_event(done)
        return self.finish(s, go)

    def cirsrad(self, ev: Evaluation, return_grad=False):
        """spec[NWAVE,NPATH] (, dspec[NWAVE,NPATH,NPAR,NLAYMAX], dtsurf[NWAVE,NPATH]) as device tensors."""
        return self._evaluate(ev, return_grad)

    def forward_jacobian(self, ev: Evaluation, M, Mlay=None):
        """Spectrum and state-vector Jacobian; layer-space gradients never leave the device:
        CIRSrad(return_grad=True) -> map2pro -> map2xvec of nemesisfmg (ForwardModel_0.py:694-714).
        M = plan.fold_projection(...) [NPATH, NPAR*NLAYMAX, NX]; Mlay = plan.fold_projection_layers(...)
        [1, NPAR*NLAY, NX] lets decks with many paths take the per-layer route (see stage_radiance).  Returns
        device tensors spec[NWAVE,NPATH], dspec_x[NWAVE,NPATH,NX], dtsurf[NWAVE,NPATH]."""
        return self._evaluate(ev, True, M, Mlay)

    def forward_jacobian_conv(self, ev: Evaluation, M, conv_op, jsurf=-1, wgeom=1.0):
        """forward_jacobian followed on the device by what nemesisfmg does with its result for one geometry
        without averaging (ForwardModel_0.py:716-768): the surface-temperature column (dSPEC1[:,0,JSURF] =
        dTSURF), the WGEOM weight, and the instrument line shape of Measurement_0.convg (k-tables, FWHM <= 0;
        ``conv_op`` = ops.ConvOperator(plan.conv_operator(...))).  Returns one device tensor
        [NCONV, 1+NX] = [SPECONV | dSPECONV]: the only array that has to travel back to the host."""
        spec, dx, dtsurf = self._evaluate(ev, True, M)
        nw, nx = dx.shape[0], dx.shape[2]
        block = torch.empty((nw, nx + 1), dtype=torch.float64, device="cuda")
        block[:, 0] = spec[:, 0]
        block[:, 1:] = dx[:, 0, :]
        if jsurf >= 0:
            block[:, 1 + jsurf] = dtsurf[:, 0]
        if wgeom != 1.0:
            block *= float(wgeom)
        self.launches += 1
        return self.ops.convolve(conv_op, block)

    def forward_jacobian_mix_conv(self, ev: Evaluation, M, mix, conv_op, Mlay=None):
        """The limb / occultation drivers' tail on the device (nemesisSOfmg / nemesisLfmg, ForwardModel_0.py:1186-1243,
        :1444-1513): forward_jacobian over all paths, the tangent-height interpolation of the path spectra
        (mix = plan.tangent_mix) and the line shape / filter integral applied to every geometry at once (IGEOM='All':
        one operator, geometry 0's grid).  Returns device tensors specmod[NWAVE,NGEOM] (subspeconv wants it) and
        [NCONV, NGEOM, 1+NX] = [SPECONV | dSPECONV]."""
        spec, dx, _ = self._evaluate(ev, True, M, Mlay)
        block = self.ops.path_mix(spec, dx, mix)
        nw, ngeom, nc = block.shape
        self.launches += 2
        out = self.ops.convolve(conv_op, block.view(nw, ngeom * nc), col0_is_spectrum=False)
        return block[:, :, 0], out.view(out.shape[0], ngeom, nc)

    def project(self, dspec, M):
        """Layer-space gradients still on the device x host-folded projection matrix -> dspec_x[NWAVE,NPATH,NX]
        (map2pro + map2xvec, ForwardModel_0.py:5319-5424)."""
        self.launches += 1
        return self.ops.jacobian_project(dspec, self.ops.to_dev(M))

    def conv_operator(self, op):
        """Device copy of a plan.conv_operator (kept by the caller for the life of a retrieval)."""
        return self.ops.ConvOperator(op)

    @staticmethod
    def to_host(t):
        """Device tensor -> numpy array (synchronises the current stream).  Results up to 64 MB go through a cached
        pinned buffer (one DMA, then a host copy out of it) instead of a pageable-memory copy."""
        t = t.detach()
        if not t.is_cuda or t.numel() == 0 or t.numel() * t.element_size() > _PINNED_OUT_MAX:
            return t.cpu().numpy()
        key = (tuple(t.shape), t.dtype)
        buf = _PINNED_OUT.get(key)
        if buf is None:
            if len(_PINNED_OUT) >= 16:
                _PINNED_OUT.clear()
            buf = _PINNED_OUT[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        buf.copy_(t.contiguous(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return buf.numpy().copy()

    def close(self):
        self.table.close()


class Staged:
    """Device-resident inputs of one evaluation (see HotPath.stage)."""


class _LblDevPlan:
    """Device views of a plan.klbl_plan inside the packed staging buffer."""


class _DevPlan:
    def __init__(self, dev, nlay):
        self.NLAY = nlay
        self.ip_lo, self.it_lo = dev["ip_lo"], dev["it_lo"]
        self.w4, self.omv, self.vv, self.dudt = dev["w4"], dev["omv"], dev["vv"], dev["dudt"]
