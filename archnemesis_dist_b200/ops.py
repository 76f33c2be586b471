"""Device-side operators: thin Python wrappers that marshal torch CUDA tensors (buffer carriers
only) into the C ABI of libansb200.so.  Every function enqueues on torch's current stream and
returns device tensors; nothing here computes on the CPU and nothing falls back to it."""
import ctypes

import numpy as np
import torch

from . import _lib
from . import plan as _plan


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("archnemesis_dist_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-contiguous tensor required"
    return ctypes.c_void_p(t.data_ptr())


def to_dev(a, dtype=torch.float64):
    """Host array (or tensor) -> contiguous device tensor of `dtype`."""
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype).contiguous()
    npdt = {torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}[dtype]
    h = np.ascontiguousarray(a, dtype=npdt)
    if not h.flags.writeable:           # e.g. the memoised, shared WAVE grid: torch refuses read-only views
        h = h.copy()
    return torch.from_numpy(h).cuda()


class Table:
    """Device-resident k-table (K and ln K), replacing the per-call read_tables of
    archnemesis/Spectroscopy_0.py:1448-1528.  `K` is [NWAVE,NG,NP,NT,NGAS] float64 (host array
    or device tensor).  `storage`: "f64" (default), "k32" (K as float32 -- lossless for .kta data, bit-identical
    results, 12 instead of 16 bytes per entry) or "f32" (K and ln K as float32: the FP32 k-interp variant, ~4e-6
    relative in k); see ansb200_table_create_ex."""

    STORAGE = {"f64": _lib.TABLE_F64, "k32": _lib.TABLE_K32, "f32": _lib.TABLE_F32}

    def __init__(self, K, storage="f64"):
        _require_cuda()
        lib = _lib.load()
        self.shape = tuple(int(x) for x in K.shape)
        assert len(self.shape) == 5, "K must be [NWAVE,NG,NP,NT,NGAS]"
        h = ctypes.c_void_p()
        if isinstance(K, torch.Tensor):
            Kd = K.to(device="cuda", dtype=torch.float64).contiguous()
            src, is_dev = ctypes.c_void_p(Kd.data_ptr()), 1
        else:
            Kh = np.ascontiguousarray(K, dtype=np.float64)
            src, is_dev = ctypes.c_void_p(Kh.ctypes.data), 0
        self.storage = storage
        _lib.check(lib.ansb200_table_create_ex(src, is_dev, *self.shape, self.STORAGE[storage], ctypes.byref(h), _stream()))
        torch.cuda.current_stream().synchronize()   # the source buffer may be released now
        self._h = h
        self.nbytes = {"f64": 16, "k32": 12, "f32": 8}[storage] * int(np.prod(self.shape))

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("table already destroyed")
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().ansb200_table_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DevicePlan:
    """kinterp plan (plan.kinterp_plan) uploaded to the device."""

    def __init__(self, host_plan, grad):
        self.host = host_plan
        self.grad = grad
        self.NLAY = len(host_plan["ip_lo"])
        self.ip_lo = to_dev(host_plan["ip_lo"], torch.int32)
        self.it_lo = to_dev(host_plan["it_lo"], torch.int32)
        self.w4 = to_dev(host_plan["w4"])
        self.omv = to_dev(host_plan["omv"])
        self.vv = to_dev(host_plan["vv"])
        self.dudt = to_dev(host_plan["dudt"])


def kinterp(table, dplan, want_grad=False):
    """calc_k / calc_kg on the device: k[NWAVE,NG,NLAY,NGAS] (and dkdT)."""
    _require_cuda()
    NWAVE, NG, NP, NT, NGAS = table.shape
    k = torch.empty((NWAVE, NG, dplan.NLAY, NGAS), dtype=torch.float64, device="cuda")
    dkdT = torch.empty_like(k) if want_grad else None
    _lib.check(_lib.load().ansb200_kinterp(table.handle, dplan.NLAY, _ptr(dplan.ip_lo), _ptr(dplan.it_lo), _ptr(dplan.w4),
                                           _ptr(dplan.omv), _ptr(dplan.vv), _ptr(dplan.dudt), int(want_grad),
                                           _ptr(k), _ptr(dkdT), _stream()))
    return (k, dkdT) if want_grad else k


class OverlapTables:
    def __init__(self, del_g):
        w, g, seq = _plan.overlap_tables(del_g)
        self.weight = to_dev(w)
        self.g_ord = to_dev(g)
        self.del_g = to_dev(np.asarray(del_g, dtype=np.float64))
        self.seq = seq
        self.NG = len(g) - 1


def koverlap(k, amount, otab, dkdT=None, force_seq=False):
    """k_overlap / k_overlapg on the device.  k[NWAVE,NG,NLAY,NGAS], amount[NGAS,NLAY] (cm-2)."""
    _require_cuda()
    NWAVE, NG, NLAY, NGAS = k.shape
    grad = dkdT is not None
    tau = torch.empty((NWAVE, NG, NLAY), dtype=torch.float64, device="cuda")
    dk = torch.empty((NWAVE, NG, NLAY, NGAS + 1), dtype=torch.float64, device="cuda") if grad else None
    flag = int(grad) | (2 if (otab.seq or force_seq) else 0)
    _lib.check(_lib.load().ansb200_koverlap(_ptr(k), _ptr(dkdT), _ptr(amount), _ptr(otab.weight), _ptr(otab.g_ord),
                                            _ptr(otab.del_g), NWAVE, NG, NLAY, NGAS, flag, _ptr(tau), _ptr(dk), _stream()))
    return (tau, dk) if grad else tau


def overlap_mode(mode=-1):
    """Diagnostics switch of the overlap kernels (include/ansb200.h): 0 = fast kernel + work list for the general one
    (default), 1 = general kernel only, 2 = as 0 with per-call statistics.  Returns the previous mode."""
    return int(_lib.load().ansb200_overlap_mode(int(mode)))


def overlap_kernel_launches(NG, NGAS, seq=False):
    """Kernels one ansb200_koverlap / ansb200_gas_opacity call launches: the fast kernel plus the general one on its
    work list where a fast kernel exists (csrc/koverlap_fast.cu: ov_fast_supported), else the general kernel alone."""
    fast = NG == 20 and 2 <= NGAS <= 14 and not seq and overlap_mode() != 1
    return 2 if fast else 1


def overlap_stats():
    """Counts of the last overlap call made in mode 2: cells handed to the general kernel and why, kinds of folds."""
    import ctypes
    buf = (ctypes.c_int32 * 9)()
    _lib.load().ansb200_overlap_stats(buf)
    names = ("handed_over", "non_monotone", "open_bin", "key_group", "exact_tie", "bins_not_monotone", "static_folds",
             "sorted_folds", "static_rejected")
    return dict(zip(names, [int(v) for v in buf]))


def gas_opacity(table, dplan, amount, otab, want_grad=False, force_seq=False):
    """Fused calc_k[g] + k_overlap[g] (K_TABLES branch of calculate_gaseous_line_opacity)."""
    _require_cuda()
    NWAVE, NG, NP, NT, NGAS = table.shape
    NLAY = dplan.NLAY
    tau = torch.empty((NWAVE, NG, NLAY), dtype=torch.float64, device="cuda")
    dk = torch.empty((NWAVE, NG, NLAY, NGAS + 1), dtype=torch.float64, device="cuda") if want_grad else None
    flag = int(want_grad) | (2 if (otab.seq or force_seq) else 0)
    _lib.check(_lib.load().ansb200_gas_opacity(table.handle, NLAY, _ptr(dplan.ip_lo), _ptr(dplan.it_lo), _ptr(dplan.w4),
                                               _ptr(dplan.omv), _ptr(dplan.vv), _ptr(dplan.dudt), _ptr(amount),
                                               _ptr(otab.weight), _ptr(otab.g_ord), _ptr(otab.del_g), flag, _ptr(tau), _ptr(dk),
                                               _stream()))
    return (tau, dk) if want_grad else tau


def continuum(kw, nplanes, plan, NWAVE, want_grad):
    """ansb200_continuum: dense continuum arrays on the device from a plan (continuum.build_plan uploaded: a dict of
    device tensors with the plan's keys).  Returns taucia (None without a CIA object), taudust, tauray [NWAVE,NLAY] and
    dtaucon[NWAVE,NPAR,NLAY] (None unless want_grad)."""
    _require_cuda()
    NLAY, NVMR, NDUST, NTERM = int(plan["NLAY"]), int(plan["NVMR"]), int(plan["NDUST"]), int(plan["NTERM"])
    NR = int(plan["ur"].shape[0])
    NPL = int(kw.shape[1]) if kw is not None and kw.dim() == 3 else 1
    has_cia = bool(plan["has_cia"])
    taucia = torch.empty((NWAVE, NLAY), dtype=torch.float64, device="cuda") if has_cia else None
    taudust = torch.empty((NWAVE, NLAY), dtype=torch.float64, device="cuda")
    tauray = torch.empty((NWAVE, NLAY), dtype=torch.float64, device="cuda")
    dtaucon = torch.empty((NWAVE, NVMR + 2 + NDUST, NLAY), dtype=torch.float64, device="cuda") if want_grad else None
    _lib.check(_lib.load().ansb200_continuum(
        _ptr(kw), _ptr(nplanes), NTERM, NPL, _ptr(plan["pl"]), _ptr(plan["wt"]), _ptr(plan["q1"]), _ptr(plan["q2"]),
        _ptr(plan["ca"]), _ptr(plan["cb"]), _ptr(plan["slots"]), _ptr(plan["xfac"]), _ptr(plan["totam"]),
        _ptr(plan["ur"]), _ptr(plan["vr"]), _ptr(plan["vrd"]), NR, _ptr(plan["ud"]), _ptr(plan["vd"]), NDUST, NWAVE, NLAY,
        NVMR, int(has_cia), int(bool(want_grad)), _ptr(taucia), _ptr(taudust), _ptr(tauray), _ptr(dtaucon), _stream()))
    return taucia, taudust, tauray, dtaucon


class LblDevicePlan:
    """plan.klbl_plan uploaded to the device."""

    def __init__(self, host_plan):
        self.host = host_plan
        self.NLAY = len(host_plan["corner"])
        self.corner = to_dev(host_plan["corner"], torch.int32)
        self.w4 = to_dev(host_plan["w4"])
        self.omv, self.vv = to_dev(host_plan["omv"]), to_dev(host_plan["vv"])
        self.du1dt, self.du2dt = to_dev(host_plan["du1dt"]), to_dev(host_plan["du2dt"])


def lbl_table_opacity(table, dplan, amount, want_grad=False):
    """calc_klbl[g] + the LBL-table branch of calculate_gaseous_line_opacity (ForwardModel_0.py:3795-3815):
    tau[NWAVE,1,NLAY] (, dk[NWAVE,1,NLAY,NGAS+1]) from a resident NG = 1 table."""
    _require_cuda()
    NWAVE, NG, NP, NT, NGAS = table.shape
    if NG != 1:
        raise ValueError("lbl_table_opacity: the table must have NG = 1")
    NLAY = dplan.NLAY
    tau = torch.empty((NWAVE, 1, NLAY), dtype=torch.float64, device="cuda")
    dk = torch.empty((NWAVE, 1, NLAY, NGAS + 1), dtype=torch.float64, device="cuda") if want_grad else None
    _lib.check(_lib.load().ansb200_lbl_table_opacity(table.handle, NLAY, _ptr(dplan.corner), _ptr(dplan.w4), _ptr(dplan.omv),
                                                     _ptr(dplan.vv), _ptr(dplan.du1dt), _ptr(dplan.du2dt), _ptr(amount),
                                                     int(want_grad), _ptr(tau), _ptr(dk), _stream()))
    return (tau, dk) if want_grad else tau


THERMAL, TRANSMISSION = 0, 1


def radiance(mode, tau, dk, gas_slot, taucia, taudust, tauray, dtaucon, layinc, scale, nlayin, emtemp, laypress,
             wave, delg, emissivity, xfac, solflux, reflectance, sol_ang, emiss_ang, ispace, tsurf, NVMR, NPAR,
             want_grad, nan_to_num=True, layer_space=False):
    """Path radiance (+ layer-space Jacobian) for all paths; see include/ansb200.h.
    Returns spec[NWAVE,NPATH] and, with gradients, dspec[NWAVE,NPATH,NPAR,NLAYMAX], dtsurf[NWAVE,NPATH].
    layer_space: gradients per layer instead of per path position, dspec[NWAVE,NPATH,NPAR,NLAY] (the visits of a layer
    added; see radiance_layer_space_ok)."""
    _require_cuda()
    NWAVE, NG, NLAY = tau.shape
    NLAYMAX, NPATH = layinc.shape
    NGAS = 0 if gas_slot is None else int(gas_slot.numel())
    spec = torch.empty((NWAVE, NPATH), dtype=torch.float64, device="cuda")
    dspec = dtsurf = None
    flags = 0
    if want_grad:
        flags |= _lib.RAD_GRAD
        if nan_to_num:
            flags |= _lib.RAD_NAN_TO_NUM
        if layer_space:
            flags |= _lib.RAD_LAYER_SPACE
        dspec = torch.empty((NWAVE, NPATH, NPAR, NLAY if layer_space else NLAYMAX), dtype=torch.float64, device="cuda")
        dtsurf = torch.zeros((NWAVE, NPATH), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().ansb200_radiance(
        int(mode), flags, _ptr(tau), _ptr(dk), _ptr(gas_slot), _ptr(taucia), _ptr(taudust), _ptr(tauray),
        _ptr(dtaucon), _ptr(layinc), _ptr(scale), _ptr(nlayin), _ptr(emtemp), _ptr(laypress), _ptr(wave), _ptr(delg),
        _ptr(emissivity), _ptr(xfac), _ptr(solflux), _ptr(reflectance), _ptr(sol_ang), _ptr(emiss_ang), int(ispace),
        float(tsurf), NWAVE, NG, NLAY, NGAS, int(NVMR), int(NPAR), NLAYMAX, NPATH, _ptr(spec), _ptr(dspec),
        _ptr(dtsurf), _stream()))
    return (spec, dspec, dtsurf) if want_grad else spec


def radiance_layer_space_ok(mode, NG, NLAY, NGAS, NPAR, NPATH, NLAYMAX, has_dk=True, has_dtaucon=True):
    """Can ansb200_radiance hand back layer-space gradients for this shape (>= 4 paths; transmission, or thermal emission
    with at most 224 positions per path)?"""
    return bool(_lib.load().ansb200_radiance_layer_space(int(mode), _lib.RAD_GRAD, int(NG), int(NLAY), int(NGAS), int(NPAR),
                                                         int(NPATH), int(NLAYMAX), int(has_dk), int(has_dtaucon)))


PROJECT_CHUNK = 16


def project_chunks(M):
    """The 16-row chunks of a host projection matrix M[P, E, NX] that hold a non-zero for any path or column, as the
    int32 device list jacobian_project(..., chunks=) takes (whole parameters without a state-vector element drop out)."""
    M = np.asarray(M)
    E = M.shape[1]
    n = (E + PROJECT_CHUNK - 1) // PROJECT_CHUNK
    rows = np.any(M != 0.0, axis=(0, 2))
    pad = np.zeros(n * PROJECT_CHUNK, dtype=bool)
    pad[:E] = rows
    return to_dev(np.nonzero(pad.reshape(n, PROJECT_CHUNK).any(axis=1))[0].astype(np.int32), torch.int32)


def jacobian_project(dspec, M, shared=False, chunks=None):
    """dspec[NWAVE,NPATH,NPAR,NLAYMAX] x M[NPATH,NPAR*NLAYMAX,NX] -> [NWAVE,NPATH,NX] (map2pro+map2xvec).
    shared: M[1,NPAR*NLAY,NX] is one layer-space matrix for every path (dspec from radiance(layer_space=True)).
    chunks: project_chunks(M) of the host copy of M (optional; rows outside the listed chunks are not read)."""
    _require_cuda()
    if dspec.dim() != 4 or M.dim() != 3:
        raise ValueError("jacobian_project: dspec must be [NWAVE,NPATH,NPAR,NLAYMAX] and M [NPATH,NPAR*NLAYMAX,NX]")
    NWAVE, NPATH, NPAR, NLM = dspec.shape
    if M.shape[0] != (1 if shared else NPATH) or M.shape[1] != NPAR * NLM:
        # (the reference raises a tensordot shape error when xmap and the gradient array disagree, ForwardModel_0.py:5420)
        raise ValueError("jacobian_project: M is %s, expected (%d, %d, NX) for dspec %s"
                         % (tuple(M.shape), NPATH, NPAR * NLM, tuple(dspec.shape)))
    for name, t in (("dspec", dspec), ("M", M)):
        if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
            raise ValueError("jacobian_project: %s must be a contiguous float64 CUDA tensor" % name)
    NX = M.shape[2]
    out = torch.empty((NWAVE, NPATH, NX), dtype=torch.float64, device="cuda")
    if chunks is not None:
        _lib.check(_lib.load().ansb200_jacobian_project_chunks(_ptr(dspec), _ptr(M), NWAVE, NPAR, NLM, NPATH, NX, int(bool(shared)),
                                                               _ptr(chunks) if chunks.numel() else None, int(chunks.numel()),
                                                               _ptr(out), _stream()))
        return out
    fn = _lib.load().ansb200_jacobian_project_shared if shared else _lib.load().ansb200_jacobian_project
    _lib.check(fn(_ptr(dspec), _ptr(M), NWAVE, NPAR, NLM, NPATH, NX, _ptr(out), _stream()))
    return out


def path_mix(spec, dx, mix):
    """Tangent-height interpolation of the path spectra (nemesisSOfmg / nemesisLfmg, ForwardModel_0.py:1206-1228):
    spec[NWAVE,NPATH], dx[NWAVE,NPATH,NX] on the device, mix = plan.tangent_mix(...) -> [NWAVE,NGEOM,1+NX]."""
    _require_cuda()
    NWAVE, NPATH, NX = dx.shape
    if tuple(spec.shape) != (NWAVE, NPATH):
        raise ValueError("path_mix: spec must be [NWAVE,NPATH] like dx[NWAVE,NPATH,NX]")
    lo, hi = np.asarray(mix["lo"], np.int32), np.asarray(mix["hi"], np.int32)
    if lo.min() < 0 or max(lo.max(), hi.max()) >= NPATH:
        raise ValueError("path_mix: path index out of range")
    NGEOM = len(lo)
    out = torch.empty((NWAVE, NGEOM, NX + 1), dtype=torch.float64, device="cuda")
    lo_d, hi_d = to_dev(lo, torch.int32), to_dev(hi, torch.int32)
    wl_d, wh_d = to_dev(np.asarray(mix["wlo"], np.float64)), to_dev(np.asarray(mix["whi"], np.float64))
    _lib.check(_lib.load().ansb200_path_mix(_ptr(spec.contiguous()), _ptr(dx.contiguous()), _ptr(lo_d), _ptr(hi_d),
                                            _ptr(wl_d), _ptr(wh_d), NWAVE, NPATH, NX, NGEOM, _ptr(out), _stream()))
    return out


def kdist_capacity(weighted):
    return int(_lib.load().ansb200_kdist_capacity(int(bool(weighted))))


def kdist(kabs, lo, hi, g_ord, w=None, woff=None):
    """k-distributions of spectral bins (tail of calc_ktable_chunk, Spectroscopy_0.py:3619-3660): kabs[ncalc] on the
    device, bins [lo[b], hi[b]) of the grid, optional weights w (bin b at w[woff[b]:...]) -> out[NBIN, NG] (device)."""
    _require_cuda()
    lo, hi = np.asarray(lo, np.int32), np.asarray(hi, np.int32)
    n = hi - lo
    if len(lo) == 0 or n.min() < 1:
        raise ValueError("kdist: every bin needs at least one grid point")
    g = to_dev(np.asarray(g_ord, np.float64))
    out = torch.empty((len(lo), g.numel()), dtype=torch.float64, device="cuda")
    lo_d, hi_d = to_dev(lo, torch.int32), to_dev(hi, torch.int32)
    w_d = woff_d = None
    if w is not None:
        w_d = w if isinstance(w, torch.Tensor) else to_dev(np.asarray(w, np.float64))
        woff_d = to_dev(np.asarray(woff, np.int64), torch.int64)
    _lib.check(_lib.load().ansb200_kdist(_ptr(kabs), _ptr(w_d) if w_d is not None else None, _ptr(lo_d), _ptr(hi_d),
                                         _ptr(woff_d) if woff_d is not None else None, len(lo), int(n.max()), _ptr(g),
                                         int(g.numel()), _ptr(out), _stream()))
    return out


class SparseProjection:
    """Device copy of a plan.sparse_projection (`put` lets the engine stage the arrays through its pinned buffers)."""

    def __init__(self, sp, put=None):
        put = put or (lambda name, a, dt: to_dev(a, dt))
        self.P, self.E, self.NX = sp["shape"]
        self.col_r0 = put("sp_col_r0", sp["col_r0"], torch.int32)
        self.col_len = put("sp_col_len", sp["col_len"], torch.int32)
        self.col_voff = put("sp_col_voff", sp["col_voff"], torch.int32)
        self.vals = put("sp_vals", sp["vals"], torch.float64)
        self.long_cols = put("sp_long_cols", sp["long_cols"], torch.int32)
        self.long_ptr = put("sp_long_ptr", sp["long_ptr"], torch.int32)


def jacobian_project_sparse(dspec, sp, shared=False):
    """dspec[NWAVE,NPATH,NPAR,NLAYMAX] x sparse M (SparseProjection) -> [NWAVE,NPATH,NX]."""
    _require_cuda()
    NWAVE, NPATH, NPAR, NLM = dspec.shape
    if sp.E != NPAR * NLM or sp.P != (1 if shared else NPATH):
        raise ValueError("jacobian_project_sparse: operator is for (%d, %d, NX), dspec is %s" % (sp.P, sp.E, tuple(dspec.shape)))
    if dspec.dtype != torch.float64 or not dspec.is_cuda or not dspec.is_contiguous():
        raise ValueError("jacobian_project_sparse: dspec must be a contiguous float64 CUDA tensor")
    out = torch.empty((NWAVE, NPATH, sp.NX), dtype=torch.float64, device="cuda")
    p = lambda t: _ptr(t) if t.numel() else None      # noqa: E731
    _lib.check(_lib.load().ansb200_jacobian_project_sparse(_ptr(dspec), _ptr(sp.col_r0), _ptr(sp.col_len), _ptr(sp.col_voff),
                                                           p(sp.vals), int(sp.vals.numel()), p(sp.long_cols), _ptr(sp.long_ptr),
                                                           NWAVE, NPAR, NLM, NPATH, sp.NX, int(bool(shared)), _ptr(out),
                                                           _stream()))
    return out


SPARSE_MAX_E = 3200


class ConvOperator:
    """Device copy of a plan.conv_operator (Measurement_0.conv / convg for k-tables)."""

    def __init__(self, op):
        self.mode, self.NCONV = int(op["mode"]), int(op["NCONV"])
        self.row_start = to_dev(op["row_start"], torch.int32)
        self.widx = to_dev(op["widx"], torch.int32)
        self.wval = to_dev(op["wval"])
        self.norm = to_dev(op["norm"])
        self.np_lo = to_dev(op["np_lo"], torch.int32)
        self.np_exact = to_dev(op["np_exact"], torch.int32)
        self.xinfo = to_dev(op["xinfo"])
        self.weighted_sum_only = bool(op.get("weighted_sum_only", False))
        self.np_interp_all = bool(op.get("np_interp_all", False))       # lblconv / lblconvg with FWHM == 0


def convolve(cop, block, col0_is_spectrum=True):
    """Apply the instrument line shape to block[NWAVE, NCOL] (e.g. [spectrum | Jacobian columns]) ->
    [NCONV, NCOL].  With col0_is_spectrum column 0 is evaluated like the reference's 1-D interpolation."""
    _require_cuda()
    if block.dim() != 2 or block.stride(1) != 1:
        raise ValueError("convolve: block must be 2-D with unit column stride")
    NWAVE, NCOL = block.shape
    out = torch.empty((cop.NCONV, NCOL), dtype=torch.float64, device="cuda")
    if not block.is_cuda or block.dtype != torch.float64:
        raise ValueError("convolve: block must be a float64 device tensor")
    _lib.check(_lib.load().ansb200_convolve(ctypes.c_void_p(block.data_ptr()), NWAVE, NCOL, block.stride(0), cop.mode,
                                            2 if cop.np_interp_all else int(bool(col0_is_spectrum) and not cop.weighted_sum_only),
                                            _ptr(cop.row_start), _ptr(cop.widx), _ptr(cop.wval), _ptr(cop.norm), _ptr(cop.np_lo), _ptr(cop.np_exact),
                                            _ptr(cop.xinfo), cop.NCONV, _ptr(out), _stream()))
    return out
