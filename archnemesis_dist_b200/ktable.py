"""k-table generation on the device: the drop-in for ``Spectroscopy_0.calc_ktable_chunk``.

The reference computes, for every (p, T) point of the table, a line-by-line spectrum of the chunk
(``calc_klbl_online`` -- under ``install()`` that already runs on the device, linedata.py) and then, per spectral bin,
masks the whole grid, argsorts the bin and interpolates its cumulative distribution at the g-ordinates
(archnemesis/Spectroscopy_0.py:3558-3667; O(NBIN x ncalc) comparisons per point in numpy).  ``calc_ktable_chunk`` below
is that function with the per-bin tail replaced by one ``ansb200_kdist`` launch per (p, T) point: the masks become index
ranges of the ascending grid (two ``searchsorted`` calls for all bins), the instrument weights -- if a Measurement is
given -- are evaluated on the host per bin with the reference's own ``np.interp`` and travel as one array.  Bins with
more points than the shared-memory sort holds (ops.kdist_capacity) are done with the library's sort on the device.

No CPU fallback: without CUDA ``k_distribution`` raises (tests substitute an oracle-backed namespace).
"""
import sys

import numpy as np

from . import ops as _ops

_INSTALLED = {}


class DeviceBackend:
    """k_distribution on the device."""

    def k_distribution(self, kabs, wavecalc, vbinmin, vbinmax, g_ord, ils=None):
        import torch
        wavecalc = np.asarray(wavecalc, dtype=np.float64)
        lo, hi = bin_ranges(wavecalc, vbinmin, vbinmax)
        if np.any(hi - lo < 1):
            raise ValueError("calc_ktable: a spectral bin holds no line-by-line grid point")
        delv = wavecalc[1] - wavecalc[0]
        w = woff = None
        if ils is not None:
            parts = [np.asarray(ils(ib, wavecalc[lo[ib]:hi[ib]]), dtype=np.float64) * delv for ib in range(len(lo))]
            woff = np.concatenate([[0], np.cumsum([len(p) for p in parts])[:-1]]).astype(np.int64)
            w = np.concatenate(parts)
        kd = kabs if isinstance(kabs, torch.Tensor) else _ops.to_dev(np.asarray(kabs, dtype=np.float64))
        cap = _ops.kdist_capacity(ils is not None)
        n = hi - lo
        small = np.nonzero(n <= cap)[0]
        out = torch.empty((len(lo), len(g_ord)), dtype=torch.float64, device="cuda")
        if len(small):
            out[torch.as_tensor(small, device="cuda")] = _ops.kdist(
                kd, lo[small], hi[small], g_ord, None if w is None else w, None if w is None else woff[small])
        g = _ops.to_dev(np.asarray(g_ord, dtype=np.float64))
        for ib in np.nonzero(n > cap)[0]:          # rare: very wide bins at very fine grids
            ks, order = torch.sort(kd[lo[ib]:hi[ib]])
            if w is None:
                gs = torch.arange(1, int(n[ib]) + 1, device="cuda", dtype=torch.float64) / float(n[ib])
            else:
                ws = _ops.to_dev(w[woff[ib]:woff[ib] + n[ib]])[order]
                gs = torch.cumsum(ws, 0) / ws.sum()
            j = torch.clamp(torch.searchsorted(gs, g, right=True) - 1, 0, int(n[ib]) - 2)
            slope = (ks[j + 1] - ks[j]) / (gs[j + 1] - gs[j])
            r = slope * (g - gs[j]) + ks[j]
            out[ib] = torch.where(g <= gs[0], ks[0], torch.where(g >= gs[-1], ks[-1], r))
        return out.cpu().numpy()


_BACKEND = DeviceBackend()


def set_backend(b):
    global _BACKEND
    old, _BACKEND = _BACKEND, b
    return old


def bin_ranges(wavecalc, vbinmin, vbinmax):
    """[lo, hi) per bin of the points with vbinmin <= wavecalc <= vbinmax (the reference's boolean mask, :3637)."""
    lo = np.searchsorted(wavecalc, np.asarray(vbinmin, dtype=np.float64), side="left")
    hi = np.searchsorted(wavecalc, np.asarray(vbinmax, dtype=np.float64), side="right")
    return lo.astype(np.int32), hi.astype(np.int32)


def k_distribution(kabs, wavecalc, vbinmin, vbinmax, g_ord, ils=None):
    """k[NBIN, NG] of one (p, T) point; see oracle.k_distribution for the reference's numpy."""
    return _BACKEND.k_distribution(kabs, wavecalc, vbinmin, vbinmax, g_ord, ils)


def make_calc_ktable_chunk(reference_fn):
    """``calc_ktable_chunk(iwaves, Spectroscopy, Spectroscopy_LBL, self_frac, Measurement)`` with the per-bin tail on
    the device.  The body follows archnemesis/Spectroscopy_0.py:3558-3667 step for step (same line selection, same
    grid, same calc_klbl_online call); `reference_fn` is kept for uninstall()."""

    def calc_ktable_chunk(iwaves, Spectroscopy, Spectroscopy_LBL, self_frac, Measurement):
        S, SL, M = Spectroscopy, Spectroscopy_LBL, Measurement
        iwaves = np.asarray(iwaves)
        iwavemin, iwavemax, nwave = iwaves[0], iwaves[-1], len(iwaves)

        def half_width(iw):
            if M is not None:
                return (M.VFIL[0:M.NFIL[iw], iw] - M.VCONV[iw, 0]).max()
            return (S.WAVE[1] - S.WAVE[0]) / 2.

        vchunkmin = S.WAVE[iwavemin] - half_width(iwavemin)
        vchunkmax = S.WAVE[iwavemax] + half_width(iwavemax)
        vchunkmean = np.mean(S.WAVE[iwaves])
        linedata, lineparams, ispace = SL.LINE_DATA[0], SL.LINE_DATA_PARAMS[0], SL.ISPACE
        um = int(ispace) == 1          # WaveUnitEnum.Wavelength_um
        if um:
            wnchunkmin, wnchunkmax = 1. / vchunkmax * 1.0e4, 1. / vchunkmin * 1.0e4
        else:
            wnchunkmin, wnchunkmax = vchunkmin, vchunkmax
        linedata.set_params(vmin=wnchunkmin - lineparams.wn_approx_window * 2.,
                            vmax=wnchunkmax + lineparams.wn_approx_window * 2., wave_unit=0).fetch_linedata()
        linedata.fetch_partition_fn()
        k_coefficients = np.zeros((nwave, S.NG, S.NP, S.NT))
        if len(linedata.combined_line_data.NU) == 0:
            return k_coefficients
        vbinmin = np.array([S.WAVE[iw] - half_width(iw) for iw in iwaves])
        vbinmax = np.array([S.WAVE[iw] + half_width(iw) for iw in iwaves])
        ils = None
        if M is not None:
            def ils(ib, wavesel):
                iw = iwaves[ib]
                return np.interp(wavesel - S.WAVE[iw], M.VFIL[0:M.NFIL[iw], iw] - M.VCONV[iw, 0], M.AFIL[0:M.NFIL[iw], iw])
        for ip in range(S.NP):
            for it in range(S.NT):
                pressx, tempx = S.PRESS[ip], S.TEMP[it]
                alpha_d = linedata.calculate_doppler_width(tempx, combined_output=True)
                gamma_l = linedata.calculate_lorentz_width(tempx, pressx, amb_frac=1. - self_frac, combined_output=True)
                hwhm_voigt = 0.5346 * gamma_l + np.sqrt(0.2166 * gamma_l ** 2. + alpha_d ** 2.)
                delwn_calc = np.min(hwhm_voigt) / 5.
                delv_calc = delwn_calc * (vchunkmean ** 2.) / 1.0e4 if um else delwn_calc
                ncalc = int((vchunkmax - vchunkmin) / delv_calc)
                wavecalc = np.linspace(vchunkmin, vchunkmax, ncalc)
                SL.NWAVE, SL.WAVE = ncalc, wavecalc
                kabs = SL.calc_klbl_online(1, [pressx], [tempx], amb_frac=1. - self_frac)[:, 0, 0]
                k_coefficients[:, :, ip, it] = k_distribution(kabs, wavecalc, vbinmin, vbinmax, S.G_ORD, ils)
        return k_coefficients

    calc_ktable_chunk.__doc__ = (reference_fn.__doc__ or "") + "\n(archnemesis_dist_b200: per-bin sort and quantiles on the device)"
    calc_ktable_chunk.b200_reference = reference_fn
    return calc_ktable_chunk


def install_ktable():
    """Rebind archnemesis.Spectroscopy_0.calc_ktable_chunk (what calc_ktable's workers call, :3338-3556)."""
    mod = sys.modules.get("archnemesis.Spectroscopy_0")
    if mod is None or "fn" in _INSTALLED:
        return
    _INSTALLED["fn"] = mod.calc_ktable_chunk
    mod.calc_ktable_chunk = make_calc_ktable_chunk(mod.calc_ktable_chunk)


def uninstall_ktable():
    mod = sys.modules.get("archnemesis.Spectroscopy_0")
    if mod is not None and "fn" in _INSTALLED:
        mod.calc_ktable_chunk = _INSTALLED.pop("fn")
