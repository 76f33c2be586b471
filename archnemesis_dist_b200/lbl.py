"""Line-by-line absorption on the device: host face of ansb200_lbl_absorption, mirroring
add_line_set_monochromatic_absorption (archnemesis/LineData_0.py:279-358) batched over (p,T)."""
import numpy as np
import torch

from . import _lib
from .ops import _ptr, _require_cuda, _stream, to_dev

SHAPE_IDS = {"voigt": 0, "lorentz": 1, "gaussian": 2}


def resident_lines(lines):
    """Upload a line list once: dict of device tensors accepted by lbl_absorption in place of the host arrays
    (a 10^6-line list is 80 MB; generating a (p,T) grid launch by launch should not re-send it)."""
    _require_cuda()
    return {k: to_dev(lines[k]) for k in ("nu", "sw", "e_lower", "stim_ref", "broadening")}


def lbl_absorption(wn_grid, lines, pts, t_ref, p_ref, abundance, mass, mix, s_floor=0.0, wn_calc_window=25.0,
                   wn_approx_window=75.0, shape="voigt", out=None):
    """wn_grid[NWAVE] ascending; lines: dict(nu, sw, e_lower, stim_ref [N], broadening [3*M,N]) -- the rows of
    LineSetSpecData._data (LineData_0.py:681); pts: iterable of (t_calc, p_calc, q_ratio); mix[M].
    Returns (and accumulates into) out[NPT,NWAVE] on the device."""
    _require_cuda()
    wn = to_dev(wn_grid)
    nu, sw, el, st = (to_dev(lines[k]) for k in ("nu", "sw", "e_lower", "stim_ref"))
    br = to_dev(lines["broadening"])
    mixd = to_dev(mix)
    ptd = to_dev(np.asarray(pts, dtype=np.float64).reshape(-1, 3))
    NPT, NWAVE, N, M = ptd.shape[0], wn.numel(), nu.numel(), mixd.numel()
    if br.shape != (3 * M, N):
        raise ValueError("broadening must be [3*len(mix), N_lines]")
    if out is None:
        out = torch.zeros((NPT, NWAVE), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().ansb200_lbl_absorption(_ptr(wn), NWAVE, _ptr(nu), _ptr(sw), _ptr(el), _ptr(st), _ptr(br), N,
                                                  _ptr(mixd), M, _ptr(ptd), NPT, float(t_ref), float(p_ref),
                                                  float(abundance), float(mass), float(s_floor), float(wn_calc_window),
                                                  float(wn_approx_window), SHAPE_IDS[shape], _ptr(out), _stream()))
    return out
