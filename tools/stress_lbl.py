#!/usr/bin/env python
"""Randomised parity sweep of the line-by-line kernel against the CPU oracle (SciPy Voigt): extreme pressures and
temperatures (pure Doppler to heavily pressure-broadened), narrow and wide grids, windows, strength floor, one or two
broadeners.   python tools/stress_lbl.py [ncases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(rng):
    from archnemesis_dist_b200 import lbl, synthetic
    from oracle import oracle as orc
    nwave = int(rng.integers(300, 2500))
    span = float(10.0 ** rng.uniform(-0.5, 2.0))              # 0.3 .. 100 cm-1
    wn0 = float(rng.choice([50.0, 800.0, 4000.0]))
    wn = np.linspace(wn0, wn0 + span, nwave)
    n_amb = int(rng.integers(1, 3))
    lines = synthetic.make_line_list(int(rng.integers(20, 300)), wn[0], wn[-1], seed=int(rng.integers(1, 10**6)),
                                     n_amb=n_amb, pad=float(rng.choice([5.0, 30.0, 80.0])))
    mix = rng.dirichlet(np.ones(1 + n_amb))
    calc_win = float(rng.choice([25.0, 5.0, 0.5]))
    approx_win = float(rng.choice([75.0, 75.0, 30.0]))
    if approx_win < calc_win:
        approx_win = calc_win
    s_floor = float(rng.choice([0.0, 0.0, 1e-24]))
    pts = [(float(rng.uniform(60, 900)), float(10.0 ** rng.uniform(-9, 2)), float(rng.uniform(0.3, 5.0))) for _ in range(3)]
    mass = float(rng.choice([2.0, 28.0, 200.0]))
    out = lbl.lbl_absorption(wn, lines, pts, t_ref=296.0, p_ref=1.0, abundance=0.97, mass=mass, mix=mix, s_floor=s_floor,
                             wn_calc_window=calc_win, wn_approx_window=approx_win).cpu().numpy()
    worst = 0.0
    for i, (t, p, q) in enumerate(pts):
        ref = orc.lbl_absorption(wn, lines, t, p, 296.0, 1.0, q, 0.97, mass, mix, s_floor=s_floor, wn_calc_window=calc_win,
                                 wn_approx_window=approx_win)
        m = np.maximum(np.abs(ref), np.abs(out[i]))
        m[m == 0] = 1.0
        worst = max(worst, float((np.abs(out[i] - ref) / m).max()))
    desc = "NWAVE=%4d span=%6.2f at %6.0f lines=%3d n_amb=%d windows=%4.1f/%4.1f floor=%g mass=%g p=%s" % (
        nwave, span, wn0, len(lines["nu"]), n_amb, calc_win, approx_win, s_floor, mass,
        ",".join("%.0e" % p for _, p, _ in pts))
    return desc, worst


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
    worst = 0.0
    for case in range(n):
        desc, e = run_case(rng)
        worst = max(worst, e)
        print("case %2d %s  relerr %.1e%s" % (case, desc, e, "" if e < 1e-10 else "   <-- CHECK"))
    print("worst %.2e" % worst)
