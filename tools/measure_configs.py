#!/usr/bin/env python
"""BASELINE.json configs 3, 4 and 5 on one B200 (config 2 is bench.py).  Output is committed under profiles/.

  config 3  line-by-line Voigt cross-sections: 1M lines x 100k wavenumbers, a sample of the 20x15 (p,T) grid
            (every (p,T) point is an independent slice of the same launch; the full grid is 300/NPT x the time)
  config 4  multi-geometry limb/solar-occultation forward+Jacobian: config-2 atmosphere, NGEOM tangent paths
            (transmission mode, NPATH = NGEOM) -- opacity once, radiance + projection for all paths
  config 5  Jacobian-column sweep: NX in {10, 60, 300, 1000} x NWAVE in {1k, 4k, 16k}
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from archnemesis_dist_b200 import engine, lbl, plan, synthetic  # noqa: E402


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def config3(nlines=1000000, nwave=100000, npt=6):
    wn = np.linspace(1000.0, 1000.0 + 0.002 * (nwave - 1), nwave)
    lines = synthetic.make_line_list(nlines, wn[0], wn[-1], seed=0)
    press = np.exp(np.linspace(-15, 2, 20))
    temps = np.linspace(70, 300, 15)
    pts = [(float(temps[(3 * i) % 15]), float(press[(7 * i + 5) % 20]), 1.0) for i in range(npt)]
    mix = np.array([0.1, 0.9])
    ms = timeit(lambda: lbl.lbl_absorption(wn, lines, pts, 296.0, 1.0, 1.0, 28.0, mix), reps=1, warm=1)
    nu = lines["nu"]
    core = float((np.searchsorted(wn, nu + 25.0) - np.searchsorted(wn, nu - 25.0)).sum()) * npt
    win = float((np.searchsorted(wn, nu + 75.0) - np.searchsorted(wn, nu - 75.0)).sum()) * npt
    print("config3 lbl: %d lines x %d wavenumbers x %d of 300 (p,T) points: %.1f ms (%.1f ms per (p,T) point, "
          "%.1f s for the 20x15 grid); %.3g line-point pairs in the +-75 cm-1 window (%.3g Voigt, rest 1/dnu^2 wings): "
          "%.1f G pairs/s" % (nlines, nwave, npt, ms, ms / npt, ms / npt * 300 / 1e3, win, core, win / ms / 1e6))


def fm_objects(nwave, nx, ngeom=1, transmission=False, seed=7):
    c = synthetic.make_fm_case(nwave=nwave, nx=nx, seed=seed)
    tab = c["tab"]
    hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    nlay = len(c["press"])
    if ngeom > 1:
        # limb / occultation paths: geometry g sees the layers above tangent layer t_g twice (in and out)
        nlaymax = 2 * nlay
        layinc = np.zeros((nlaymax, ngeom), np.int32)
        scale = np.zeros((nlaymax, ngeom))
        nlayin = np.zeros(ngeom, np.int32)
        for g in range(ngeom):
            t = (g * (nlay - 2)) // ngeom
            seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay))
            nlayin[g] = len(seq)
            layinc[:len(seq), g] = seq
            scale[:len(seq), g] = 1.0 + 20.0 / (1.0 + np.abs(np.array(seq) - t))
        c["LAYINC"], c["SCALE"], c["NLAYIN"] = layinc, scale, nlayin
        c["EMTEMP"] = c["temp"][layinc]
    ev = engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                           NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"],
                           EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"],
                           mode=engine.TRANSMISSION if transmission else engine.THERMAL, ISPACE=c["ISPACE"],
                           TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"])
    M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    return hp, ev, M


def config4(ngeoms=(16, 32, 64)):
    for ng in ngeoms:
        hp, ev, M = fm_objects(4000, 60, ngeom=ng, transmission=True)
        s = hp.stage(ev, True, M)
        ms = timeit(lambda: hp.run(s))
        go = hp.gas_opacity(s)
        ms_rad = timeit(lambda: hp.finish(s, go))
        M_saved, s.M = s.M, None
        ms_r = timeit(lambda: hp.finish(s, go))          # radiance only (layer-space gradients)
        s.M = M_saved
        print("config4 limb/SO transmission forward+Jacobian: NGEOM=%2d paths (NLAYIN up to %d), NWAVE=4000 NX=60: "
              "%.2f ms per evaluation of all geometries (gas opacity %.2f + radiance %.2f + projection %.2f) = %.0f "
              "geometry-spectra/s" % (ng, int(ev.NLAYIN.max()), ms, ms - ms_rad, ms_r, ms_rad - ms_r, ng * 1e3 / ms))
        hp.close()
        del hp, s, go
        torch.cuda.empty_cache()


def config5(nwaves=(1000, 4000, 16000), nxs=(10, 60, 300, 1000)):
    print("config5 sweep (ms per forward+Jacobian evaluation, device-resident inputs; Jacobian columns/s = NX*1e3/ms)")
    print("%8s " % "NWAVE" + " ".join("%22s" % ("NX=%d" % nx) for nx in nxs))
    for nw in nwaves:
        row = []
        for i, nx in enumerate(nxs):
            hp, ev, M = fm_objects(nw, nx, seed=1000 + i)
            s = hp.stage(ev, True, M)
            ms = timeit(lambda: hp.run(s))
            row.append("%8.2f ms %9.0f col/s" % (ms, nx * 1e3 / ms))
            hp.close()
            del hp, s
            torch.cuda.empty_cache()
        print("%8d " % nw + " ".join(row))


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "4", "5"]
    t0 = time.time()
    if "5" in which:
        config5()
    if "4" in which:
        config4()
    if "3" in which:
        config3()
    print("# wall %.0f s" % (time.time() - t0))
