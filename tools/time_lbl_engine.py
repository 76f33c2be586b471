#!/usr/bin/env python
"""Line-by-line-table mode through the engine: one nadir thermal forward+Jacobian evaluation on a table of NWAVE
monochromatic points (NG = 1), stage by stage with CUDA events, then end to end with the Gaussian ILS applied on the
device (forward_jacobian_conv: only [NCONV, 1+NX] returns to the host)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from archnemesis_dist_b200 import engine, plan, synthetic as syn  # noqa: E402


def main(nwave=100000, ngas=4, nlay=60, nx=60, nconv=400):
    c = syn.make_fm_case(nwave=nwave, ng=1, ngas=ngas, nlay=nlay, npro=nlay, nx=nx, nvmr=ngas + 1, seed=5, tsurf=150.0)
    tab = c["tab"]
    K4 = np.ascontiguousarray(tab["K"][:, 0])
    hp = engine.HotPath(K4, tab["PRESS"], tab["TEMP"], np.array([1.0]), tab["WAVE"])
    npath = int(os.environ.get("LIMB_PATHS", "0"))
    if npath:
        # solar-occultation shape (the reference's mars_solocc example: ILBL = 2, one path per tangent height)
        nlm = 2 * nlay
        layinc, scale, nlayin = np.zeros((nlm, npath), np.int32), np.zeros((nlm, npath)), np.zeros(npath, np.int32)
        for p in range(npath):
            t = (p * (nlay - 2)) // npath
            seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay))
            nlayin[p] = len(seq)
            layinc[:len(seq), p] = seq
            scale[:len(seq), p] = 1.0 + 20.0 / (1.0 + np.abs(np.array(seq) - t))
        c["LAYINC"], c["SCALE"], c["NLAYIN"], c["EMTEMP"] = layinc, scale, nlayin, c["temp"][layinc]
    ev = engine.Evaluation(mode=engine.TRANSMISSION if npath else engine.THERMAL,
                           press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                           NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"],
                           EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"],
                           TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"])
    M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    wave = tab["WAVE"]
    step = wave[1] - wave[0]
    vconv = np.linspace(wave[200], wave[-201], nconv)
    t0 = time.perf_counter()
    op = plan.lbl_conv_operator(wave, vconv, 40.0 * step, plan.ILS_GAUSSIAN)
    t_op = (time.perf_counter() - t0) * 1e3
    cop = hp.conv_operator(op)
    sync = torch.cuda.synchronize
    for _ in range(3):
        hp.forward_jacobian_conv(ev, M, cop, 2, 1.0)
    sync()
    ev_ = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    s = hp.stage(ev, True, M)
    sync()
    reps = 9
    samples = []
    for _ in range(reps):
        torch.cuda._sleep(30_000_000)          # keep the device busy while the host queues the whole evaluation
        ev_[0].record()
        go = hp.gas_opacity(s)
        ev_[1].record()
        out = hp.ops.radiance(s.mode, go[0], go[1], s.gas_slot, s.taucia, s.taudust, s.tauray, s.dtaucon, s.layinc, s.scale,
                              s.nlayin, s.emtemp, s.laypress, hp.wave_d, hp.delg_d, s.emissivity, s.xfac, s.solflux,
                              s.reflectance, s.sol_ang, s.emiss_ang, s.ISPACE, s.TSURF, s.NVMR, s.NPAR, True)
        ev_[2].record()
        dx = hp.ops.jacobian_project_sparse(out[1], s.M_sparse) if getattr(s, "M_sparse", None) is not None else \
            hp.ops.jacobian_project(out[1], s.M)
        ev_[3].record()
        block = torch.cat([out[0][:, :1], dx[:, 0, :]], dim=1)
        hp.ops.convolve(cop, block)
        ev_[4].record()
        sync()
        samples.append([ev_[i].elapsed_time(ev_[i + 1]) for i in range(4)])
    acc = np.median(np.array(samples), axis=0)
    print("LBL table NWAVE=%d NLAY=%d NGAS=%d NX=%d NCONV=%d %s(%d operator entries, built in %.0f ms on the host)"
          % (nwave, nlay, ngas, nx, nconv, ("%d limb paths, transmission " % npath) if npath else "", len(op["widx"]), t_op))
    for name, v in zip(("lbl_table_opacity (calc_klblg + gas sum)", "radiance + layer Jacobian", "projection",
                        "block assembly + lblconvg"), acc):
        print("  %-45s %8.3f ms (median of %d)" % (name, v, reps))
    if npath:
        hp.close()
        return          # the convolved end-to-end route is per geometry (one path); the stages above are the point here
    t0 = time.perf_counter()
    for _ in range(reps):
        o = hp.to_host(hp.forward_jacobian_conv(ev, M, cop, 2, 1.0))
    print("  %-45s %8.3f ms  (h2d %.1f MB, d2h %.2f MB)" % ("end to end, dense continuum arrays in / [NCONV,1+NX] out",
                                                             (time.perf_counter() - t0) * 1e3 / reps,
                                                             ev.h2d_bytes / 1e6, o.nbytes / 1e6))
    # the same evaluation with the continuum terms as a PLAN (what the drop-in hands over since round 2: the per-layer
    # coefficients and the resident cross-section planes; the dense [NWAVE, NPAR, NLAY] arrays are made on the device)
    cont = syn.make_continuum(nwave, nlay, c["NVMR"], c["NDUST"], seed=5, temp=c["temp"])
    ev2 = engine.Evaluation(mode=engine.THERMAL, press_atm=c["press"], temp=c["temp"], amount=c["amount"],
                            gas_slot=c["gas_slot"], NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"],
                            NLAYIN=c["NLAYIN"], EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], continuum=cont,
                            TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"])
    for _ in range(3):
        hp.to_host(hp.forward_jacobian_conv(ev2, M, cop, 2, 1.0))
    t0 = time.perf_counter()
    for _ in range(reps):
        o = hp.to_host(hp.forward_jacobian_conv(ev2, M, cop, 2, 1.0))
    print("  %-45s %8.3f ms  (h2d %.2f MB, d2h %.2f MB)" % ("end to end, continuum plan in / [NCONV,1+NX] out",
                                                             (time.perf_counter() - t0) * 1e3 / reps,
                                                             ev2.h2d_bytes / 1e6, o.nbytes / 1e6))
    hp.close()


if __name__ == "__main__":
    main()
