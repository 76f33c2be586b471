#!/usr/bin/env python
"""Per-kernel timing at BASELINE config 2 with CUDA events (not under a profiler): time, algorithmic bytes,
achieved GB/s against MEASURED_PEAKS.json.  Output is committed under profiles/."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from archnemesis_dist_b200 import engine, ops, plan, synthetic, lbl  # noqa: E402


def timeit(fn, reps=5, warm=2):
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


def main():
    nwave = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    c = synthetic.make_fm_case(nwave=nwave, seed=7)
    tab = c["tab"]
    NW, NG, NP, NT, NGAS = tab["K"].shape
    NLAY, NPAR, NX = 100, c["NPAR"], c["NX"]
    hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    ev = engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                           NVMR=c["NVMR"], NPAR=NPAR, LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"],
                           EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"],
                           TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"])
    M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    s = hp.stage(ev, True, M)
    U = plan.planes_touched(s.plan_host, NT)
    plane = NW * NG * NGAS * 8
    rows = []

    def rec(name, ms, nbytes, note=""):
        gbs = nbytes / ms / 1e6
        rows.append((name, ms, nbytes / 1e6, gbs, gbs / peak, note))

    for grad in (False, True):
        sg = hp.stage(ev, grad, M)
        ms = timeit(lambda: ops.kinterp(hp.table, sg.dplan, grad))
        rec("kinterp%s" % ("_grad" if grad else ""), ms, U * plane + (1 + grad) * NW * NG * NLAY * NGAS * 8,
            "U=%d planes + k%s out" % (U, ",dkdT" if grad else ""))
    k, d = ops.kinterp(hp.table, s.dplan, True)
    ms = timeit(lambda: ops.koverlap(k, s.amount, hp.otab, dkdT=d))
    rec("koverlap_grad (unfused)", ms, 2 * NW * NG * NLAY * NGAS * 8 + NW * NG * NLAY * 8 * (2 + NGAS), "k,dkdT in; tau,dk out")
    ms = timeit(lambda: ops.koverlap(k, s.amount, hp.otab))
    rec("koverlap (unfused)", ms, NW * NG * NLAY * NGAS * 8 + NW * NG * NLAY * 8, "k in; tau out")
    ms = timeit(lambda: hp.gas_opacity(s))
    rec("gas_opacity_grad (fused)", ms, U * plane + NW * NG * NLAY * 8 * (2 + NGAS), "B_kio")
    s0 = hp.stage(ev, False)
    ms = timeit(lambda: hp.gas_opacity(s0))
    rec("gas_opacity (fused)", ms, U * plane + NW * NG * NLAY * 8, "B_kio no grad")
    tau, dk = hp.gas_opacity(s)

    def rad(grad):
        return ops.radiance(s.mode, tau, dk if grad else None, s.gas_slot, s.taucia, None, None, s.dtaucon if grad else None,
                            s.layinc, s.scale, s.nlayin, s.emtemp, s.laypress, hp.wave_d, hp.delg_d, s.emissivity, s.xfac,
                            None, None, None, None, s.ISPACE, s.TSURF, s.NVMR, s.NPAR, grad)
    ms = timeit(lambda: rad(True))
    rec("radiance_grad", ms, NW * NG * NLAY * 8 * (2 + NGAS) + NW * NLAY * 8 * (1 + NPAR) + NW * NPAR * NLAY * 8,
        "tau,dk,continuum in; dspec out")
    ms = timeit(lambda: rad(False))
    rec("radiance", ms, NW * NG * NLAY * 8 + NW * NLAY * 8, "tau,continuum in")
    spec, dspec, _ = rad(True)
    ms = timeit(lambda: ops.jacobian_project(dspec, s.M))
    rec("jacobian_project", ms, NW * NPAR * NLAY * 8 + NW * NX * 8, "%.2f GFLOP" % (2e-9 * NW * NPAR * NLAY * NX))
    ms = timeit(lambda: hp.run(s))
    rec("forward+jacobian (3 kernels)", ms, U * plane + (1 + NPAR) * NW * NLAY * 8 + NW * (1 + NX) * 8, "B_fwdjac")
    # line by line: 20k lines x 20k wavenumbers x 4 (p,T)
    wn = np.linspace(1000.0, 1040.0, 20001)
    lines = synthetic.make_line_list(20000, 1000.0, 1040.0, seed=0)
    pts = [(200.0, 0.1, 1.0), (250.0, 0.5, 1.0), (296.0, 1.0, 1.0), (150.0, 0.01, 1.0)]
    ms = timeit(lambda: lbl.lbl_absorption(wn, lines, pts, 296.0, 1.0, 1.0, 28.0, np.array([0.1, 0.9])), reps=2, warm=1)
    pairs = 4 * 20000 * 20001
    rows.append(("lbl_absorption 20k lines x 20k pts x 4 pT", ms, 0, 0, 0, "%.2f G line-point pairs/s" % (pairs / ms / 1e6)))
    print("# config: NWAVE=%d NG=%d NGAS=%d NLAY=%d NPAR=%d NX=%d table %dx%d, U=%d planes; HBM peak %.0f GB/s (measured)" % (
        NW, NG, NGAS, NLAY, NPAR, NX, NP, NT, U, peak))
    print("%-42s %10s %12s %10s %8s  %s" % ("kernel", "ms", "alg MB", "GB/s", "frac", "note"))
    for r in rows:
        print("%-42s %10.3f %12.1f %10.1f %8.4f  %s" % r)


if __name__ == "__main__":
    main()
