#!/usr/bin/env python
"""Time the fused gas-opacity kernel at config 2 (or NWAVE=argv[1]) in its rebin variants."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from archnemesis_dist_b200 import engine, ops, plan, synthetic  # noqa: E402
from tools.measure_kernels import timeit  # noqa: E402

nwave = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
c = synthetic.make_fm_case(nwave=nwave, seed=7)
tab = c["tab"]
hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
ev = engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                       NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"],
                       EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"],
                       TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"])
M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
s = hp.stage(ev, True, M)
for grad in (True, False):
    for name, kw in (("parallel", {}), ("sequential", dict(force_seq=True))):
        ms = timeit(lambda: ops.gas_opacity(hp.table, s.dplan, s.amount, hp.otab, want_grad=grad, **kw), reps=3, warm=1)
        print("gas_opacity grad=%d %-10s %8.3f ms" % (grad, name, ms))
