#!/usr/bin/env python
"""Where the end-to-end step spends its time (host phases vs device), config 2."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from archnemesis_dist_b200 import engine, plan  # noqa: E402

cfg = dict(bench.CFG)
c = bench.make_case(cfg)
tab = c["tab"]
hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
ev = engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                       NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"],
                       EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"],
                       mode=engine.THERMAL, ISPACE=c["ISPACE"], TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"])
M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
sync = torch.cuda.synchronize
for _ in range(3):
    hp.forward_jacobian(ev, M); sync()
T = {}
def lap(name, t0):
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
N = 5
for _ in range(N):
    sync(); t = time.perf_counter()
    s = hp.stage_opacity(ev, True); lap("host: stage_opacity (plan + small H2D)", t)
    t = time.perf_counter(); go = hp.gas_opacity(s); lap("host: launch gas_opacity (alloc + call)", t)
    t = time.perf_counter(); hp.stage_radiance(s, ev, M); lap("host: stage_radiance (pinned memcpy + H2D enqueue)", t)
    t = time.perf_counter(); sync(); lap("wait: gas_opacity kernel + copies", t)
    t = time.perf_counter(); out = hp.finish(s, go); sync(); lap("radiance + project (launch + run)", t)
    t = time.perf_counter(); h = out[1].cpu(); lap("D2H", t)
for k, v in T.items():
    print("%-55s %8.3f ms" % (k, v / N))
t0 = time.perf_counter()
for _ in range(N):
    spec, dx, _ = hp.forward_jacobian(ev, M); dx.cpu()
print("%-55s %8.3f ms" % ("forward_jacobian + D2H per call (wall)", (time.perf_counter() - t0) * 1e3 / N))
