#!/usr/bin/env python
"""Thermal emission + Jacobian over NGEOM limb paths (config-2 atmosphere): radiance kernel time."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.measure_configs import fm_objects, timeit  # noqa: E402
for ng in (16, 64):
    hp, ev, M = fm_objects(4000, 60, ngeom=ng, transmission=False)
    ev.EMTEMP = np.asarray(ev.EMTEMP, dtype=np.float64)
    s = hp.stage(ev, True, M)
    go = hp.gas_opacity(s)
    Ms, s.M = s.M, None
    ms = timeit(lambda: hp.finish(s, go))
    print("thermal limb NGEOM=%d: radiance %.2f ms" % (ng, ms))
    hp.close(); del hp, s, go; torch.cuda.empty_cache()
