#!/usr/bin/env python
"""One launch each of the multi-path radiance kernels (64 limb paths, NWAVE=592) for ncu."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.measure_configs import fm_objects  # noqa: E402
for trans in (True, False):
    hp, ev, M = fm_objects(592, 60, ngeom=64, transmission=trans)
    s = hp.stage(ev, True, M)
    go = hp.gas_opacity(s)
    for _ in range(2):
        hp.finish(s, go)
    torch.cuda.synchronize()
    hp.close()
