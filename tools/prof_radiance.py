#!/usr/bin/env python
"""One multi-geometry radiance launch for ncu: python tools/prof_radiance.py [NWAVE] [NGEOM]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.measure_configs import fm_objects  # noqa: E402
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 592
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hp, ev, M = fm_objects(nw, 60, ngeom=ng, transmission=ng > 1)
s = hp.stage(ev, True, M)
go = hp.gas_opacity(s)
for _ in range(2):
    hp.finish(s, go)
torch.cuda.synchronize()
