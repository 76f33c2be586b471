#!/usr/bin/env python
"""Stand-alone k-interp launches (no-grad, grad) at config 2 for ncu."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from archnemesis_dist_b200 import ops, plan, synthetic  # noqa: E402
c = synthetic.make_fm_case(nwave=4000, seed=7); tab = c["tab"]
T = ops.Table(tab["K"])
for grad in (False, True):
    dp = ops.DevicePlan(plan.kinterp_plan(tab["PRESS"], tab["TEMP"], c["press"], c["temp"], grad), grad)
    for _ in range(2):
        ops.kinterp(T, dp, grad)
torch.cuda.synchronize()
