#!/usr/bin/env python
"""Per-rank cases of bench.py (perturb_case) on one GPU: gas-opacity time and the fast kernel's hand-over statistics."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from archnemesis_dist_b200 import engine, ops  # noqa: E402
from tools.measure_kernels import timeit  # noqa: E402

cfg = dict(bench.CFG)
c0 = bench.make_case(cfg)
tab = c0["tab"]
hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
for rank in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    c = bench.perturb_case(c0, rank)
    s = hp.stage(bench.make_evaluation(c), True, bench.fold_M(c))
    ops.overlap_mode(0)
    ms = timeit(lambda: hp.gas_opacity(s), reps=5, warm=2)
    ops.overlap_mode(2)
    hp.gas_opacity(s)
    torch.cuda.synchronize()
    print("rank %d: gas opacity %.3f ms  %s" % (rank, ms, ops.overlap_stats()), flush=True)
    ops.overlap_mode(0)
