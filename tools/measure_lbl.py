#!/usr/bin/env python
"""Line-by-line kernel throughput: synthetic HITRAN-shaped list (scaled down)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from archnemesis_dist_b200 import lbl, synthetic  # noqa: E402

nlines = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nwave = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
npt = int(sys.argv[3]) if len(sys.argv) > 3 else 4
wn = np.linspace(1000.0, 1000.0 + 0.002 * (nwave - 1), nwave)
lines = synthetic.make_line_list(nlines, wn[0], wn[-1], seed=0)
press = np.exp(np.linspace(-15, 2, 20))
temps = np.linspace(70, 300, 15)
pts = [(float(temps[(3 * i) % 15]), float(press[(7 * i + 5) % 20]), 1.0) for i in range(npt)]
mix = np.array([0.1, 0.9])
out = lbl.lbl_absorption(wn, lines, pts, 296.0, 1.0, 1.0, 28.0, mix)
torch.cuda.synchronize()
times = []
for _ in range(int(os.environ.get("LBL_REPS", "3"))):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = lbl.lbl_absorption(wn, lines, pts, 296.0, 1.0, 1.0, 28.0, mix)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
ms = sorted(times)[len(times) // 2]
print("launches (ms):", " ".join("%.1f" % t for t in times))
# pairs inside the +-75 cm-1 window
span = wn[-1] - wn[0]
nu = lines["nu"]
lo = np.searchsorted(wn, nu - 75.0)
hi = np.searchsorted(wn, nu + 75.0)
pairs = float((hi - lo).sum()) * npt
print("lbl: %d lines x %d wavenumbers x %d (p,T): %.1f ms, %.3g line-point pairs in window, %.1f G pairs/s" % (
    nlines, nwave, npt, ms, pairs, pairs / ms / 1e6))
