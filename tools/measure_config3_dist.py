#!/usr/bin/env python
"""BASELINE config 3 on N GPUs (torchrun): line-by-line Voigt cross-sections, 10^6 lines x 10^5 wavenumbers, the
(p,T) grid sharded across ranks (dist.pt_grid) with the final NCCL all-gather of k[NPT, NWAVE].  Weak scaling:
PTS_PER_RANK state points per rank.  Device-timed, max over ranks.
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tools/measure_config3_dist.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from archnemesis_dist_b200 import dist as adist, lbl, synthetic  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
NLINES, NWAVE, PTS_PER_RANK = int(os.environ.get("NLINES", "1000000")), 100000, int(os.environ.get("PTS_PER_RANK", "6"))
wn = np.linspace(1000.0, 1000.0 + 0.002 * (NWAVE - 1), NWAVE)
lines = synthetic.make_line_list(NLINES, wn[0], wn[-1], seed=0)
press = np.exp(np.linspace(-15, 2, 20))
temps = np.linspace(70, 300, 15)
npt = PTS_PER_RANK * world
pts = [(float(temps[(3 * i) % 15]), float(press[(7 * i + 5) % 20]), 1.0) for i in range(npt)]
mix = np.array([0.1, 0.9])


dev_lines = lbl.resident_lines(lines)
wn_d = torch.from_numpy(wn).cuda()


def step():
    return adist.pt_grid(lambda chunk: lbl.lbl_absorption(wn_d, dev_lines, chunk, 296.0, 1.0, 1.0, 28.0, mix), pts)


step()
torch.cuda.synchronize()
times = []
for _ in range(3):
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = step()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
ms = sorted(times)[1]
nu = lines["nu"]
win = float((np.searchsorted(wn, nu + 75.0) - np.searchsorted(wn, nu - 75.0)).sum()) * npt
if rank == 0:
    print("config3 on %d GPU(s): %d lines x %d wavenumbers x %d (p,T) points (%d per rank): %.1f ms (max over ranks, incl. all-gather "
          "of %.0f MB) = %.2f ms per point, %.0f G line-point pairs/s, %.1f s for the 20x15 grid"
          % (world, NLINES, NWAVE, npt, PTS_PER_RANK, ms, out.numel() * 8 / 1e6, ms / npt, win / ms / 1e6, ms / npt * 300 / 1e3), flush=True)
if world > 1:
    dist.destroy_process_group()
