#!/usr/bin/env python
"""BASELINE config 4 on N GPUs (torchrun): 64 limb paths forward+Jacobian, sharded by geometry (every rank repeats
the gas opacity, evaluates NPATH/N paths) or by wavenumber (every rank holds NWAVE/N rows of the table), with
the final NCCL all-gather.  Device-timed, max over ranks.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/measure_config4_dist.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from archnemesis_dist_b200 import dist as adist, engine, plan, synthetic  # noqa: E402
from tools.measure_configs import fm_objects  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
NW, NX, NGEOM = 4000, 60, 64


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- geometry sharding: full table, this rank's paths --------------------------------------------------
hp, ev, M = fm_objects(NW, NX, ngeom=NGEOM, transmission=True)
lo, hi = adist.my_chunk(NGEOM, rank, world)
ev_r = engine.Evaluation(**{k: getattr(ev, k) for k in ev.__dataclass_fields__ if k != "h2d_bytes"})
ev_r.LAYINC, ev_r.SCALE, ev_r.NLAYIN = (np.ascontiguousarray(ev.LAYINC[:, lo:hi]), np.ascontiguousarray(ev.SCALE[:, lo:hi]),
                                        np.ascontiguousarray(ev.NLAYIN[lo:hi]))
ev_r.EMTEMP = np.ascontiguousarray(ev.EMTEMP[:, lo:hi])
s = hp.stage(ev_r, True, np.ascontiguousarray(M[lo:hi]))


def step_geom():
    spec, dx, _ = hp.run(s)
    block = torch.cat([spec.unsqueeze(2), dx], dim=2).movedim(1, 0).contiguous()     # [paths, NWAVE, 1+NX]
    return adist.all_gather_rows(block, NGEOM, dim=0)


ms_g = timed(step_geom)
hp.close()
del hp, s
torch.cuda.empty_cache()

# ---- wavenumber sharding: NWAVE/N rows of the table, every path ---------------------------------------------
c = synthetic.make_fm_case(nwave=NW, nx=NX, seed=7)
tab = c["tab"]
ws = adist.WavenumberShard(engine.HotPath, tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"], rank, world)
s2 = ws.hotpath.stage(ws.slice_evaluation(ev), True, M)


def step_wave():
    spec, dx, _ = ws.hotpath.run(s2)
    block = torch.cat([spec.unsqueeze(2), dx], dim=2)                                # [NWAVE/N, paths, 1+NX]
    return adist.all_gather_rows(block, NW, dim=0)


ms_w = timed(step_wave)
if rank == 0:
    print("config4 x%d GPUs, NGEOM=%d NWAVE=%d NX=%d: geometry sharding %.2f ms (%.0f geometry-spectra/s), "
          "wavenumber sharding %.2f ms (%.0f geometry-spectra/s)" % (world, NGEOM, NW, NX, ms_g, NGEOM * 1e3 / ms_g,
                                                                      ms_w, NGEOM * 1e3 / ms_w))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
