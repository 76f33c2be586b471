#!/usr/bin/env python
"""Where the extra time of a step goes at N > 1 (run under torch.distributed.run): the config-2 step of bench.py with
(a) kernels only, (b) + block assembly, (c) + all-gather as bench.py issues it (async, two buffer sets),
(d) + all-gather on the compute stream, (f) the all-gather of a step issued after the NEXT step's kernels (it runs
beside that step's radiance kernel instead of holding SMs while the persistent overlap kernel starts)."""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from archnemesis_dist_b200 import engine  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = dict(bench.CFG)
c0 = bench.make_case(cfg)
c = bench.perturb_case(c0, rank)
tab = c["tab"]
hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
ev = bench.make_evaluation(c)
M = bench.fold_M(c)
NW, NX = cfg["nwave"], cfg["nx"]
staged = hp.stage(ev, True, M)
blocks = [torch.empty((NW, NX + 1), dtype=torch.float64, device="cuda") for _ in range(2)]
gath = [torch.empty((world, NW, NX + 1), dtype=torch.float64, device="cuda") for _ in range(2)]
if world > 1:
    for _ in range(16):
        dist.all_gather_into_tensor(gath[0], blocks[0])
torch.cuda.synchronize()
pend = [None, None]

ready = [None, None]


def step(mode, i):
    if mode == "f":
        slot = i & 1
        if pend[slot] is not None:
            pend[slot].wait()            # the gather out of / into this buffer set, two steps ago
            pend[slot] = None
        spec, dx, _ = hp.run(staged)
        b = blocks[slot]
        b[:, 0] = spec[:, 0]
        b[:, 1:] = dx[:, 0, :]
        ready[slot] = True
        prev = slot ^ 1
        if ready[prev] and world > 1:
            pend[prev] = dist.all_gather_into_tensor(gath[prev], blocks[prev], async_op=True)
            ready[prev] = None
        return
    spec, dx, _ = hp.run(staged)
    if mode == "a":
        return
    slot = i & 1
    if mode == "c" and pend[slot] is not None:
        pend[slot].wait()
    b = blocks[slot]
    b[:, 0] = spec[:, 0]
    b[:, 1:] = dx[:, 0, :]
    if mode == "b" or world == 1:
        return
    if mode == "c":
        pend[slot] = dist.all_gather_into_tensor(gath[slot], b, async_op=True)
    else:
        dist.all_gather_into_tensor(gath[slot], b)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

for mode in ("a", "c", "d", "f", "f", "c"):
    for i in range(3):
        step(mode, i)
    for p in pend:
        if p is not None: p.wait()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for i in range(K):
        step(mode, i)
    for j in (0, 1):
        if mode == "f" and ready[j] and world > 1:
            pend[j] = dist.all_gather_into_tensor(gath[j], blocks[j], async_op=True)
            ready[j] = None
        if pend[j] is not None:
            pend[j].wait(); pend[j] = None
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device="cuda")
    if world > 1:
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
        ts = [float(x.item()) for x in tl]
    else:
        ts = [float(t.item())]
    if rank == 0:
        print("mode %s: ms per step per rank %s" % (mode, " ".join("%.3f" % x for x in ts)), flush=True)
if world > 1:
    dist.destroy_process_group()
