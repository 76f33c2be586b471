"""CPU, world_size 2 over gloo: the sharding + all-gather plumbing of archnemesis_dist_b200/dist.py
(column, geometry and wavenumber sharding) with the oracle-backed engine standing in for the device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from archnemesis_dist_b200 import dist as adist


def test_chunk_bounds_match_reference_split():
    for n, w in ((60, 8), (7, 3), (5, 8), (1000, 8)):
        b = adist.chunk_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        base, rem = divmod(n, w)
        assert [hi - lo for lo, hi in b] == [base + (1 if i < rem else 0) for i in range(w)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from archnemesis_dist_b200 import plan, synthetic as syn, engine
        from tests import cpu_engine
        c = syn.make_fm_case(nwave=3, ng=10, ngas=2, nlay=8, nvmr=3, npro=8, nx=7, seed=5)
        tab = c["tab"]
        hp = cpu_engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])

        def make_ev(scale):
            return engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"] * scale,
                                     gas_slot=c["gas_slot"], NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"],
                                     SCALE=c["SCALE"], NLAYIN=c["NLAYIN"], EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"],
                                     taucia=c["taucon"], dtaucon=c["dtaucon"], TSURF=c["TSURF"],
                                     EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"])
        M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        # (1) column sharding (NX=7 over 2 ranks: ragged 4 + 3)
        spec, dx, _ = adist.jacobian_columns(hp, make_ev(1.0), M, to_tensor=t)
        full = hp.forward_jacobian(make_ev(1.0), M)
        assert np.array_equal(dx.numpy(), full[1])
        # (2) geometry sharding (3 geometries over 2 ranks: ragged 2 + 1)
        scales = [1.0, 0.5, 2.0]

        def evaluate(i):
            s, d, _ = hp.forward_jacobian(make_ev(scales[i]), M)
            return t(s[:, 0]), t(d[:, 0, :])
        YN, KK = adist.geometries(evaluate, 3)
        for i in range(3):
            s, d, _ = hp.forward_jacobian(make_ev(scales[i]), M)
            assert np.array_equal(YN[i].numpy(), s[:, 0]) and np.array_equal(KK[i].numpy(), d[:, 0, :])
        # (3) wavenumber sharding (NWAVE=3 over 2 ranks: ragged 2 + 1), two paths
        ws = adist.WavenumberShard(cpu_engine.HotPath, tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
        got = ws.forward_jacobian(make_ev(1.0), M, to_tensor=t)
        for a, b in zip(got, full):
            # (the stand-in engine projects with einsum, whose blocking depends on the row count: compare to rounding)
            assert a.shape == b.shape and np.abs(a.numpy() - b).max() <= 1e-13 * np.abs(b).max()
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharding_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
