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
        # (4) the same on a line-by-line table (4-D K, NG = 1): 5 wavenumbers over 2 ranks
        cl = syn.make_fm_case(nwave=5, ng=1, ngas=2, nlay=8, nvmr=3, npro=8, nx=7, seed=6)
        tl = cl["tab"]
        K4, one = np.ascontiguousarray(tl["K"][:, 0]), np.array([1.0])
        evl = engine.Evaluation(press_atm=cl["press"], temp=cl["temp"], amount=cl["amount"], gas_slot=cl["gas_slot"],
                                NVMR=cl["NVMR"], NPAR=cl["NPAR"], LAYINC=cl["LAYINC"], SCALE=cl["SCALE"],
                                NLAYIN=cl["NLAYIN"], EMTEMP=cl["EMTEMP"], LAYPRESS=cl["LAYPRESS"], taucia=cl["taucon"],
                                dtaucon=cl["dtaucon"], TSURF=cl["TSURF"], EMISSIVITY=cl["EMISSIVITY"], xfac=cl["xfac"])
        Ml = plan.fold_projection(cl["xmap"], cl["LAYINC"], cl["NLAYIN"], cl["DTE"], cl["DAM"], cl["DCO"], cl["NVMR"],
                                  cl["NDUST"])
        full_l = cpu_engine.HotPath(K4, tl["PRESS"], tl["TEMP"], one, tl["WAVE"]).forward_jacobian(evl, Ml)
        wsl = adist.WavenumberShard(cpu_engine.HotPath, K4, tl["PRESS"], tl["TEMP"], one, tl["WAVE"])
        assert wsl.hotpath.lbl_table and wsl.hotpath.K.shape[0] == (3 if rank == 0 else 2)
        for a, b in zip(wsl.forward_jacobian(evl, Ml, to_tensor=t), full_l):
            assert a.shape == b.shape and np.abs(a.numpy() - b).max() <= 1e-13 * np.abs(b).max()
        # (4b) the same evaluation with its continuum terms as a device plan: every rank cuts its rows of the CIA tables
        #      and of the Rayleigh / aerosol spectra
        import dataclasses
        tables, cplan = syn.make_continuum(5, 8, 3, ndust=0, seed=2, temp=cl["temp"])
        evp = dataclasses.replace(evl, taucia=None, dtaucon=None, continuum=(tables, cplan))
        full_p = cpu_engine.HotPath(K4, tl["PRESS"], tl["TEMP"], one, tl["WAVE"]).forward_jacobian(evp, Ml)
        assert np.abs(full_p[0] - full_l[0]).max() > 0.0            # (other continuum numbers than case 4)
        for a, b in zip(wsl.forward_jacobian(evp, Ml, to_tensor=t), full_p):
            assert a.shape == b.shape and np.abs(a.numpy() - b).max() <= 1e-13 * np.abs(b).max()
        sl = wsl.slice_evaluation(evp).continuum
        assert sl[0].kw.shape[2] == wsl.hi - wsl.lo and sl[1]["ur"].shape[1] == wsl.hi - wsl.lo
        # (5) (p,T)-grid sharding of the line-by-line generation: 5 state points over 2 ranks (3 + 2)
        from oracle import oracle as orc
        wn = np.linspace(1000.0, 1001.0, 41)
        lines = syn.make_line_list(12, 1000.0, 1001.0, seed=3, pad=5.0)
        mix = np.array([0.1, 0.9])
        pts = [(200.0, 0.1, 1.0), (296.0, 1.0, 1.1), (150.0, 1e-3, 0.9), (250.0, 0.5, 1.0), (120.0, 1e-4, 2.0)]

        def absorb(chunk):
            return t(np.stack([orc.lbl_absorption(wn, lines, tc, pc, 296.0, 1.0, q, 1.0, 28.0, mix) for tc, pc, q in chunk]))
        grid = adist.pt_grid(absorb, pts)
        assert grid.shape == (5, 41) and np.array_equal(grid.numpy(), absorb(np.asarray(pts)).numpy())
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharding_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
