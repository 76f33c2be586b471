"""Config 4 (64 limb paths): time the radiance kernel and the projection separately, path space vs layer space."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from archnemesis_dist_b200 import engine, ops  # noqa: E402


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    nwave = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    cfg = dict(bench.CFG, nwave=nwave)
    c = bench.make_case(cfg)
    c4 = bench.limb_case(c, 64)
    tab = c["tab"]
    hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    M, Ml = bench.fold_M(c4), bench.fold_Mlay(c4)
    for mode, name in ((engine.TRANSMISSION, "transmission"), (engine.THERMAL, "thermal")):
        ev = bench.make_evaluation(c4, mode=mode)
        for lay in (False, True):
            s = hp.stage(ev, True, M, Mlay=Ml if lay else None)
            if lay and not s.layer_space:
                continue
            tau, dk = hp.gas_opacity(s)
            args = (s.mode, tau, dk, s.gas_slot, s.taucia, s.taudust, s.tauray, s.dtaucon, s.layinc, s.scale, s.nlayin,
                    s.emtemp, s.laypress, hp.wave_d, hp.delg_d, s.emissivity, s.xfac, None, None, None, None, s.ISPACE,
                    s.TSURF, s.NVMR, s.NPAR, True)
            kw = dict(layer_space=True) if lay else {}
            t_rad = timeit(lambda: ops.radiance(*args, **kw))
            _, dspec, _ = ops.radiance(*args, **kw)
            if getattr(s, "M_sparse", None) is not None:
                t_prj = timeit(lambda: ops.jacobian_project_sparse(dspec, s.M_sparse, shared=lay))
                name += " (sparse M)"
            else:
                t_prj = timeit(lambda: ops.jacobian_project(dspec, s.M, shared=lay))
            print("%-12s %-11s radiance %.3f ms  projection %.3f ms  dspec %.2f GB" %
                  (name, "layer space" if lay else "path space", t_rad, t_prj, dspec.numel() * 8 / 1e9))
            del dspec, tau, dk
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
