#!/usr/bin/env python
"""Randomised parity sweep of the overlap kernels against the CPU oracle: shapes, dynamic ranges that force
static orders / ties / the exact-sort fall-back, dead and negative gases, scrambled g-ordering, near-equal
gases (packed-key collisions), float32 and float64 quadrature weights.
   python tools/stress_overlap.py [ncases] [seed]
tests/test_gpu_stress.py runs a short sweep of the same cases with assertions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KINDS = ("plain", "dynamic range 1e44", "scrambled gas", "non-positive gas", "near-equal gases", "1e-290 gas")


def make_case(rng, case):
    """k, dkdT, amount, del_g of one random case (numpy), and its description."""
    from archnemesis_dist_b200 import synthetic
    from oracle import oracle as orc
    ng = int(rng.choice([4, 5, 8, 10, 16, 20, 20, 20, 22]))
    ngas = int(rng.integers(2, 9))
    nlay = int(rng.integers(3, 9))
    nwave = int(rng.integers(3, 8))
    c = synthetic.make_fm_case(nwave=nwave, ng=ng, ngas=ngas, nlay=nlay, npro=nlay, nx=4, nvmr=max(ngas, 2),
                               seed=int(rng.integers(1, 10**6)), zero_fraction=float(rng.choice([0.0, 0.0, 0.2])))
    tab = c["tab"]
    k, dkdT = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    kind = case % len(KINDS)
    if kind == 1:      # huge dynamic range between gases: static orders and exact ties
        f = 10.0 ** rng.uniform(-22, 22, size=ngas)
        k *= f
        dkdT *= f
    elif kind == 2:    # scrambled g order of one gas
        g = int(rng.integers(0, ngas))
        perm = rng.permutation(ng)
        k[..., g] = k[..., g][:, perm, :]
        dkdT[..., g] = dkdT[..., g][:, perm, :]
    elif kind == 3:    # a negative / zero gas in some cells (the linear branch of calc_k can give k <= 0)
        g = int(rng.integers(0, ngas))
        m = rng.uniform(size=k.shape[:1] + (1,) + k.shape[2:3]) < 0.4
        k[..., g] = np.where(m, -np.abs(k[..., g]) * rng.choice([0.0, 1.0]), k[..., g])
    elif kind == 4:    # nearly equal gases: many near-collisions of the packed sort keys
        k[..., 1:] = k[..., :1] * (1.0 + 1e-9 * rng.normal(size=k[..., 1:].shape))
    elif kind == 5:    # a gas whose opacities are denormal next to normal ones
        k[..., 0] *= 1e-290
        dkdT[..., 0] *= 1e-290
    dg = tab["DELG"] if case % 5 else tab["DELG"].astype(np.float64)
    desc = "NG=%2d NGAS=%d NLAY=%d NWAVE=%d %-20s delg=%s" % (ng, ngas, nlay, nwave, KINDS[kind], dg.dtype)
    return k, dkdT, c["amount"], dg, desc


def rel(a, b):
    m = np.maximum(np.abs(a), np.abs(b))
    m[m == 0] = 1.0
    return float((np.abs(a - b) / m).max())


def col(a, b):
    s = np.abs(b).max()
    return float(np.abs(a - b).max() / s) if s > 1e-280 else 0.0      # (a denormal column has no digits to compare)


def run_case(k, dkdT, amount, dg):
    """Errors of the parallel kernels against the oracle and bit-exactness of the sequential rebin."""
    from archnemesis_dist_b200 import ops
    from oracle import oracle as orc
    orc.set_sort_mode(orc.NUMBA_ORDER)
    otab = ops.OverlapTables(dg)
    kd, dd, am = ops.to_dev(k), ops.to_dev(dkdT), ops.to_dev(amount)
    rt, rd = orc.k_overlap(dg, k, amount, dkdT=dkdT)
    rt0 = orc.k_overlap(dg, k, amount)
    tau, dk = (x.cpu().numpy() for x in ops.koverlap(kd, am, otab, dkdT=dd))
    tau0 = ops.koverlap(kd, am, otab).cpu().numpy()
    ts, ds = (x.cpu().numpy() for x in ops.koverlap(kd, am, otab, dkdT=dd, force_seq=True))
    return dict(tau=rel(tau, rt), tau_nograd=rel(tau0, rt0),
                dk=max(col(dk[..., q], rd[..., q]) for q in range(rd.shape[-1])),
                seq_exact=bool(np.array_equal(ts, rt) and np.array_equal(ds, rd)),
                finite=bool(np.isfinite(tau).all() == np.isfinite(rt).all()))


if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
    worst = dict(tau=0.0, dk=0.0)
    for case in range(ncases):
        k, dkdT, amount, dg, desc = make_case(rng, case)
        r = run_case(k, dkdT, amount, dg)
        worst["tau"] = max(worst["tau"], r["tau"], r["tau_nograd"])
        worst["dk"] = max(worst["dk"], r["dk"])
        ok = r["tau"] < 2e-13 and r["tau_nograd"] < 2e-13 and r["dk"] < 1e-11 and r["seq_exact"] and r["finite"]
        print("case %2d %s  tau %.1e / %.1e  dk %.1e  seq bit-exact %s%s" % (
            case, desc, r["tau"], r["tau_nograd"], r["dk"], r["seq_exact"], "" if ok else "   <-- CHECK"))
    print("worst: tau %.2e dk %.2e" % (worst["tau"], worst["dk"]))
