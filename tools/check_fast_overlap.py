"""Fast overlap kernel (csrc/koverlap_fast.cu) against the general kernel and the oracle, with timings.

    python tools/check_fast_overlap.py [NWAVE] [--oracle] [--nograd]

Runs ansb200_gas_opacity on a config-2 slice twice -- through the default dispatch (fast kernel + work list)
and with the general kernel forced (ANSB200_OVERLAP=general in a child process) -- and compares tau / dk;
--oracle adds the CPU oracle (NWAVE <= 512 or so).  Prints CUDA-event times of both.
"""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(nwave, want_grad, out=None):
    import torch
    from archnemesis_dist_b200 import ops, plan, synthetic
    c = synthetic.make_fm_case(nwave=nwave, ng=20, ngas=6, nlay=100, npro=100, nx=60, nvmr=8, seed=7)
    tab = c["tab"]
    hp = plan.kinterp_plan(tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad)
    T = ops.Table(tab["K"])
    dp = ops.DevicePlan(hp, want_grad)
    otab = ops.OverlapTables(tab["DELG"])
    am = ops.to_dev(c["amount"])
    res = ops.gas_opacity(T, dp, am, otab, want_grad)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(10):
        ev0.record()
        res = ops.gas_opacity(T, dp, am, otab, want_grad)
        ev1.record()
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1))
    tau, dk = (res if want_grad else (res, None))
    tau = tau.cpu().numpy()
    dk = dk.cpu().numpy() if dk is not None else np.zeros(0)
    if out:
        np.savez(out, tau=tau, dk=dk, ms=np.array(times))
    return c, tau, dk, times


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    nwave = int(args[0]) if args else 400
    want_grad = "--nograd" not in sys.argv
    if "--child" in sys.argv:
        run(nwave, want_grad, out="/tmp/kf_general.npz")
        return
    env = dict(os.environ, ANSB200_OVERLAP="general")
    subprocess.check_call([sys.executable, os.path.abspath(__file__), str(nwave), "--child"] +
                          ([] if want_grad else ["--nograd"]), env=env)
    ref = np.load("/tmp/kf_general.npz")
    if "--stats" in sys.argv:
        os.environ.setdefault("ANSB200_OVERLAP", "stats")
    c, tau, dk, times = run(nwave, want_grad)

    def rel(a, b):
        m = np.maximum(np.abs(a), np.abs(b))
        m[m == 0] = 1.0
        return float((np.abs(a - b) / m).max()) if a.size else 0.0

    print("NWAVE %d grad %s" % (nwave, want_grad))
    print("general kernel ms: %s" % np.round(ref["ms"], 3))
    print("fast dispatch  ms: %s" % np.round(times, 3))
    print("fast vs general: tau rel %.3e" % rel(tau, ref["tau"]))
    if want_grad:
        for col in range(dk.shape[-1]):
            a, b = dk[..., col], ref["dk"][..., col]
            print("   dk col %d: rel-to-max %.3e   elementwise %.3e" % (col, np.abs(a - b).max() / max(np.abs(b).max(), 1e-300), rel(a, b)))
    bad = ~np.isfinite(tau)
    print("non-finite tau entries: %d (general: %d)" % (bad.sum(), (~np.isfinite(ref["tau"])).sum()))
    if "--oracle" in sys.argv:
        from oracle import oracle
        oracle.set_sort_mode(oracle.NUMBA_ORDER)
        tab = c["tab"]
        nt = oracle.max_threads()
        k, dkdT = oracle.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True, nthreads=nt)
        if want_grad:
            rt, rd = oracle.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT, nthreads=nt)
            print("fast vs oracle: tau rel %.3e" % rel(tau, rt))
            for col in range(rd.shape[-1]):
                a, b = dk[..., col], rd[..., col]
                print("   dk col %d: rel-to-max %.3e" % (col, np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)))
        else:
            rt = oracle.k_overlap(tab["DELG"], k, c["amount"], nthreads=nt)
            print("fast vs oracle: tau rel %.3e" % rel(tau, rt))


if __name__ == "__main__":
    main()
