"""Time ansb200_lbl_table_opacity on a line-by-line-table sized case and report the achieved fraction of the HBM
roofline.  Algorithmic bytes per launch: 8*(NGAS+2) written per (wavenumber, layer) with gradients (8 without) plus
2 x 8*NGAS read per (wavenumber, distinct plane touched) -- ln K and, only for non-positive corners, K."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from archnemesis_dist_b200 import ops, plan  # noqa: E402


def main(nwave=200000, npg=20, ntg=15, ngas=4, nlay=100, reps=10):
    rng = np.random.default_rng(0)
    K = torch.exp(torch.empty((nwave, 1, npg, ntg, ngas), dtype=torch.float64, device="cuda").uniform_(-60.0, -40.0))
    tab = ops.Table(K)
    del K
    P = np.exp(np.linspace(np.log(1e-7), np.log(20.0), npg)).astype(np.float32)
    T = np.linspace(70.0, 400.0, ntg).astype(np.float32)
    press = np.exp(np.linspace(np.log(5.0), np.log(1e-6), nlay))
    temp = 110.0 + 60.0 * np.abs(np.linspace(-1.0, 1.0, nlay))
    amount = ops.to_dev(np.exp(rng.uniform(40.0, 55.0, size=(ngas, nlay))))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    for grad in (False, True):
        hp = plan.klbl_plan(P, T, press, temp, grad)
        dplan = ops.LblDevicePlan(hp)
        planes = len(set(hp["corner"].reshape(-1).tolist()))
        for _ in range(3):
            ops.lbl_table_opacity(tab, dplan, amount, grad)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.lbl_table_opacity(tab, dplan, amount, grad)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = nwave * nlay * 8 * ((ngas + 2) if grad else 1) + nwave * planes * 8 * ngas
        print("lbl_table_opacity grad=%d: NWAVE=%d NLAY=%d NGAS=%d planes=%d  %.3f ms  %.1f GB/s algorithmic (%.0f MB)"
              % (grad, nwave, nlay, ngas, planes, ms, nbytes / ms / 1e6, nbytes / 1e6), flush=True)
    print("peaks:", {k: v for k, v in peaks.items() if "hbm" in k.lower()})


if __name__ == "__main__":
    main()
