#!/usr/bin/env python
"""ansb200_convolve on a line-by-line sized block: Gaussian ILS rows of ~150 calculation points, 61 columns."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from archnemesis_dist_b200 import ops, plan  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main(nwave=100000, ncol=61, nconv=400, width=40.0):
    wave = np.linspace(100.0, 100.0 + 0.25 * (nwave - 1), nwave)
    vconv = np.linspace(wave[200], wave[-201], nconv)
    op = plan.lbl_conv_operator(wave, vconv, width * 0.25, plan.ILS_GAUSSIAN)
    cop = ops.ConvOperator(op)
    spec = torch.rand((nwave, 1), dtype=torch.float64, device="cuda")
    dx = torch.rand((nwave, 1, ncol - 1), dtype=torch.float64, device="cuda")
    block = torch.cat([spec[:, :1], dx[:, 0, :]], dim=1)
    print("entries %d, rows of ~%d; block %s" % (len(op["widx"]), len(op["widx"]) // nconv, tuple(block.shape)))
    print("torch.cat   %.3f ms" % timed(lambda: torch.cat([spec[:, :1], dx[:, 0, :]], dim=1)))
    print("convolve    %.3f ms" % timed(lambda: ops.convolve(cop, block)))
    op0 = plan.lbl_conv_operator(wave, vconv, 0.0)
    cop0 = ops.ConvOperator(op0)
    print("np.interp   %.3f ms" % timed(lambda: ops.convolve(cop0, block)))


if __name__ == "__main__":
    main()
