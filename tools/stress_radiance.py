#!/usr/bin/env python
"""Randomised parity sweep of the radiance / projection kernels against the CPU oracle: thermal and transmission,
1..40 ragged limb / nadir paths (all three kernels: one path per CTA, staged multi-path, warp-per-path), with and
without gradients, surface, dust / Rayleigh terms, wavelength space.   python tools/stress_radiance.py [ncases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    m = np.maximum(np.abs(a), np.abs(b))
    m[m == 0] = 1.0
    return float((np.abs(a - b) / m).max())


def col(a, b):
    s = np.abs(b).max()
    return float(np.abs(a - b).max() / s) if s > 0 else float(np.abs(a).max())


def run_case(rng):
    import torch
    from archnemesis_dist_b200 import ops, plan, synthetic
    from oracle import oracle as orc
    ng = int(rng.choice([5, 10, 20]))
    ngas = int(rng.integers(1, 6))
    nlay = int(rng.integers(4, 24))
    nwave = int(rng.integers(2, 6))
    nvmr = ngas + int(rng.integers(0, 3))
    ndust = int(rng.integers(0, 2))
    thermal = bool(rng.integers(0, 2))
    grad = bool(rng.integers(0, 2))
    npath = int(rng.choice([1, 2, 3, 4, 7, 33]))
    tsurf = float(rng.choice([-1.0, 180.0])) if thermal else -1.0
    ispace = int(rng.integers(0, 2)) if thermal else 0
    c = synthetic.make_fm_case(nwave=nwave, ng=ng, ngas=ngas, nlay=nlay, npro=nlay, nx=5, nvmr=nvmr, ndust=ndust,
                               seed=int(rng.integers(1, 10**6)), tsurf=tsurf)
    tab = c["tab"]
    d = ops.to_dev
    if npath == 1 and rng.integers(0, 2):
        layinc, scale, nlayin = c["LAYINC"], c["SCALE"], c["NLAYIN"]          # nadir
    else:
        nlm = 2 * nlay
        layinc = np.zeros((nlm, npath), np.int32)
        scale = np.zeros((nlm, npath))
        nlayin = np.zeros(npath, np.int32)
        for p in range(npath):
            t = int(rng.integers(0, nlay - 1))
            seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay))
            nlayin[p] = len(seq)
            layinc[:len(seq), p] = seq
            scale[:len(seq), p] = rng.uniform(1.0, 15.0, len(seq))
    nlm = layinc.shape[0]
    emtemp = np.zeros((nlm, npath))
    for p in range(npath):
        emtemp[:nlayin[p], p] = c["temp"][layinc[:nlayin[p], p]]
    wave = tab["WAVE"] if ispace == 0 else 1.0e4 / tab["WAVE"]
    kr, dr = orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True)
    if ngas == 1:
        tau, dk = kr[..., 0] * c["amount"][0], np.stack([kr[..., 0], dr[..., 0] * c["amount"][0]], axis=-1)
    else:
        tau, dk = orc.k_overlap(tab["DELG"], kr, c["amount"], dkdT=dr)
    f = 10.0 ** rng.uniform(-4, -1)
    tau, dk = tau * f, dk * f
    taucia = c["taucon"]
    taudust = 10.0 ** rng.uniform(-6, -3, size=taucia.shape) if rng.integers(0, 2) else None
    tauray = 10.0 ** rng.uniform(-6, -3, size=taucia.shape) if rng.integers(0, 2) else None
    xfac = rng.uniform(0.5, 2.0, nwave)
    emis = np.full(nwave, 0.9) if tsurf > 0 else np.zeros(nwave)
    out = ops.radiance(ops.THERMAL if thermal else ops.TRANSMISSION, d(tau), d(dk) if grad else None,
                       d(c["gas_slot"], torch.int32), d(taucia), d(taudust), d(tauray), d(c["dtaucon"]) if grad else None,
                       d(layinc, torch.int32), d(scale), d(nlayin, torch.int32), d(emtemp) if thermal else None,
                       d(c["LAYPRESS"]) if thermal else None, d(wave) if thermal else None, d(tab["DELG"].astype(np.float64)),
                       d(emis) if thermal else None, d(xfac), None, None, None, None, ispace, tsurf, c["NVMR"], c["NPAR"], grad)
    tcon = taucia + (taudust if taudust is not None else 0.0) + (tauray if tauray is not None else 0.0)
    # (the reference adds the three continuum terms to the gas opacity one after the other; summing them first
    #  differs at rounding level only)
    tl, tp, dtl = orc.assemble_opacity(tau, dk if grad else None, c["gas_slot"], c["NVMR"], c["NPAR"], tcon, c["dtaucon"],
                                       layinc, scale)
    dg = tab["DELG"]
    z = np.zeros(nwave)
    if thermal:
        S, dS, dT = orc.thermal_paths(ispace, wave, tl, dtl, c["NVMR"], nlayin, emtemp, c["LAYPRESS"], layinc, tsurf, emis,
                                      xfac, z, z, np.full(npath, 100.0), np.full(npath, 100.0))
    else:
        S, dS = orc.transmission(tp, dtl, xfac)
        dT = np.zeros_like(S) if grad else None
    desc = "%s grad=%d NPATH=%2d NLAY=%2d NG=%2d NGAS=%d NVMR=%d NDUST=%d tsurf=%5.0f ispace=%d dust=%d ray=%d" % (
        "thermal     " if thermal else "transmission", grad, npath, nlay, ng, ngas, nvmr, ndust, tsurf, ispace,
        taudust is not None, tauray is not None)
    if not grad:
        return desc, rel(out.cpu().numpy(), orc.g_integrate(S, None, None, dg)), 0.0, 0.0
    s_ref, d_ref, t_ref = orc.g_integrate(S, dS, dT, dg)
    spec, dspec, dts = (x.cpu().numpy() for x in out)
    got = np.transpose(dspec, (0, 2, 3, 1))
    e_d = max(col(got[:, kk], d_ref[:, kk]) for kk in range(d_ref.shape[1]))
    # projection of the same gradients
    M = plan.fold_projection(c["xmap"], layinc, nlayin, c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    dx = ops.jacobian_project(out[1], d(M)).cpu().numpy()
    d2 = orc.map2pro(d_ref, nwave, c["NVMR"], c["NDUST"], c["NPRO"], npath, nlayin, layinc, c["DTE"], c["DAM"], c["DCO"],
                     INCPAR=orc.included_params(c["xmap"]))
    x_ref = orc.map2xvec(d2, c["xmap"])
    e_x = max(col(dx[:, :, ix], x_ref[:, :, ix]) for ix in range(x_ref.shape[2]))
    return desc, max(rel(spec, s_ref), rel(dts, t_ref) if thermal else 0.0), e_d, e_x


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
    worst = [0.0, 0.0, 0.0]
    for case in range(n):
        desc, e_s, e_d, e_x = run_case(rng)
        worst = [max(a, b) for a, b in zip(worst, (e_s, e_d, e_x))]
        ok = e_s < 1e-11 and e_d < 1e-10 and e_x < 1e-10
        print("case %2d %s  spec %.1e  dspec %.1e  dx %.1e%s" % (case, desc, e_s, e_d, e_x, "" if ok else "   <-- CHECK"))
    print("worst: spec %.2e dspec %.2e dx %.2e" % tuple(worst))
