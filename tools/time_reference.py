#!/usr/bin/env python
"""Time the UNMODIFIED reference (archNEMESIS, numpy + numba, single-threaded) on the stages of the hot path, in the
build container where /root/reference is mounted (it cannot travel to the GPU box).  Config-2 shapes on a sample of
wavenumbers (every stage is independent per wavenumber), scaled to NWAVE=4000; numba is warmed by one discarded call.
Output is committed under profiles/ next to the C-port number that bench.py reports as cpu_baseline."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402
from archnemesis_dist_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402

ans = import_reference()
from archnemesis.ForwardModel_0 import k_overlapg, calc_thermal_emission_spectrumg, map2pro, map2xvec  # noqa: E402

NS = int(sys.argv[1]) if len(sys.argv) > 1 else 40          # sampled wavenumbers
c = syn.make_fm_case(nwave=NS, seed=7)
tab = c["tab"]
S = ans.Spectroscopy_0(ILBL=0)
for kk in ("K", "PRESS", "TEMP", "G_ORD", "DELG", "WAVE", "NWAVE", "NG", "NP", "NT", "NGAS"):
    setattr(S, kk, tab[kk])


def best(fn, reps=2):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


t_k, (kg, dkdT) = best(lambda: S.calc_kg(len(c["press"]), c["press"], c["temp"]))
t_o, (taug, dk) = best(lambda: k_overlapg(tab["DELG"], kg, dkdT, c["amount"]))
tl, tp, dtl = orc.assemble_opacity(taug, dk, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"], c["dtaucon"], c["LAYINC"],
                                   c["SCALE"])
emt, emp = c["EMTEMP"][:, 0], c["LAYPRESS"][c["LAYINC"][:, 0]]
t0_, d0_ = np.ascontiguousarray(tl[..., 0]), np.ascontiguousarray(dtl[..., 0])
t_t, (sg, dsg, dts) = best(lambda: calc_thermal_emission_spectrumg(0, tab["WAVE"], t0_, d0_, c["NVMR"], emt, emp, -1.0,
                                                                   np.zeros(NS)))
dspec = np.tensordot(dsg, tab["DELG"], axes=([1], [0]))[..., None]
inc = [i for i in range(c["NPAR"]) if np.mean(c["xmap"][:, i, :]) != 0.0]
t_m, _ = best(lambda: map2xvec(map2pro(dspec, NS, c["NVMR"], c["NDUST"], c["NPRO"], 1, c["NLAYIN"], c["LAYINC"], c["DTE"],
                                       c["DAM"], c["DCO"], INCPAR=inc), NS, c["NVMR"], c["NDUST"], c["NPRO"], 1, c["NX"],
                               c["xmap"]))
f = 4000.0 / NS
tot = (t_k + t_o + t_t + t_m) * f
print("# reference (archNEMESIS %s, numba %s, 1 core of the build container), %d sampled wavenumbers scaled to 4000" % (
    getattr(ans, "__version__", "?"), __import__("numba").__version__, NS))
for name, t in (("calc_kg", t_k), ("k_overlapg", t_o), ("calc_thermal_emission_spectrumg", t_t), ("map2pro+map2xvec", t_m)):
    print("%-34s %8.2f s  (%.3f s on the sample)" % (name, t * f, t))
print("%-34s %8.2f s  -> %.5f spectra/s on one core" % ("forward+Jacobian stages, total", tot, 1.0 / tot))
