"""ansb200_continuum (csrc/continuum.cu): continuum opacities on the device from a host plan, against the dense arrays
of the LIVE reference stored in tests/golden/continuum.npz (tests/test_continuum_plan.py writes it) and, through the
engine, against the same evaluation with dense host arrays."""
import numpy as np
import pytest

from tests.util import relerr, colerr, cpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", ["jupiter", "co2_n2", "dusty"])
def test_device_continuum_matches_reference_arrays(variant):
    import os
    import torch
    from archnemesis_dist_b200 import ops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "continuum.npz"))
    pre = variant + "_"
    plan = {k[len(pre) + 5:]: g[k] for k in g.files if k.startswith(pre + "plan_")}
    dev = {}
    for k, v in plan.items():
        if v.ndim == 0:
            dev[k] = v.item()
        else:
            dev[k] = ops.to_dev(v, torch.int32 if v.dtype.kind == "i" else torch.float64)
    kw = ops.to_dev(g[pre + "kw"]) if g[pre + "kw"].size else None
    npl = ops.to_dev(g[pre + "nplanes"], torch.int32) if g[pre + "nplanes"].size else None
    NW = g[pre + "TAURAY"].shape[0]
    for want_grad in (True, False):
        taucia, taudust, tauray, dtaucon = ops.continuum(kw, npl, dev, NW, want_grad)
        assert relerr(cpu(taucia), g[pre + "TAUCIA"]) < 1e-13
        assert relerr(cpu(tauray), g[pre + "TAURAY"]) < 1e-15 and relerr(cpu(taudust), g[pre + "TAUDUST"]) < 1e-15
        if want_grad:
            ref = g[pre + "dTAUCON"]
            got = cpu(dtaucon)
            assert got.shape == ref.shape
            for k in range(ref.shape[1]):
                assert colerr(got[:, k, :], ref[:, k, :]) < 1e-13, k
        else:
            assert dtaucon is None


def test_engine_evaluation_with_a_plan_equals_dense_arrays():
    """HotPath.forward_jacobian with Evaluation.continuum = (tables, plan) against the same call with the dense arrays the
    oracle makes from that plan: equal spectra and Jacobians, and kilobytes instead of megabytes over PCIe."""
    from archnemesis_dist_b200 import engine, plan as b2plan, synthetic, continuum
    from oracle import oracle as orc
    c = synthetic.make_fm_case(nwave=48, ng=20, ngas=6, nlay=100, npro=100, nx=30, nvmr=8, seed=7)
    tab = c["tab"]
    hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    tables, cplan = synthetic.make_continuum(48, 100, 8, ndust=0, seed=3, temp=c["temp"])
    taucia, taudust, tauray, dtaucon = orc.continuum_eval(tables, cplan, True)
    assert 1e-8 < np.median(taucia) < 1.0 and np.abs(dtaucon).max() > 0.0
    M = b2plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    common = dict(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"], NVMR=c["NVMR"],
                  NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"], EMTEMP=c["EMTEMP"],
                  LAYPRESS=c["LAYPRESS"], TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"], ISPACE=c["ISPACE"])
    dense = engine.Evaluation(taucia=taucia, taudust=taudust, tauray=tauray, dtaucon=dtaucon, **common)
    planned = engine.Evaluation(continuum=(tables, cplan), **common)
    s1, x1, _ = hp.forward_jacobian(dense, M)
    s2, x2, _ = hp.forward_jacobian(planned, M)
    assert relerr(cpu(s2), cpu(s1)) < 1e-13
    for ix in range(x1.shape[-1]):
        assert colerr(cpu(x2)[..., ix], cpu(x1)[..., ix]) < 1e-12, ix
    # (both carry the 240 KB projection matrix; the dense continuum arrays are 0.5 MB at this size, the plan 40 KB)
    assert dense.h2d_bytes - planned.h2d_bytes > 400 * 1024
    assert continuum.plan_bytes(cplan) < 100 * 1024
    s0 = hp.cirsrad(engine.Evaluation(continuum=(tables, cplan), **common), False)
    d0 = hp.cirsrad(engine.Evaluation(taucia=taucia, taudust=taudust, tauray=tauray, **common), False)
    assert relerr(cpu(s0), cpu(d0)) < 1e-13
    hp.close()
