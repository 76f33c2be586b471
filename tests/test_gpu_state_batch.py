"""Perturbed atmosphere states as an extra dimension of one launch (engine.combine_evaluations): what the numerical
columns of jacobian_nemesis (archnemesis/ForwardModel_0.py:2184-2361) run through.  The merged evaluation must give,
per state, what the state gives alone."""
import numpy as np
import pytest

from tests.util import relerr, cpu

pytestmark = pytest.mark.gpu


def _evaluation(engine, c, **kw):
    a = dict(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"], NVMR=c["NVMR"],
             NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"], EMTEMP=c["EMTEMP"],
             LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"], xfac=c["xfac"],
             ISPACE=c["ISPACE"])
    a.update(kw)
    return engine.Evaluation(**a)


@pytest.mark.parametrize("mode", ["thermal", "transmission"])
def test_states_side_by_side_equal_states_alone(mode):
    from archnemesis_dist_b200 import engine, synthetic
    c = synthetic.make_fm_case(nwave=24, ng=20, ngas=6, nlay=40, npro=40, nx=8, nvmr=8, seed=3)
    tab = c["tab"]
    hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    rng = np.random.default_rng(1)
    evs = []
    for s in range(5):
        temp = c["temp"] * (1.0 + 0.05 * rng.uniform(-1, 1, len(c["temp"])))
        kw = dict(temp=temp, amount=c["amount"] * rng.uniform(0.8, 1.2, c["amount"].shape),
                  taucia=c["taucon"] * rng.uniform(0.5, 1.5), EMTEMP=temp[c["LAYINC"][:, 0]].reshape(-1, 1).copy())
        if mode == "transmission":
            kw.update(mode=engine.TRANSMISSION, EMTEMP=None)
        if s == 3:
            kw["taucia"] = None                     # a state without that continuum term: zeros in the merged array
        evs.append(_evaluation(engine, c, **kw))
    alone = [cpu(hp.cirsrad(e, False)) for e in evs]
    merged, cols = engine.combine_evaluations(evs)
    assert merged.amount.shape == (6, 5 * 40) and merged.LAYINC.shape[1] == 5 and cols == [(i, 1) for i in range(5)]
    got = cpu(hp.cirsrad(merged, False))
    for (c0, n), a in zip(cols, alone):
        assert relerr(got[:, c0:c0 + n], a) < 1e-13
    # evaluations that differ in a per-evaluation quantity are not merged
    with pytest.raises(ValueError):
        engine.combine_evaluations([evs[0], _evaluation(engine, c, TSURF=250.0)])
    hp.close()
