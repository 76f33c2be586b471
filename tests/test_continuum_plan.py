"""Continuum terms as a device plan (archnemesis_dist_b200.continuum): collision-induced absorption, Rayleigh and
aerosol opacities and their fold into dTAUCON.

CPU part (live reference): the plan built from the reference's objects, evaluated in numpy by the oracle, against the
dense arrays the reference's own routines produce (calc_tau_cia :4516-4788, calc_tau_rayleigh :4869-4937, calc_tau_dust
:4790-4867, calculate_layer_opacity :3938-3981) on the Jupiter CIRS deck -- plus a CO2 / N2 / H2 variant of the gas list
that switches the fixed CO2-CO2, N2-N2 and N2-H2 spectra on.  `make_golden()` stores plan + expected arrays for the
GPU test of ansb200_continuum (python -m tests.test_continuum_plan).
"""
import os
import sys
import tempfile

import numpy as np
import pytest

from tests.util import relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "continuum.npz")


def _prepared_forward_model(variant):
    """A drop-in forward model on the Jupiter deck, taken to the point where CIRSrad would run."""
    from oracle.ref_import import import_reference
    from oracle import make_golden as mg
    from archnemesis_dist_b200 import forward_model as fmod
    from tests import cpu_engine
    from copy import deepcopy
    ans = import_reference()
    deck = mg.build_jupiter_deck(os.path.join(tempfile.mkdtemp(prefix="ansb200_c_"), "deck"))
    objs = mg.load_jupiter(ans, deck)
    if variant == "co2_n2":
        A = objs["Atmosphere"]
        A.ID[0], A.ISO[0] = int(ans.enum.GasEnum.CO2), 0
        A.ID[1], A.ISO[1] = int(ans.enum.GasEnum.N2), 0
    if variant == "no_rayleigh":
        objs["Scatter"].IRAY = 0
    cls = fmod.make_forward_model_class(sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0)
    cls.b200_engine = cpu_engine
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        fm = mg.make_forward_model(ans, cls, objs, deck)
        M = fm.Measurement
        M.build_ils(IGEOM=0)
        wmin, wmax = M.calc_wave_range(apply_doppler=True, IGEOM=0)
        fm.SpectroscopyX = deepcopy(fm.Spectroscopy)
        fm.SpectroscopyX.read_tables(wavemin=wmin, wavemax=wmax)
        fm.select_Measurement(0, 0)
        for n in ("Atmosphere", "Scatter", "Stellar", "Surface", "Layer", "CIA", "Telluric"):
            setattr(fm, n + "X", deepcopy(getattr(fm, n)))
        if variant == "no_cia":
            fm.CIAX = None
        fm.ScatterX.SOL_ANG = fm.MeasurementX.SOL_ANG[0, 0]
        fm.ScatterX.EMISS_ANG = fm.MeasurementX.EMISS_ANG[0, 0]
        fm.ScatterX.AZI_ANG = fm.MeasurementX.AZI_ANG[0, 0]
        fm.subprofretg()
        fm.LayerX.DUST_UNITS_FLAG = fm.AtmosphereX.DUST_UNITS_FLAG
        fm.calc_pathg()
        if variant == "dusty":
            fm.LayerX.CONT[:, 0] = np.linspace(1.0e3, 5.0e5, fm.LayerX.NLAY)
            fm.ScatterX.KEXT[:, 0] = np.linspace(2.0e-9, 1.0e-9, fm.ScatterX.NWAVE)
            fm.ScatterX.KSCA[:, 0] = 0.3 * fm.ScatterX.KEXT[:, 0]
    finally:
        os.chdir(cwd)
    return fm


def _both(fm):
    from oracle import oracle as orc
    from tests import cpu_engine
    dense = fm._b200_continuum(True)                 # the reference's own routines + the fold of :3938-3981
    hp = cpu_engine.HotPath(fm.SpectroscopyX.K, fm.SpectroscopyX.PRESS, fm.SpectroscopyX.TEMP, fm.SpectroscopyX.DELG,
                            fm.SpectroscopyX.WAVE)
    cp = fm._b200_continuum_plan(hp)
    assert cp is not None
    return dense, cp, orc.continuum_eval(cp[0], cp[1], True)


@pytest.mark.reference
@pytest.mark.parametrize("variant", ["jupiter", "co2_n2", "no_cia", "no_rayleigh", "dusty"])
def test_plan_reproduces_the_reference_arrays(variant):
    fm = _prepared_forward_model(variant)
    (TAUCIA, TAUDUST, TAURAY, dTAUCON), (tables, plan), (pcia, pdust, pray, pdcon) = _both(fm)
    NW, NLAY = fm.SpectroscopyX.NWAVE, fm.LayerX.NLAY
    terms = [t[0] for t in tables.meta["terms"]]
    if variant == "no_cia":
        assert TAUCIA is None and pcia is None and terms == []
    else:
        assert np.abs(TAUCIA).max() > 0.0 and relerr(pcia, TAUCIA) < 1e-13
        assert ("co2" in terms and "n2n2" in terms and "n2h2" in terms) == (variant == "co2_n2")
    assert relerr(pray, TAURAY) < 1e-15 and (np.abs(TAURAY).max() > 0.0) == (variant != "no_rayleigh")
    assert relerr(pdust, TAUDUST) < 1e-15 and (np.abs(TAUDUST).max() > 0.0) == (variant == "dusty")
    if dTAUCON is None:
        assert not np.any(pdcon)
    else:
        assert pdcon.shape == dTAUCON.shape == (NW, fm.AtmosphereX.NVMR + 2 + fm.ScatterX.NDUST, NLAY)
        for k in range(dTAUCON.shape[1]):
            s = np.abs(dTAUCON[:, k, :]).max()
            assert np.abs(pdcon[:, k, :] - dTAUCON[:, k, :]).max() <= 1e-13 * s, k
    # what travels per evaluation: per-layer coefficients and two or three spectra -- not NWAVE x NPAR x NLAY
    from archnemesis_dist_b200 import continuum
    assert continuum.plan_bytes(plan) < 64 * 1024 + 4 * NW * 8


def make_golden():
    out = {}
    for variant in ("jupiter", "co2_n2", "dusty"):
        fm = _prepared_forward_model(variant)
        (TAUCIA, TAUDUST, TAURAY, dTAUCON), (tables, plan), _ = _both(fm)
        pre = variant + "_"
        out[pre + "kw"], out[pre + "nplanes"] = tables.kw, tables.nplanes
        for k, v in plan.items():
            out[pre + "plan_" + k] = np.asarray(v)
        out[pre + "TAUCIA"], out[pre + "TAUDUST"], out[pre + "TAURAY"], out[pre + "dTAUCON"] = TAUCIA, TAUDUST, TAURAY, dTAUCON
    np.savez_compressed(GOLD, **out)
    print("wrote", GOLD, os.path.getsize(GOLD) // 1024, "KiB")


if __name__ == "__main__":
    make_golden()
