"""k-table generation: the drop-in for Spectroscopy_0.calc_ktable_chunk (archnemesis/Spectroscopy_0.py:
3558-3667) against the unmodified reference on synthetic line data, without and with an instrument function, and the
golden vectors for the device kernel (tests/golden/kdist.npz; regenerate with ANSB200_REGEN_GOLDEN=1)."""
import os
import sys
import types

import numpy as np
import pytest

from tests.util import relerr
from tests.test_lbl_dropin import OracleBackend, _reference_objects, _line_set_spec, _line_data, _spectroscopy

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kdist.npz")


class OracleKdist:
    """oracle.k_distribution behind ktable's backend interface (tests only)."""
    calls = 0

    def k_distribution(self, kabs, wavecalc, vbinmin, vbinmax, g_ord, ils=None):
        from oracle import oracle
        OracleKdist.calls += 1
        return oracle.k_distribution(kabs, wavecalc, vbinmin, vbinmax, g_ord, ils)


def _case(ans, ld, LineSetData, with_ils):
    from numpy.polynomial.legendre import leggauss
    lss = _line_set_spec(ans, ld, LineSetData, 5, 1, 300, 1990.0, 2110.0, 3)
    obj = _line_data(ans, ld, [lss], 5, 1)
    obj._params_fetched_lines_last = True          # hand-made line sets: nothing to fetch from a database
    obj._params_fetched_partition_last = True
    obj.set_params = lambda **k: obj
    S_LBL = _spectroscopy(ans, ld, np.linspace(2040.0, 2060.0, 11), [(5, 1, obj)])
    S = ans.Spectroscopy_0(ILBL=ans.enum.SpectralCalculationModeEnum.K_TABLES)
    S.ISPACE = ans.enum.WaveUnitEnum.Wavenumber_cm
    S.WAVE, S.NWAVE, S.NG = np.linspace(2045.0, 2055.0, 21), 21, 10
    x, w = leggauss(10)
    S.G_ORD, S.DELG = 0.5 * (x + 1), 0.5 * w
    S.NP, S.NT, S.PRESS, S.TEMP = 2, 2, np.array([1e-2, 1.0]), np.array([180.0, 260.0])
    M = None
    if with_ils:
        # a triangular instrument function of half-width 0.6 cm-1 around every bin centre (bins then overlap)
        nf = 7
        M = types.SimpleNamespace(NFIL=np.full(21, nf), VCONV=S.WAVE.reshape(-1, 1).copy(),
                                  VFIL=S.WAVE[None, :] + np.linspace(-0.6, 0.6, nf)[:, None],
                                  AFIL=np.repeat((1.0 - np.abs(np.linspace(-1.0, 1.0, nf)))[:, None] + 0.05, 21, axis=1))
    return S, S_LBL, M


@pytest.mark.reference
@pytest.mark.parametrize("with_ils", [False, True])
def test_calc_ktable_chunk_dropin_matches_reference(with_ils):
    from archnemesis_dist_b200 import ktable, linedata
    ans, ld, LineSetData = _reference_objects()
    sp_mod = sys.modules["archnemesis.Spectroscopy_0"]
    iwaves = np.arange(3, 12)
    S, S_LBL, M = _case(ans, ld, LineSetData, with_ils)
    ref = sp_mod.calc_ktable_chunk(iwaves, S, S_LBL, 0.3, M)
    assert ref.shape == (9, 10, 2, 2) and np.all(np.diff(ref, axis=1) >= 0.0) and ref.max() > 0.0
    old_l, old_k = linedata.set_backend(OracleBackend()), ktable.set_backend(OracleKdist())
    OracleKdist.calls = 0
    linedata.install_lbl()
    ktable.install_ktable()
    try:
        assert sp_mod.calc_ktable_chunk is not ktable._INSTALLED["fn"]
        S2, S_LBL2, M2 = _case(ans, ld, LineSetData, with_ils)
        got = sp_mod.calc_ktable_chunk(iwaves, S2, S_LBL2, 0.3, M2)
    finally:
        ktable.uninstall_ktable()
        linedata.uninstall_lbl()
        linedata.set_backend(old_l)
        ktable.set_backend(old_k)
    assert OracleKdist.calls == 4 and not ktable._INSTALLED          # one launch per (p, T) point
    assert sp_mod.calc_ktable_chunk.__name__ == "calc_ktable_chunk" and not hasattr(sp_mod.calc_ktable_chunk, "b200_reference")
    assert got.shape == ref.shape and relerr(got, ref) < 1e-12


@pytest.mark.reference
def test_kdist_golden_is_current():
    """The golden vectors of the device kernel: the line-by-line spectrum of one (p, T) point of the case above and the
    k-distributions the REFERENCE's calc_ktable_chunk makes of it (captured at its calc_klbl_online call)."""
    ans, ld, LineSetData = _reference_objects()
    sp_mod = sys.modules["archnemesis.Spectroscopy_0"]
    out = {}
    for tag, with_ils in (("plain", False), ("ils", True)):
        S, S_LBL, M = _case(ans, ld, LineSetData, with_ils)
        S.NP, S.NT, S.PRESS, S.TEMP = 1, 1, np.array([1e-2]), np.array([180.0])
        cap = {}
        orig = S_LBL.calc_klbl_online

        def spy(*a, **k):
            r = orig(*a, **k)
            cap["kabs"], cap["wavecalc"] = r[:, 0, 0].copy(), np.array(S_LBL.WAVE)
            return r
        S_LBL.calc_klbl_online = spy
        iwaves = np.arange(3, 12)
        k = sp_mod.calc_ktable_chunk(iwaves, S, S_LBL, 0.3, M)[:, :, 0, 0]
        hw = 0.6 if with_ils else (S.WAVE[1] - S.WAVE[0]) / 2.
        out.update({tag + "_kabs": cap["kabs"], tag + "_wavecalc": cap["wavecalc"], tag + "_k": k,
                    tag + "_vbinmin": S.WAVE[iwaves] - hw, tag + "_vbinmax": S.WAVE[iwaves] + hw, "g_ord": S.G_ORD,
                    "centres": S.WAVE[iwaves]})
        if with_ils:
            out["vfil_rel"], out["afil"] = M.VFIL[:, 0] - M.VCONV[0, 0], M.AFIL[:, 0]
    if os.environ.get("ANSB200_REGEN_GOLDEN") == "1" or not os.path.exists(GOLD):
        np.savez_compressed(GOLD, **out)
    z = np.load(GOLD)
    for name, a in out.items():
        assert relerr(np.asarray(a, dtype=float), z[name]) < 1e-13, name
    # the oracle restatement on the captured spectrum reproduces the reference's k-distributions
    from oracle import oracle
    assert relerr(oracle.k_distribution(z["plain_kabs"], z["plain_wavecalc"], z["plain_vbinmin"], z["plain_vbinmax"],
                                        z["g_ord"]), z["plain_k"]) < 1e-13
    # (the golden keeps ONE relative filter grid, VFIL[:, 0] - VCONV[0]; the reference forms VFIL[:, iw] - VCONV[iw] per
    # bin, which differs from it by an ulp of the wavenumber: 1e-12 in the weights)
    ils = lambda ib, wv: np.interp(wv - z["centres"][ib], z["vfil_rel"], z["afil"])      # noqa: E731
    assert relerr(oracle.k_distribution(z["ils_kabs"], z["ils_wavecalc"], z["ils_vbinmin"], z["ils_vbinmax"], z["g_ord"],
                                        ils), z["ils_k"]) < 1e-11


def test_oracle_kdist_against_golden():
    """(runs without the reference) the oracle restatement against the committed vectors."""
    from oracle import oracle
    if not os.path.exists(GOLD):
        pytest.skip("golden not generated")
    z = np.load(GOLD)
    got = oracle.k_distribution(z["plain_kabs"], z["plain_wavecalc"], z["plain_vbinmin"], z["plain_vbinmax"], z["g_ord"])
    assert relerr(got, z["plain_k"]) < 1e-13
    from archnemesis_dist_b200 import ktable
    lo, hi = ktable.bin_ranges(z["plain_wavecalc"], z["plain_vbinmin"], z["plain_vbinmax"])
    for ib in range(len(lo)):
        m = (z["plain_wavecalc"] >= z["plain_vbinmin"][ib]) & (z["plain_wavecalc"] <= z["plain_vbinmax"][ib])
        assert np.array_equal(np.nonzero(m)[0], np.arange(lo[ib], hi[ib]))
