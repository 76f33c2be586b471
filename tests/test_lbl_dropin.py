"""Run-time line-by-line cross sections behind the reference API (archnemesis_dist_b200.linedata).

CPU part (needs the live reference): `install_lbl()` rebinds LineData_0.add_line_set_monochromatic_absorption and
LineSetSpecData.add_monochromatic_absorption; with the oracle as the backend the reference's own callers --
LineSetSpecData / LineData_0.add_monochromatic_absorption, Spectroscopy_0.calc_klbl_online, calc_klblg_online and
calc_lbltable_chunk (LineData_0.py:822-915, :2282-2461; Spectroscopy_0.py:1922-2143, :3308-3333) -- must return what
they return with the unmodified numba function, on hand-made line sets (the HITRAN files need h5py, absent here).

GPU part: the same replacement method on a stub line set with the device backend against tests/golden/lbl_dropin.npz,
written by this file's `make_golden()` from the live reference (python -m tests.test_lbl_dropin).
"""
import os
import sys
import types

import numpy as np
import pytest

from tests.util import relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lbl_dropin.npz")


class OracleBackend:
    """oracle.lbl_absorption behind the backend interface of linedata (tests only)."""

    calls = 0
    shapes = ("voigt",)

    def available(self):
        return True

    def absorption(self, wn_grid, lines, t, p, q, t_ref, p_ref, ab, mass, mix, s_floor, cw, aw, shape):
        from oracle import oracle
        OracleBackend.calls += 1
        return oracle.lbl_absorption(wn_grid, lines, t, p, t_ref, p_ref, q, ab, mass, mix, s_floor, cw, aw, shape)


def synthetic_line_set(n, lo, hi, seed, n_amb=1):
    rng = np.random.default_rng(seed)
    nu = np.sort(rng.uniform(lo, hi, n))
    return dict(nu=nu, sw=10.0 ** rng.uniform(-26.0, -20.0, n), a=rng.uniform(0.1, 10.0, n),
                elower=rng.uniform(0.0, 3000.0, n), gamma_self=rng.uniform(0.05, 0.1, n), n_self=rng.uniform(0.5, 0.8, n),
                gamma_amb=rng.uniform(0.03, 0.09, (n, n_amb)), n_amb=rng.uniform(0.5, 0.8, (n, n_amb)),
                delta_amb=rng.uniform(-0.01, 0.0, (n, n_amb)))


def partition(T):
    return 100.0 * (np.asarray(T, dtype=float) / 296.0) ** 1.5


def _reference_objects():
    from oracle.ref_import import import_reference
    ans = import_reference()
    ld = sys.modules["archnemesis.LineData_0"]
    from archnemesis.database.datatypes.line_set_data import LineSetData
    return ans, ld, LineSetData


def _line_set_spec(ans, ld, LineSetData, mol, iso, n, lo, hi, seed):
    a = synthetic_line_set(n, lo, hi, seed)
    lsd = LineSetData(1e-30, 296.0, 1.0, (lo, hi), np.full(n, mol), np.full(n, iso), a["nu"], a["sw"], a["a"], a["elower"],
                      a["gamma_self"], a["n_self"], a["gamma_amb"], a["n_amb"], a["delta_amb"])
    return ld.LineSetSpecData.create_from(mol, iso, (ans.enum.AmbientGasEnum.AIR,), lsd)


def _line_data(ans, ld, lss_list, mol, iso):
    """A LineData_0 with hand-made per-isotopologue line sets (no database behind it)."""
    obj = ld.LineData_0(ID=mol, ISO=iso, LINE_DATABASE="synthetic", cache=None)
    obj.line_data = list(lss_list)
    obj.continuum_data = [None] * len(lss_list)
    obj.partition_fn_data = [partition] * len(lss_list)
    return obj


def _spectroscopy(ans, ld, wave, gases):
    sp_mod = sys.modules["archnemesis.Spectroscopy_0"]
    S = ans.Spectroscopy_0(ILBL=ans.enum.SpectralCalculationModeEnum.LINE_BY_LINE_RUNTIME)
    S.ISPACE = ans.enum.WaveUnitEnum.Wavenumber_cm
    S.NGAS = len(gases)
    S.ID = np.array([g[0] for g in gases])
    S.ISO = np.array([g[1] for g in gases])
    S.WAVE, S.NWAVE = wave, len(wave)
    S.LINE_DATA = [g[2] for g in gases]
    S.LINE_DATA_PARAMS = [sp_mod.MolLineDataParams(include_continuum=False, use_cache=False, s_floor=1e-25,
                                                   include_pressure_shift=(i % 2 == 0)) for i in range(len(gases))]
    S.N_AMB_GASSES = 1
    return S


@pytest.mark.reference
def test_callers_of_the_numba_function_get_the_same_numbers():
    from archnemesis_dist_b200 import linedata
    ans, ld, LineSetData = _reference_objects()
    lss_a = _line_set_spec(ans, ld, LineSetData, 5, 1, 160, 1990.0, 2110.0, 3)
    lss_b = _line_set_spec(ans, ld, LineSetData, 2, 1, 90, 1985.0, 2115.0, 4)
    wave = np.linspace(2040.0, 2060.0, 401)
    press = np.array([0.3, 1e-3, 2.0])
    temp = np.array([210.0, 150.0, 296.0])

    def run_all():
        gases = [(5, 1, _line_data(ans, ld, [lss_a], 5, 1)), (2, 1, _line_data(ans, ld, [lss_b], 2, 1))]
        S = _spectroscopy(ans, ld, wave.copy(), gases)
        k = S.calc_klbl_online(3, press, temp, amb_frac=0.8)
        kg, dk = S.calc_klblg_online(3, press, temp, amb_frac=np.array([0.7]))
        one = lss_a.add_monochromatic_absorption(wave.copy(), ans.lineshape.voigt, 250.0, 0.05, partition,
                                                 np.array([0.1, 0.9]), 0.97, use_cache=False, wn_calc_range=(2000.0, 2100.0),
                                                 include_pressure_shift=False)
        # calc_lbltable_chunk works on a Spectroscopy with a (p, T) grid and one gas
        S1 = _spectroscopy(ans, ld, wave.copy(), gases[:1])
        S1.NP, S1.NT, S1.PRESS, S1.TEMP = 2, 2, np.array([1e-2, 1.0]), np.array([180.0, 260.0])
        chunk = sys.modules["archnemesis.Spectroscopy_0"].calc_lbltable_chunk(np.arange(50, 250), S1, 0.3)
        # a shape the device does not know goes to the reference either way
        o = np.zeros(len(wave))
        ld.add_line_set_monochromatic_absorption(wave, ans.lineshape.lorentz, 200.0, 296.0, 0.1, 1.0, 1.2, 1.0, 28.0,
                                                 np.array([0.5, 0.5]), lss_a._data[5:], *lss_a._data[:4], out=o)
        return k, kg, dk, one, chunk, o

    ref = run_all()
    old = linedata.set_backend(OracleBackend())
    OracleBackend.calls = 0
    linedata.install_lbl()
    try:
        assert ld.LineSetSpecData.add_monochromatic_absorption is not linedata._INSTALLED["method"]
        got = run_all()
    finally:
        linedata.uninstall_lbl()
        linedata.set_backend(old)
    assert OracleBackend.calls == 6 + 12 + 1 + 4   # every (gas, point, T / T+5) evaluation went through the backend
    assert ld.add_line_set_monochromatic_absorption is not None and not linedata._INSTALLED
    for r, g in zip(ref[:5], got[:5]):
        assert r.shape == g.shape and relerr(g, r) < 1e-12
    assert np.abs(ref[5]).max() > 0.0 and relerr(got[5], ref[5]) < 1e-12
    # (the oracle backend declares only the Voigt shape, so the Lorentz call above ran in the reference function)


@pytest.mark.reference
def test_unknown_shapes_and_result_cache_follow_the_reference():
    from archnemesis_dist_b200 import linedata
    ans, ld, LineSetData = _reference_objects()
    lss = _line_set_spec(ans, ld, LineSetData, 5, 1, 60, 1990.0, 2110.0, 8)
    wave = np.linspace(2045.0, 2055.0, 101)
    mix = np.array([0.2, 0.8])
    sub = ans.lineshape.gaussian                  # not among the backend's shapes: must run in the reference function
    ref = lss.add_monochromatic_absorption(wave.copy(), sub, 230.0, 0.5, partition, mix.copy(), use_cache=False)
    old = linedata.set_backend(OracleBackend())
    linedata.install_lbl()
    try:
        OracleBackend.calls = 0
        got = lss.add_monochromatic_absorption(wave.copy(), sub, 230.0, 0.5, partition, mix.copy(), use_cache=False)
        assert OracleBackend.calls == 0 and np.array_equal(got, ref)
        # result cache: the second identical call is served from the reference's cache object
        a = lss.add_monochromatic_absorption(wave.copy(), ans.lineshape.voigt, 230.0, 0.5, partition, mix.copy(), use_cache=True)
        n = OracleBackend.calls
        b = lss.add_monochromatic_absorption(wave.copy(), ans.lineshape.voigt, 230.0, 0.5, partition, mix.copy(), use_cache=True)
        assert OracleBackend.calls == n and np.array_equal(a, b)
        # an empty set adds nothing
        empty = types.SimpleNamespace(has_data=False)
        z = ld.LineSetSpecData.add_monochromatic_absorption(empty, wave, ans.lineshape.voigt, 230.0, 0.5, partition, mix)
        assert not z.any()
    finally:
        linedata.uninstall_lbl()
        linedata.set_backend(old)


def make_golden():
    """Golden for the GPU box: LineSetSpecData.add_monochromatic_absorption of the LIVE reference on a synthetic set."""
    ans, ld, LineSetData = _reference_objects()
    lss = _line_set_spec(ans, ld, LineSetData, 5, 1, 400, 1900.0, 2200.0, 21)
    wave = np.linspace(2000.0, 2100.0, 4001)
    mix = np.array([0.15, 0.85])
    cases = [(210.0, 0.3, None, True), (150.0, 1e-3, (1950.0, 2150.0), False), (296.0, 2.0, None, True)]
    outs = []
    for (t, p, rng_, shift) in cases:
        outs.append(lss.add_monochromatic_absorption(wave.copy(), ans.lineshape.voigt, t, p, partition, mix.copy(), 0.97,
                                                     use_cache=False, s_floor=1e-25, wn_calc_range=rng_,
                                                     include_pressure_shift=shift))
    np.savez_compressed(GOLD, data=np.array(lss._data), mass=lss._molecular_mass, wave=wave, mix=mix,
                        cases=np.array([(t, p, -1.0 if r is None else r[0], -1.0 if r is None else r[1], float(s))
                                        for (t, p, r, s) in cases]), out=np.array(outs))
    print("wrote", GOLD, os.path.getsize(GOLD) // 1024, "KiB")


@pytest.mark.gpu
def test_device_backend_matches_the_live_reference_golden():
    """The replacement method with the DEVICE backend and a resident line list on the golden inputs."""
    from archnemesis_dist_b200 import linedata
    g = np.load(GOLD)

    class Cache:
        def get(self, *a):
            return None

        def set(self, *a):
            return None

    data = np.array(g["data"])
    stub = types.SimpleNamespace(_data=data, _data_hash=hash(bytes(data)), has_data=True, t_ref=296.0, p_ref=1.0,
                                 _molecular_mass=float(g["mass"]), _result_cache=Cache(), cache_identity=lambda: ("stub",))
    voigt = object()
    sys.modules.setdefault("archnemesis.lineshape", types.SimpleNamespace(voigt=voigt))
    voigt = sys.modules["archnemesis.lineshape"].voigt
    method = linedata.make_line_set_method(lambda *a, **k: (_ for _ in ()).throw(AssertionError("reference called")))
    old = linedata.set_backend(None)
    try:
        for (t, p, lo, hi, shift), ref in zip(g["cases"], g["out"]):
            rng_ = None if lo < 0 else (float(lo), float(hi))
            for _ in range(2):                                  # second call: resident arrays
                got = method(stub, np.array(g["wave"]), voigt, float(t), float(p), partition, np.array(g["mix"]), 0.97,
                             use_cache=False, s_floor=1e-25, wn_calc_range=rng_, include_pressure_shift=bool(shift))
                assert relerr(got, ref) < 1e-10 and np.abs(got - ref).max() < 1e-11 * np.abs(ref).max()
        assert 1 <= len(method._b200_resident) <= 3
    finally:
        linedata.set_backend(old)


if __name__ == "__main__":
    make_golden()
