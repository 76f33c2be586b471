"""GPU + live reference: forward_model.install() with the CUDA engine (engine.HotPath, no substitution) against the
unmodified reference classes on the reference's own Jupiter CIRS nadir deck -- the seam the CPU suite can only
exercise with the oracle-backed engine of tests/cpu_engine.py.  On the GPU box the reference is the byte-for-byte
mirror staged by oracle/make_ref.py under oracle/_ref (see oracle/ref_import.py); the FP64 bar is north_star's 1e-9."""
import os
import sys
import tempfile

import numpy as np
import pytest

from tests.util import relerr, colerr

pytestmark = [pytest.mark.gpu, pytest.mark.reference]

TOL = 1e-9


@pytest.fixture(scope="module")
def jupiter():
    from oracle.ref_import import import_reference
    from oracle import make_golden as mg
    ans = import_reference()
    deck = mg.build_jupiter_deck(os.path.join(tempfile.mkdtemp(prefix="ansb200_g_"), "deck"))
    return ans, deck, mg


def _columns_close(dS, dS_ref, tol, floor=0.0):
    """Every Jacobian column within `tol` of its own largest entry.  `floor` (a fraction of the largest entry of the
    WHOLE Jacobian) bounds the absolute error instead for columns that are themselves a rounding-level residue: in the
    thermal limb case the top temperature levels have columns 1e-7 ... 1e-10 of the others, each entry the difference
    T_j B_j - sum_m (T_m-1 - T_m) B_m of terms that nearly cancel along an opaque path, so two FP64 evaluations in a
    different order (the reference's recurrence, the closed form here, the per-path and the per-layer kernels) agree
    to ~1e-16 of the terms, not of the residue."""
    scale = np.abs(dS_ref).max()
    for ix in range(dS_ref.shape[-1]):
        cmax = np.abs(dS_ref[..., ix]).max()
        if cmax == 0.0:
            assert np.abs(dS[..., ix]).max() == 0.0, ix
            continue
        assert np.abs(dS[..., ix] - dS_ref[..., ix]).max() <= max(tol * cmax, floor * scale), ix


def test_install_runs_the_cuda_engine_on_the_reference_deck(jupiter):
    ans, deck, mg = jupiter
    from archnemesis_dist_b200 import engine, forward_model as fmod
    ref_cls = sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0
    objs = mg.load_jupiter(ans, deck)
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        ref = mg.make_forward_model(ans, ref_cls, objs, deck)
        S_ref, dS_ref = ref.nemesisfmg()
        S0_ref = ref.nemesisfm()
        tg_ref, dtg_ref = ref.calculate_gaseous_line_opacity(True)
        cls = fmod.install(ans)
        try:
            assert cls.b200_engine is engine                                 # nothing substituted
            fm = mg.make_forward_model(ans, ans.ForwardModel_0, objs, deck)
            S, dS = fm.nemesisfmg()
            hp = fm._b200_hotpath()
            assert isinstance(hp, engine.HotPath) and hp.table.nbytes > 0    # the resident table of the CUDA engine
            S0 = fm.nemesisfm()
            spec, dspec, dts = fm.CIRSrad(return_grad=True)
            assert isinstance(dspec, fmod.DeviceGradient)
            S_lazy, dS_lazy = ref_cls.nemesisfmg(fm)      # the reference's own driver body over the device gradient
            tg, dtg = fm.calculate_gaseous_line_opacity(True)
        finally:
            fmod.uninstall(ans)
    finally:
        os.chdir(cwd)
    assert relerr(S, S_ref) < TOL and relerr(S0, S0_ref) < TOL and relerr(S_lazy, S_ref) < TOL
    _columns_close(dS, dS_ref, TOL)
    _columns_close(dS_lazy, dS_ref, TOL)
    assert relerr(tg, tg_ref) < TOL and colerr(dtg, dtg_ref) < TOL


def test_line_by_line_table_deck_on_the_cuda_engine():
    from oracle.ref_import import import_reference
    from oracle import make_golden as mg
    from archnemesis_dist_b200 import forward_model as fmod
    ans = import_reference()
    deck = mg.build_jupiter_lbl_deck(os.path.join(tempfile.mkdtemp(prefix="ansb200_gl_"), "deck"), fwhm=1.5)
    ref_cls = sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0
    objs = mg.load_jupiter(ans, deck)
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        ref = mg.make_forward_model(ans, ref_cls, objs, deck)
        S_ref, dS_ref = ref.nemesisfmg()
        fmod.install(ans)
        try:
            fm = mg.make_forward_model(ans, ans.ForwardModel_0, objs, deck)
            S, dS = fm.nemesisfmg()
            assert fm._b200_hotpath().lbl_table and fm.b200_device_conv_ok(0)
        finally:
            fmod.uninstall(ans)
    finally:
        os.chdir(cwd)
    assert relerr(S, S_ref) < TOL
    _columns_close(dS, dS_ref, TOL)


def test_numerical_jacobian_states_meet_in_one_launch_group(jupiter):
    ans, deck, mg = jupiter
    from archnemesis_dist_b200 import forward_model as fmod
    ref_cls = sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        def variables(objs):
            V = objs["Variables"]
            V.FIX[:] = 1
            V.FIX[[3, 40, 77]] = 0
            return objs
        ref = mg.make_forward_model(ans, ref_cls, variables(mg.load_jupiter(ans, deck)), deck)
        XN0 = np.array(ref.Variables.XN)
        YN_ref, KK_ref = ref.jacobian_nemesis(NCores=1, analytical_gradient=False)
        KK_ref[:, 77] *= ref.Variables.XN[77] / XN0[77]     # in-process quirk of NCores=1, see test_reference_dropin.py
        fmod.install(ans)
        try:
            fm = mg.make_forward_model(ans, ans.ForwardModel_0, variables(mg.load_jupiter(ans, deck)), deck)
            YN, KK = fm.jacobian_nemesis(NCores=4, analytical_gradient=False)
            st = fm.b200_batch_stats
        finally:
            fmod.uninstall(ans)
    finally:
        os.chdir(cwd)
    assert st == dict(forward_models=4, rounds=1, launch_groups=1, evaluations=4)
    assert relerr(YN, YN_ref) < TOL
    for ix in (3, 40, 77):
        # a forward difference (y(x + dx) - y(x)) / dx of spectra that each carry a relative error e differs by up to
        # 2 e |y| / dx whatever the size of the column itself: the bar is e = 1e-12 on the spectra (element 77 barely
        # moves the spectrum, so its column is far below that scale and a column-relative bound would be meaningless)
        dx = 0.05 * abs(XN0[ix])
        assert np.abs(KK[:, ix] - KK_ref[:, ix]).max() <= 2e-12 * np.abs(YN_ref).max() / dx, ix


def test_coreretOE_with_device_forward_model_and_solver(jupiter):
    ans, deck, mg = jupiter
    from archnemesis_dist_b200 import forward_model as fmod
    oe_mod = sys.modules["archnemesis.OptimalEstimation_0"]
    cwd = os.getcwd()
    os.chdir(deck)

    def retrieve():
        o = mg.load_jupiter(ans, deck)
        return oe_mod.coreretOE(os.path.join(deck, "cirstest"), o["Variables"], o["Measurement"], o["Atmosphere"],
                                o["Spectroscopy"], o["Scatter"], o["Stellar"], o["Surface"], o["CIA"], o["Layer"], None,
                                NITER=2, PHILIMIT=0.0, NCores=1)
    try:
        ref = retrieve()
        fmod.install(ans)
        try:
            got = retrieve()
        finally:
            fmod.uninstall(ans)
    finally:
        os.chdir(cwd)
    assert type(got).__name__ == "OE_B200"
    assert relerr(got.YN, ref.YN) < 1e-8 and relerr(got.XN, ref.XN) < 1e-8
    assert abs(got.PHI - ref.PHI) <= 1e-7 * abs(ref.PHI) and abs(got.CHISQ - ref.CHISQ) <= 1e-7 * abs(ref.CHISQ)
    for name in ("KK", "DD", "AA", "SM", "SN", "ST"):
        assert colerr(getattr(got, name), getattr(ref, name)) < 1e-6, name


@pytest.mark.parametrize("driver,kind", [("nemesisSOfmg", "lbl"), ("nemesisLfmg", "k"), ("nemesisSOfmg", "lbl_fil")])
def test_limb_and_occultation_drivers_on_the_cuda_engine(driver, kind):
    """nemesisSOfmg / nemesisLfmg through install() with the CUDA engine: all tangent paths in one evaluation (layer-space
    gradients when there are >= 4 paths), ansb200_path_mix and the line shape for every geometry at once."""
    from oracle.ref_import import import_reference
    from oracle import make_golden as mg
    from archnemesis_dist_b200 import engine, forward_model as fmod
    from tests.test_reference_dropin import _tangent_objects
    ans = import_reference()
    root = os.path.join(tempfile.mkdtemp(prefix="ansb200_gso_"), "deck")
    deck = mg.build_jupiter_lbl_deck(root, fwhm=1.5) if kind.startswith("lbl") else mg.build_jupiter_deck(root)
    ref_cls = sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        ref = mg.make_forward_model(ans, ref_cls, _tangent_objects(ans, mg, deck, kind), deck)
        S_ref, dS_ref = getattr(ref, driver)()
        cls = fmod.install(ans)
        try:
            assert cls.b200_engine is engine
            calls = []
            orig = engine.HotPath.forward_jacobian_mix_conv

            def counting(self, ev, M, mix, conv_op, Mlay=None):
                calls.append((ev.LAYINC.shape[1], Mlay is not None))
                return orig(self, ev, M, mix, conv_op, Mlay)
            engine.HotPath.forward_jacobian_mix_conv = counting
            try:
                fm = mg.make_forward_model(ans, ans.ForwardModel_0, _tangent_objects(ans, mg, deck, kind), deck)
                S, dS = getattr(fm, driver)()
            finally:
                engine.HotPath.forward_jacobian_mix_conv = orig
        finally:
            fmod.uninstall(ans)
    finally:
        os.chdir(cwd)
    assert calls == [(calls[0][0], True)] and calls[0][0] >= 4
    assert relerr(S, S_ref) < TOL
    _columns_close(dS, dS_ref, TOL, floor=1e-15)


@pytest.mark.parametrize("with_ils", [False, True])
def test_calc_ktable_chunk_on_the_device(with_ils):
    """k-table generation: Spectroscopy_0.calc_ktable_chunk under install_lbl() + install_ktable() -- the line-by-line
    spectrum of every (p, T) point from ansb200_lbl_absorption, the per-bin sort and quantiles from ansb200_kdist --
    against the unmodified reference on synthetic line data."""
    from archnemesis_dist_b200 import ktable, linedata
    from tests.test_lbl_dropin import _reference_objects
    from tests.test_ktable_dropin import _case
    ans, ld, LineSetData = _reference_objects()
    sp_mod = sys.modules["archnemesis.Spectroscopy_0"]
    iwaves = np.arange(3, 12)
    S, S_LBL, M = _case(ans, ld, LineSetData, with_ils)
    ref = sp_mod.calc_ktable_chunk(iwaves, S, S_LBL, 0.3, M)
    linedata.install_lbl()
    ktable.install_ktable()
    try:
        assert isinstance(ktable._BACKEND, ktable.DeviceBackend)
        S2, S_LBL2, M2 = _case(ans, ld, LineSetData, with_ils)
        got = sp_mod.calc_ktable_chunk(iwaves, S2, S_LBL2, 0.3, M2)
    finally:
        ktable.uninstall_ktable()
        linedata.uninstall_lbl()
    assert got.shape == ref.shape and ref.max() > 0.0
    assert relerr(got, ref) < TOL
