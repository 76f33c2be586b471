"""CPU, needs the live reference: the drop-in class ForwardModel_B200 (forward_model.install) on the
reference's own Jupiter CIRS nadir deck against the unmodified ForwardModel_0.  The device layer is
replaced by the oracle-backed engine of tests/cpu_engine.py, so this exercises exactly the host
logic of the drop-in: name rebinding, continuum assembly, surface terms, projection folding, the
nemesisfmg override and the return layouts.  The CUDA kernels are checked against the same golden
data in tests/test_gpu_golden.py."""
import os
import sys
import tempfile

import numpy as np
import pytest

from tests.util import relerr, colerr

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def jupiter():
    from oracle.ref_import import import_reference
    from oracle import make_golden as mg
    ans = import_reference()
    deck = mg.build_jupiter_deck(os.path.join(tempfile.mkdtemp(prefix="ansb200_t_"), "deck"))
    return ans, deck, mg


def test_install_rebinds_three_names_and_matches_reference(jupiter):
    ans, deck, mg = jupiter
    from archnemesis_dist_b200 import forward_model as fmod
    from tests import cpu_engine
    ref_cls = sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0
    objs = mg.load_jupiter(ans, deck)
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        ref = mg.make_forward_model(ans, ref_cls, objs, deck)
        S_ref, dS_ref = ref.nemesisfmg()
        S0_ref = ref.nemesisfm()
        cls = fmod.install(ans)
        try:
            assert ans.ForwardModel_0 is cls
            assert sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0 is cls
            assert sys.modules["archnemesis.OptimalEstimation_0"].ForwardModel_0 is cls
            assert issubclass(cls, ref_cls) and issubclass(cls, fmod.B200HotPathMixin)
            cls.b200_engine = cpu_engine
            fm = mg.make_forward_model(ans, ans.ForwardModel_0, objs, deck)
            S, dS = fm.nemesisfmg()
            S0 = fm.nemesisfm()          # inherited driver, overridden CIRSrad
            # CIRSrad keeps the reference's return layout
            spec, dspec, dts = fm.CIRSrad(return_grad=True)
            assert dspec.shape == (fm.SpectroscopyX.NWAVE, fm.AtmosphereX.NVMR + 2 + fm.ScatterX.NDUST,
                                   fm.PathX.NLAYIN.max(), fm.PathX.NPATH)
            tg, dtg = fm.calculate_gaseous_line_opacity(True)
            # the REFERENCE's own driver body on the drop-in instance: CIRSrad hands back a device-resident
            # gradient, the rebound module functions map2pro / map2xvec defer and project it on the device --
            # the route nemesisSOfmg / nemesisLfmg / process_IAV take without being overridden
            mod = sys.modules["archnemesis.ForwardModel_0"]
            assert mod.map2pro is not fmod._INSTALLED["map2pro"]
            assert isinstance(dspec, fmod.DeviceGradient) and np.asarray(dspec).shape == dspec.shape
            S_lazy, dS_lazy = ref_cls.nemesisfmg(fm)
            # table residency: read_tables is memoised while installed -- the second evaluation re-uses the very
            # same host K (no disk read) and therefore the same resident engine table
            spec_cls = sys.modules["archnemesis.Spectroscopy_0"].Spectroscopy_0
            rt = spec_cls.read_tables
            assert rt is not fmod._INSTALLED["read_tables"] and rt.hits >= 2
            K_first, hp_first = fm.SpectroscopyX.K, fm._b200_hotpath()
            hits = rt.hits
            S_again, dS_again = fm.nemesisfmg()
            assert rt.hits == hits + 1 and fm.SpectroscopyX.K is K_first and fm._b200_hotpath() is hp_first
            assert not K_first.flags.writeable
            assert np.array_equal(S_again, S) and np.array_equal(dS_again, dS)
        finally:
            fmod.uninstall(ans)
        assert sys.modules["archnemesis.Spectroscopy_0"].Spectroscopy_0.read_tables.__name__ == "read_tables"
        assert not hasattr(sys.modules["archnemesis.Spectroscopy_0"].Spectroscopy_0.read_tables, "hits")
        assert ans.ForwardModel_0 is ref_cls
        assert sys.modules["archnemesis.ForwardModel_0"].map2pro.__module__ == "archnemesis.ForwardModel_0"
        tg_ref, dtg_ref = ref_cls.calculate_gaseous_line_opacity(fm, True)
    finally:
        os.chdir(cwd)
    assert relerr(S_lazy, S_ref) < 1e-12
    for ix in range(dS_ref.shape[2]):
        assert colerr(dS_lazy[:, :, ix], dS_ref[:, :, ix]) < 1e-11, ix
    assert relerr(S, S_ref) < 1e-12 and relerr(S0, S0_ref) < 1e-12
    for ix in range(dS_ref.shape[2]):
        assert colerr(dS[:, :, ix], dS_ref[:, :, ix]) < 1e-11, ix
    assert relerr(tg, tg_ref) < 1e-13 and colerr(dtg, dtg_ref) < 1e-13


def test_line_by_line_table_deck_matches_reference():
    """The same deck in LINE_BY_LINE_TABLES mode (synthetic .lta tables, Gaussian ILS of FWHM 1.5 cm-1): calc_klblg,
    the gas sum, CIRSrad, the projection and lblconvg through the drop-in (device-side convolution route) against
    the unmodified reference, plus the reference's own nemesisfmg body on the drop-in instance."""
    from oracle.ref_import import import_reference
    from oracle import make_golden as mg
    from archnemesis_dist_b200 import forward_model as fmod
    from tests import cpu_engine
    ans = import_reference()
    deck = mg.build_jupiter_lbl_deck(os.path.join(tempfile.mkdtemp(prefix="ansb200_l_"), "deck"), fwhm=1.5)
    ref_cls = sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0
    objs = mg.load_jupiter(ans, deck)
    assert int(objs["Spectroscopy"].ILBL) == 2
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        ref = mg.make_forward_model(ans, ref_cls, objs, deck)
        S_ref, dS_ref = ref.nemesisfmg()
        cls = fmod.install(ans)
        try:
            cls.b200_engine = cpu_engine
            fm = mg.make_forward_model(ans, ans.ForwardModel_0, objs, deck)
            S, dS = fm.nemesisfmg()
            assert fm._b200_mode() is not None and fm.b200_device_conv_ok(0)
            assert fm._b200_hotpath().lbl_table
            S_lazy, dS_lazy = ref_cls.nemesisfmg(fm)
            tg, dtg = fm.calculate_gaseous_line_opacity(True)
        finally:
            fmod.uninstall(ans)
        tg_ref, dtg_ref = ref_cls.calculate_gaseous_line_opacity(fm, True)
    finally:
        os.chdir(cwd)
    assert np.array_equal(tg, tg_ref) and np.array_equal(dtg, dtg_ref)
    assert relerr(S, S_ref) < 1e-13 and relerr(S_lazy, S_ref) < 1e-13
    for ix in range(dS_ref.shape[2]):
        assert colerr(dS[:, :, ix], dS_ref[:, :, ix]) < 1e-12, ix
        assert colerr(dS_lazy[:, :, ix], dS_ref[:, :, ix]) < 1e-12, ix


def test_numerical_jacobian_columns_are_batched(jupiter):
    """jacobian_nemesis(analytical_gradient=False) (ForwardModel_0.py:2184-2361): the reference runs one forward model
    per free state-vector element (joblib workers); the drop-in runs the same drivers in threads whose CIRSrad calls
    meet in ONE device evaluation (states laid side by side on the layer / path axes).  Same YN and KK."""
    ans, deck, mg = jupiter
    from archnemesis_dist_b200 import forward_model as fmod
    from tests import cpu_engine
    ref_cls = sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        def variables(objs):
            V = objs["Variables"]
            V.FIX[:] = 1
            V.FIX[[3, 40, 77]] = 0          # three free elements -> 1 + 3 forward models
            return objs
        ref = mg.make_forward_model(ans, ref_cls, variables(mg.load_jupiter(ans, deck)), deck)
        XN0 = np.array(ref.Variables.XN)
        YN_ref, KK_ref = ref.jacobian_nemesis(NCores=1, analytical_gradient=False)
        # With NCores = 1 joblib runs the forward models in THIS process and execute_fm (:2154) leaves
        # ref.Variables.XN at the last perturbed state, so the reference divides the last column's difference by
        # 0.05 * (perturbed value); its worker processes (NCores > 1) and the drop-in keep XN.  Undo that for the
        # comparison (loky workers cannot import the reference here: h5py is stubbed in this process only).
        assert ref.Variables.XN[77] != XN0[77] and np.array_equal(np.delete(ref.Variables.XN, 77), np.delete(XN0, 77))
        KK_ref[:, 77] *= ref.Variables.XN[77] / XN0[77]
        cls = fmod.install(ans)
        try:
            cls.b200_engine = cpu_engine
            calls = []
            orig = cpu_engine.HotPath.cirsrad

            def counting(self, ev, return_grad=False):
                calls.append((len(ev.press_atm), ev.LAYINC.shape[1]))
                return orig(self, ev, return_grad)
            cpu_engine.HotPath.cirsrad = counting
            try:
                fm = mg.make_forward_model(ans, ans.ForwardModel_0, variables(mg.load_jupiter(ans, deck)), deck)
                YN, KK = fm.jacobian_nemesis(NCores=4, analytical_gradient=False)
            finally:
                cpu_engine.HotPath.cirsrad = orig
            st = fm.b200_batch_stats
            # mixed analytic / numerical: NUM = 0 elements come from nemesisfmg, the rest from the batch
            fm2 = mg.make_forward_model(ans, ans.ForwardModel_0, variables(mg.load_jupiter(ans, deck)), deck)
            fm2.Variables.NUM[:] = 0
            fm2.Variables.NUM[40] = 1
            YN2, KK2 = fm2.jacobian_nemesis()
        finally:
            fmod.uninstall(ans)
        ref2 = mg.make_forward_model(ans, ref_cls, variables(mg.load_jupiter(ans, deck)), deck)
        ref2.Variables.NUM[:] = 0
        ref2.Variables.NUM[40] = 1
        YN2_ref, KK2_ref = ref2.jacobian_nemesis()
        KK2_ref[:, 40] *= ref2.Variables.XN[40] / XN0[40]       # (the same in-process quirk)
    finally:
        os.chdir(cwd)
    assert st == dict(forward_models=4, rounds=1, launch_groups=1, evaluations=4)
    nlay = calls[0][0] // 4
    assert calls == [(4 * nlay, 4)]                 # ONE engine call: four states' layers and paths side by side
    assert relerr(YN, YN_ref) < 1e-12
    assert np.array_equal(KK[:, [0, 1, 2, 50]], np.zeros((KK.shape[0], 4)))        # fixed elements stay zero
    for ix in (3, 40, 77):
        assert np.abs(KK_ref[:, ix]).max() > 0.0
        # a forward difference of two spectra that agree to ~1e-13: compare on the scale of the difference quotient
        assert colerr(KK[:, ix], KK_ref[:, ix]) < 1e-8, ix
    assert relerr(YN2, YN2_ref) < 1e-12
    for ix in range(KK2_ref.shape[1]):
        assert colerr(KK2[:, ix], KK2_ref[:, ix]) < (1e-8 if ix == 40 else 1e-11), ix


def test_coreretOE_runs_on_the_dropin_classes(jupiter):
    """One optimal-estimation retrieval (OptimalEstimation_0.coreretOE, :1173-1585; two iterations) with install():
    the forward model AND the solver object coreretOE builds are the drop-in classes (ForwardModel_B200, OE_B200 --
    gain matrix, cost function, state update, error covariances on the engine), against the unmodified reference."""
    ans, deck, mg = jupiter
    from archnemesis_dist_b200 import forward_model as fmod
    from tests import cpu_engine
    from oracle import oracle as orc
    import types
    oe_mod = sys.modules["archnemesis.OptimalEstimation_0"]
    oracle_oe = types.SimpleNamespace(calc_gain_matrix=orc.oe_gain_matrix, calc_phiret=orc.oe_phiret,
                                      calc_next_xn=orc.oe_next_xn, calc_serr=orc.oe_serr)
    cwd = os.getcwd()
    os.chdir(deck)

    def retrieve():
        o = mg.load_jupiter(ans, deck)
        return oe_mod.coreretOE(os.path.join(deck, "cirstest"), o["Variables"], o["Measurement"], o["Atmosphere"],
                                o["Spectroscopy"], o["Scatter"], o["Stellar"], o["Surface"], o["CIA"], o["Layer"], None,
                                NITER=2, PHILIMIT=0.0, NCores=1)
    try:
        ref = retrieve()
        ref_cls = type(ref)
        cls = fmod.install(ans)
        try:
            cls.b200_engine = cpu_engine
            oe_cls = fmod._INSTALLED["oe_cls"]
            assert ans.OptimalEstimation_0 is oe_cls and oe_mod.OptimalEstimation_0 is oe_cls
            assert issubclass(oe_cls, ref_cls) and oe_cls.__name__ == "OE_B200"
            oe_cls.b200_oe = oracle_oe
            calls = []
            for name in ("calc_gain_matrix", "calc_phiret", "calc_next_xn", "calc_serr"):
                def wrap(f, name=name):
                    def g(*a, **k):
                        calls.append(name)
                        return f(*a, **k)
                    return g
                setattr(oracle_oe, name, wrap(getattr(oracle_oe, name)))
            got = retrieve()
        finally:
            fmod.uninstall(ans)
        assert ans.OptimalEstimation_0 is ref_cls and oe_mod.OptimalEstimation_0 is ref_cls
    finally:
        os.chdir(cwd)
    assert type(got).__name__ == "OE_B200" and {"calc_gain_matrix", "calc_phiret", "calc_next_xn", "calc_serr"} <= set(calls)
    assert relerr(got.YN, ref.YN) < 1e-10 and relerr(got.XN, ref.XN) < 1e-9
    assert abs(got.PHI - ref.PHI) <= 1e-8 * abs(ref.PHI) and abs(got.CHISQ - ref.CHISQ) <= 1e-8 * abs(ref.CHISQ)
    for name in ("KK", "DD", "AA", "SM", "SN", "ST"):
        assert colerr(getattr(got, name), getattr(ref, name)) < 1e-8, name


def _tangent_geometries(objs, ngeom=4, lo=20.0, hi=160.0):
    """Turn the nadir measurement of the Jupiter deck into `ngeom` limb / occultation geometries on one spectral grid
    (what nemesisSOfmg / nemesisLfmg expect: TANHE per geometry, the same VCONV everywhere)."""
    M = objs["Measurement"]
    nc = int(M.NCONV[0])
    M.NGEOM = ngeom
    M.NAV = np.ones(ngeom, dtype="int32")
    M.NCONV = np.full(ngeom, nc, dtype="int32")
    M.VCONV = np.repeat(M.VCONV[:, :1], ngeom, axis=1)
    M.MEAS = np.full((M.VCONV.shape[0], ngeom), 0.5)
    M.ERRMEAS = np.repeat(M.ERRMEAS[:, :1], ngeom, axis=1)
    M.FLAT, M.FLON = np.zeros((ngeom, 1)), np.zeros((ngeom, 1))
    M.SOL_ANG, M.EMISS_ANG, M.AZI_ANG = np.full((ngeom, 1), 90.0), np.full((ngeom, 1), 90.0), np.zeros((ngeom, 1))
    M.WGEOM = np.ones((ngeom, 1))
    M.TANHE = np.linspace(lo, hi, ngeom).reshape(ngeom, 1)
    M.NY = nc * ngeom
    return objs


def _tangent_objects(ans, mg, deck, kind):
    objs = _tangent_geometries(mg.load_jupiter(ans, deck))
    if kind == "lbl_fil":
        # filter functions instead of an analytic shape (FWHM < 0: lblconvg_fil_ngeom, Measurement_0.py:2240-2248); the
        # deck's VFIL / AFIL are the tabulated Gaussian build_ils made
        objs["Measurement"].FWHM = -1.0
    return objs


@pytest.mark.parametrize("driver,kind", [("nemesisSOfmg", "lbl"), ("nemesisLfmg", "lbl"), ("nemesisSOfmg", "k"),
                                         ("nemesisLfmg", "k"), ("nemesisSOfmg", "lbl_fil")])
def test_limb_and_occultation_drivers_keep_their_tail_on_the_engine(driver, kind):
    """nemesisSOfmg / nemesisLfmg (ForwardModel_0.py:983-1243, :1372-1518) through install(): all tangent paths in one
    evaluation, then the tangent-height interpolation and the line shape for every geometry at once (IGEOM='All':
    lblconvg_ngeom for the line-by-line tables with a Gaussian of FWHM 1.5 cm-1, scipy interpolation for k-tables with
    FWHM = 0) on the engine -- only SPECMOD and [NCONV, NGEOM, 1+NX] come back -- against the unmodified reference."""
    from oracle.ref_import import import_reference
    from oracle import make_golden as mg
    from archnemesis_dist_b200 import forward_model as fmod
    from tests import cpu_engine
    ans = import_reference()
    root = os.path.join(tempfile.mkdtemp(prefix="ansb200_so_"), "deck")
    deck = mg.build_jupiter_lbl_deck(root, fwhm=1.5) if kind.startswith("lbl") else mg.build_jupiter_deck(root)
    ref_cls = sys.modules["archnemesis.ForwardModel_0"].ForwardModel_0
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        objs = _tangent_objects(ans, mg, deck, kind)
        if kind == "k":
            assert float(objs["Measurement"].FWHM) == 0.0
        ref = mg.make_forward_model(ans, ref_cls, objs, deck)
        S_ref, dS_ref = getattr(ref, driver)()
        cls = fmod.install(ans)
        try:
            cls.b200_engine = cpu_engine
            calls = []
            orig = cpu_engine.HotPath.forward_jacobian_mix_conv

            def counting(self, ev, M, mix, conv_op, Mlay=None):
                calls.append((ev.LAYINC.shape[1], len(mix["lo"]), int(np.sum(np.asarray(mix["hi"]) >= 0))))
                return orig(self, ev, M, mix, conv_op, Mlay)
            cpu_engine.HotPath.forward_jacobian_mix_conv = counting
            try:
                fm = mg.make_forward_model(ans, ans.ForwardModel_0, _tangent_objects(ans, mg, deck, kind), deck)
                S, dS = getattr(fm, driver)()
            finally:
                cpu_engine.HotPath.forward_jacobian_mix_conv = orig
        finally:
            fmod.uninstall(ans)
    finally:
        os.chdir(cwd)
    assert len(calls) == 1 and calls[0][0] >= 4 and calls[0][1] == 4 and calls[0][2] >= 3     # one evaluation, all paths
    assert S.shape == S_ref.shape and dS.shape == dS_ref.shape
    assert relerr(S, S_ref) < 1e-12
    assert np.abs(dS_ref).max() > 0.0
    for ix in range(dS_ref.shape[2]):
        assert colerr(dS[:, :, ix], dS_ref[:, :, ix]) < 1e-11, ix
