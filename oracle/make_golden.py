"""TEST INFRASTRUCTURE: generate the golden vectors under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden

Two fixture files are written (and committed, they are small):

  tests/golden/stages.npz     inputs and outputs of every hot-path function of the reference called
                              in isolation on small seeded arrays: Spectroscopy_0.calc_k / calc_kg,
                              k_overlap / k_overlapg, calc_thermal_emission_spectrum[g], map2pro,
                              map2xvec, add_line_set_monochromatic_absorption (numba + SciPy Voigt).
  tests/golden/jupiter.npz    the Jupiter CIRS nadir deck of the reference's own test
                              (tests/files/Jupiter_CIRS_nadir_thermal_emission, test_zzz_forward_models.py:155)
                              run through the unmodified ForwardModel_0.nemesisfmg with synthetic
                              .kta tables written by the reference's write_ktable (the real tables
                              are absent, .MISSING_LARGE_BLOBS): every array CIRSrad consumed on a
                              subset of wavenumber rows, and what CIRSrad / map2pro / map2xvec returned.

The script also asserts that the CPU oracle reproduces the captured reference outputs, i.e. it is
the pin of oracle/ against the reference.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference, REFERENCE_ROOT  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from archnemesis_dist_b200 import synthetic as syn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
JUPITER_GASES = [(26, 0), (27, 0), (6, 1), (6, 2), (6, 3), (28, 0), (11, 0)]   # cirstest.kls order


def rel(a, b):
    d = np.abs(np.asarray(a) - np.asarray(b))
    m = np.maximum(np.abs(a), np.abs(b))
    m[m == 0] = 1.0
    return float((d / m).max())


def build_jupiter_deck(dst, nwave=60, vmin=5.0, delv=1.0, conv_lo=12.0, conv_hi=56.0):
    """Copy the reference's Jupiter CIRS deck, write synthetic k-tables with the reference's own
    writer and shrink the .spx to a few convolution points so the case runs in seconds."""
    import_reference()
    from archnemesis.Spectroscopy_0 import write_ktable
    src = os.path.join(REFERENCE_ROOT, "tests", "files", "Jupiter_CIRS_nadir_thermal_emission")
    shutil.rmtree(dst, ignore_errors=True)
    shutil.copytree(src, dst)
    os.chmod(dst, 0o755)
    for f in os.listdir(dst):
        os.chmod(os.path.join(dst, f), 0o644)
    paths = []
    for n, (gid, iso) in enumerate(JUPITER_GASES):
        t = syn.make_ktable(nwave, 20, 8, 6, 1, seed=50 + n)
        fn = os.path.join(dst, "gas%d_%d.kta" % (gid, iso))
        write_ktable(fn, gid, iso, t["G_ORD"], t["DELG"], t["PRESS"], t["TEMP"], nwave, vmin, delv, 2.5, t["K"][..., 0])
        paths.append(fn)
    with open(os.path.join(dst, "cirstest.kls"), "w") as f:
        f.write("\n".join(paths) + "\n")
    lines = open(os.path.join(src, "cirstest.spx")).read().split("\n")
    pts = [ln for ln in lines[4:] if ln.strip()]
    sel = [p for p in pts if conv_lo <= float(p.split()[0]) <= conv_hi][::2]
    with open(os.path.join(dst, "cirstest.spx"), "w") as f:
        f.write("\n".join(lines[:1] + ["%10d" % len(sel)] + lines[2:4] + sel) + "\n")
    return dst


def build_jupiter_lbl_deck(dst, nwave=221, vmin=5.0, delv=0.25, conv_lo=12.0, conv_hi=52.0, fwhm=None):
    """The same deck in line-by-line-table mode (ILBL = 2): synthetic .lta tables written with the reference's own
    writer, a .lls list instead of the .kls, and optionally a positive FWHM in the .spx (Gaussian ILS through
    the .sha file) so that lblconvg's analytic line shape is exercised."""
    import_reference()
    from archnemesis.Spectroscopy_0 import write_lbltable
    dst = build_jupiter_deck(dst, conv_lo=conv_lo, conv_hi=conv_hi)
    paths = []
    for n, (gid, iso) in enumerate(JUPITER_GASES):
        t = syn.make_ktable(nwave, 1, 8, 6, 1, seed=150 + n)
        fn = os.path.join(dst, "gas%d_%d.lta" % (gid, iso))
        write_lbltable(fn, 8, 6, gid, iso, t["PRESS"], t["TEMP"], nwave, vmin, delv, t["K"][:, 0, :, :, 0])
        paths.append(fn)
    with open(os.path.join(dst, "cirstest.lls"), "w") as f:
        f.write("\n".join(paths) + "\n")
    os.remove(os.path.join(dst, "cirstest.kls"))
    inp = open(os.path.join(dst, "cirstest.inp")).read().split("\n")
    inp[0] = "0 0 2\t\t\t! ISPACE, ISCAT, ILBL"
    with open(os.path.join(dst, "cirstest.inp"), "w") as f:
        f.write("\n".join(inp))
    if fwhm is not None:
        spx = open(os.path.join(dst, "cirstest.spx")).read().split("\n")
        head = spx[0].split()
        head[0] = "%.5f" % fwhm
        spx[0] = "  ".join(head)
        with open(os.path.join(dst, "cirstest.spx"), "w") as f:
            f.write("\n".join(spx))
        with open(os.path.join(dst, "cirstest.sha"), "w") as f:
            f.write("2\n")
    return dst


def load_jupiter(ans, deck):
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        objs = ans.Files.read_input_files("cirstest")
    finally:
        os.chdir(cwd)
    names = ["Atmosphere", "Measurement", "Spectroscopy", "Scatter", "Stellar", "Surface", "CIA", "Layer", "Variables"]
    return dict(zip(names, objs[:9]))


def make_forward_model(ans, cls, objs, deck):
    return cls(runname=os.path.join(deck, "cirstest"), Atmosphere=objs["Atmosphere"], Surface=objs["Surface"],
               Measurement=objs["Measurement"], Spectroscopy=objs["Spectroscopy"], Stellar=objs["Stellar"],
               Scatter=objs["Scatter"], CIA=objs["CIA"], Layer=objs["Layer"], Variables=objs["Variables"])


def golden_jupiter(ans):
    from archnemesis_dist_b200.forward_model import B200HotPathMixin
    fm_mod = sys.modules["archnemesis.ForwardModel_0"]   # the package attribute of that name is the class
    deck = build_jupiter_deck(os.path.join(tempfile.mkdtemp(prefix="ansb200_"), "deck"))
    objs = load_jupiter(ans, deck)
    cap = {}

    class Capture(fm_mod.ForwardModel_0):
        def CIRSrad(self, return_grad=False):
            out = super().CIRSrad(return_grad)
            if return_grad:
                cap["self"] = self
                for nm in ("SpectroscopyX", "LayerX", "PathX", "AtmosphereX", "SurfaceX", "MeasurementX", "ScatterX"):
                    cap[nm] = getattr(self, nm)     # nemesisfm[g] makes fresh copies per call: keep these
                cap["out"] = out
                # continuum terms exactly as the drop-in assembles them from the reference's routines
                cap["cont"] = B200HotPathMixin._b200_continuum(self, True)
                cia = self.calculate_vertical_cia_opacity(True)
                ray = self.calc_tau_rayleigh(MakePlot=False)
                dust = self.calc_tau_dust()
                cap["raw"] = dict(TAUCIA=cia[0], dTAUCIA=cia[1], TAURAY=ray[0], dTAURAY=ray[1], TAUDUST1=dust[0],
                                  TAUCLSCAT=dust[1], dTAUDUST1=dust[2], dTAUCLSCAT=dust[3])
            return out

    orig_m2x = fm_mod.map2xvec

    def m2x(*a, **k):
        r = orig_m2x(*a, **k)
        cap["dSPEC1"] = r.copy()
        cap["xmap"] = a[-1].copy()
        return r

    fm_mod.map2xvec = m2x
    cwd = os.getcwd()
    os.chdir(deck)
    try:
        FM = make_forward_model(ans, Capture, objs, deck)
        SPECONV_fm = FM.nemesisfm()
        SPECONV, dSPECONV = FM.nemesisfmg()
    finally:
        os.chdir(cwd)
        fm_mod.map2xvec = orig_m2x
    s = cap["self"]
    sp, lay, path, atm = cap["SpectroscopyX"], cap["LayerX"], cap["PathX"], cap["AtmosphereX"]
    SPECOUT, dSPECOUT, dTSURF = cap["out"]
    TAUCIA, TAUDUST, TAURAY, dTAUCON = cap["cont"]
    gas_slot = np.array([atm.locate_gas(sp.ID[i], sp.ISO[i]) for i in range(sp.NGAS)], np.int32)
    amount = np.stack([lay.AMOUNT[:, g] * 1.0e-4 for g in gas_slot])

    # ---- pin the oracle against what the reference just computed (all rows) -----------------------
    NPAR = atm.NVMR + 2 + cap["ScatterX"].NDUST
    k, dkdT = orc.calc_k(sp.K, sp.PRESS, sp.TEMP, lay.PRESS / 101325.0, lay.TEMP, want_grad=True)
    tau, dk = orc.k_overlap(sp.DELG, k, amount, dkdT=dkdT)
    t = tau + TAUCIA[:, None, :] + TAUDUST[:, None, :] + TAURAY[:, None, :]
    z = np.zeros_like(TAUCIA)
    tl, tp, dtl = orc.assemble_opacity(t, dk, gas_slot, atm.NVMR, NPAR, z, dTAUCON, path.LAYINC, path.SCALE)
    emis = np.zeros(sp.NWAVE)
    S, dS, dT = orc.thermal_paths(int(cap["MeasurementX"].ISPACE), sp.WAVE, tl, dtl, atm.NVMR, path.NLAYIN, path.EMTEMP,
                                  lay.PRESS, path.LAYINC, float(cap["SurfaceX"].TSURF), emis, np.ones(sp.NWAVE))
    o_spec, o_dspec, o_dts = orc.g_integrate(S, dS, dT, sp.DELG)
    e1, e2 = rel(o_spec, SPECOUT), float(np.abs(o_dspec - dSPECOUT).max() / np.abs(dSPECOUT).max())
    print("jupiter: oracle vs reference CIRSrad  spec %.2e  dspec(col) %.2e" % (e1, e2))
    assert e1 < 1e-12 and e2 < 1e-12
    for kpar in range(dSPECOUT.shape[1]):       # every parameter on its own scale (tie-sensitive gases included)
        m = np.abs(dSPECOUT[:, kpar]).max()
        if m > 0:
            assert np.abs(o_dspec[:, kpar] - dSPECOUT[:, kpar]).max() / m < 1e-12, kpar
    inc = orc.included_params(cap["xmap"])
    d2 = orc.map2pro(o_dspec, sp.NWAVE, atm.NVMR, atm.NDUST, atm.NP, path.NPATH, path.NLAYIN, path.LAYINC, lay.DTE,
                     lay.DAM, lay.DCO, INCPAR=inc)
    e3 = float(np.abs(orc.map2xvec(d2, cap["xmap"]) - cap["dSPEC1"]).max() / np.abs(cap["dSPEC1"]).max())
    print("jupiter: oracle vs reference map2xvec %.2e" % e3)
    assert e3 < 1e-12

    rows = np.unique(np.linspace(0, sp.NWAVE - 1, 8).astype(int))
    # read_ktable divides the float32 file values by 1e20 in float32 (Spectroscopy_0.py:2848): exact in f32
    assert np.array_equal(sp.K.astype(np.float32).astype(np.float64), sp.K)
    raw = cap["raw"]
    out = dict(
        rows=rows, K_f32=sp.K[rows].astype(np.float32), PRESS=sp.PRESS, TEMP=sp.TEMP, DELG=sp.DELG,
        G_ORD=sp.G_ORD, WAVE=sp.WAVE[rows], ID=np.asarray(sp.ID), ISO=np.asarray(sp.ISO),
        LAY_PRESS=lay.PRESS, LAY_TEMP=lay.TEMP, LAY_AMOUNT=lay.AMOUNT, LAY_TOTAM=lay.TOTAM, DTE=lay.DTE, DAM=lay.DAM,
        DCO=lay.DCO, LAYINC=path.LAYINC, SCALE=path.SCALE, NLAYIN=path.NLAYIN, EMTEMP=path.EMTEMP,
        IMOD=np.asarray(path.IMOD).astype(np.int64), SOL_ANG=np.asarray(path.SOL_ANG, float),
        EMISS_ANG=np.asarray(path.EMISS_ANG, float), ATM_ID=np.asarray(atm.ID), ATM_ISO=np.asarray(atm.ISO),
        NVMR=atm.NVMR, NDUST=atm.NDUST, NP=atm.NP, TSURF=float(cap["SurfaceX"].TSURF), ISPACE=int(cap["MeasurementX"].ISPACE),
        IFORM=int(cap["MeasurementX"].IFORM), NX=s.Variables.NX, JSURF=int(s.Variables.JSURF), xmap=cap["xmap"],
        TAUCIA=raw["TAUCIA"][rows], dTAUCIA=raw["dTAUCIA"][rows], TAURAY=raw["TAURAY"][rows],
        dTAURAY=raw["dTAURAY"][rows], TAUDUST1=raw["TAUDUST1"][rows], TAUCLSCAT=raw["TAUCLSCAT"][rows],
        dTAUDUST1=raw["dTAUDUST1"][rows], dTAUCLSCAT=raw["dTAUCLSCAT"][rows],
        ref_SPECOUT=SPECOUT[rows], ref_dSPECOUT=dSPECOUT[rows], ref_dTSURF=dTSURF[rows], ref_dSPEC1=cap["dSPEC1"][rows],
        ref_SPECONV=SPECONV, ref_dSPECONV=dSPECONV, ref_SPECONV_fm=SPECONV_fm,
        gas_slot=gas_slot)
    np.savez_compressed(os.path.join(GOLD, "jupiter.npz"), **out)
    print("wrote jupiter.npz", os.path.getsize(os.path.join(GOLD, "jupiter.npz")) // 1024, "KiB")


def golden_stages(ans):
    from archnemesis.ForwardModel_0 import (k_overlap, k_overlapg, calc_thermal_emission_spectrum,
                                            calc_thermal_emission_spectrumg, map2pro, map2xvec)
    from archnemesis.LineData_0 import add_line_set_monochromatic_absorption
    from archnemesis.lineshape import voigt
    out = {}
    # ---- k-interp + overlap -------------------------------------------------------------------------
    c = syn.make_fm_case(nwave=5, ng=20, npress=6, ntemp=5, ngas=3, nlay=9, nvmr=4, ndust=1, npro=9, nx=7, seed=17,
                         zero_fraction=0.2)
    tab = c["tab"]
    S = ans.Spectroscopy_0(ILBL=0)
    for kk in ("K", "PRESS", "TEMP", "G_ORD", "DELG", "WAVE", "NWAVE", "NG", "NP", "NT", "NGAS"):
        setattr(S, kk, tab[kk])
    press, temp = c["press"].copy(), c["temp"].copy()
    press[0], press[-1], temp[2], temp[4] = 20.0, 1e-8, 50.0, 400.0
    k = S.calc_k(len(press), press, temp)
    kg, dkdT = S.calc_kg(len(press), press, temp)
    tau = k_overlap(tab["DELG"], k, c["amount"])
    taug, dk = k_overlapg(tab["DELG"], kg, dkdT, c["amount"])
    tau64 = k_overlap(tab["DELG"].astype(np.float64), k, c["amount"])
    out.update(ko_seed=17, ko_press=press, ko_temp=temp, ko_amount=c["amount"], ko_k=k, ko_kg=kg, ko_dkdT=dkdT, ko_tau=tau,
               ko_taug=taug, ko_dk=dk, ko_tau_f64delg=tau64)
    # oracle pin
    assert rel(orc.calc_k(tab["K"], tab["PRESS"], tab["TEMP"], press, temp), k) < 1e-15 * 4
    assert np.array_equal(orc.k_overlap(tab["DELG"], k, c["amount"]), tau)
    og = orc.k_overlap(tab["DELG"], kg, c["amount"], dkdT=dkdT)
    assert np.array_equal(og[0], taug) and np.array_equal(og[1], dk)
    # ---- thermal emission ---------------------------------------------------------------------------
    tl, tp, dtl = orc.assemble_opacity(taug, dk, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"], c["dtaucon"],
                                       c["LAYINC"], c["SCALE"])
    emt, emp = c["EMTEMP"][:, 0], c["LAYPRESS"][c["LAYINC"][:, 0]]
    z = np.zeros(5)
    em = np.full(5, 0.9)
    t0 = np.ascontiguousarray(tl[:, :, :, 0])
    d0 = np.ascontiguousarray(dtl[:, :, :, :, 0])
    out.update(th_tau=t0, th_dtau=d0, th_emtemp=emt, th_empress=emp, th_wave=tab["WAVE"], th_nvmr=c["NVMR"])
    for tag, ispace, wave, tsurf, emis in (("a", 0, tab["WAVE"], -1.0, z), ("b", 0, tab["WAVE"], 150.0, em),
                                           ("c", 1, 1e4 / tab["WAVE"], 150.0, em)):
        s = calc_thermal_emission_spectrum(ispace, wave, t0, None, emt, emp, tsurf, emis, z, z, 100.0, 10.0)
        sg, dsg, dts = calc_thermal_emission_spectrumg(ispace, wave, t0, d0, c["NVMR"], emt, emp, tsurf, emis)
        out.update({"th_%s_spec" % tag: s, "th_%s_specg" % tag: sg, "th_%s_dspec" % tag: dsg, "th_%s_dts" % tag: dts})
        assert np.array_equal(orc.thermal(ispace, wave, t0, None, emt, emp, tsurf, emis, z, z, 100.0, 10.0), s)
        o = orc.thermalg(ispace, wave, t0, d0, c["NVMR"], emt, emp, tsurf, emis)
        assert np.array_equal(o[0], sg) and np.array_equal(o[1], dsg) and np.array_equal(o[2], dts)
    # ---- transmission, several ragged limb paths (calculate_transmission_spectrum, :4104-4129) ----------------
    import types
    nlay = 9
    tang = [1, 4, 6]                                   # tangent layers: path p sees layers above it, in and out
    nlayin = np.array([2 * (nlay - t) for t in tang], np.int32)
    nlm = int(nlayin.max())
    layinc = np.zeros((nlm, 3), np.int32)
    scale = np.zeros((nlm, 3))
    for p_, t in enumerate(tang):
        seq = list(range(nlay - 1, t - 1, -1)) + list(range(t, nlay))
        layinc[:len(seq), p_] = seq
        scale[:len(seq), p_] = 1.0 + 12.0 / (1.0 + np.abs(np.array(seq) - t))
    # (rows past NLAYIN carry SCALE 0 and layer 0, like Path_0 pads them: zero opacity, zero gradient)
    ttl, ttp, dttl = orc.assemble_opacity(taug * 1e-3, dk * 1e-3, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"],
                                          c["dtaucon"], layinc, scale)
    stub = types.SimpleNamespace(SpectroscopyX=types.SimpleNamespace(NWAVE=5, NG=20),
                                 MeasurementX=types.SimpleNamespace(IFORM=0), PathX=types.SimpleNamespace(NPATH=3))
    tr_spec, tr_dspec = ans.ForwardModel_0.calculate_transmission_spectrum(stub, ttp, dttl, return_grad=True)
    # g-integration exactly as CIRSrad does it (:4504-4507)
    delg = tab["DELG"]
    tr_s = np.tensordot(tr_spec, delg, axes=([1], [0]))
    tr_d = np.nan_to_num(np.tensordot(tr_dspec, delg, axes=([1], [0])))
    o_s, o_d = orc.transmission(ttp, dttl)
    assert np.array_equal(o_s, tr_spec) and np.array_equal(o_d, tr_dspec)
    og_s, og_d, _ = orc.g_integrate(o_s, o_d, None, delg)
    assert np.array_equal(og_s, tr_s) and np.array_equal(og_d, tr_d)
    out.update(tr_layinc=layinc, tr_scale=scale, tr_nlayin=nlayin, tr_tau=taug * 1e-3, tr_dk=dk * 1e-3,
               tr_gas_slot=c["gas_slot"], tr_taucon=c["taucon"], tr_dtaucon=c["dtaucon"], tr_spec=tr_s, tr_dspec=tr_d)
    # ---- instrument line shape: Measurement_0.convg, FWHM == 0 and FWHM < 0 (:2467-2692) --------------------------
    from archnemesis_dist_b200 import plan as b2plan
    rng = np.random.default_rng(23)
    cw = np.linspace(600.0, 640.0, 161)                      # calculation grid
    cy = rng.uniform(0.5, 2.0, 161) * 1e-7
    cg = rng.normal(size=(161, 5)) * 1e-9
    vconv = np.sort(rng.uniform(603.0, 637.0, 11))
    vconv[3], vconv[10] = cw[40], cw[160]                    # on a knot, and on the last knot
    M0 = ans.Measurement_0(NGEOM=1, FWHM=0.0, NCONV=np.array([11], dtype="int32"))
    M0.VCONV = vconv.reshape(11, 1)
    y0, g0 = M0.convg(cw, cy, cg, IGEOM=0)
    op0 = b2plan.conv_operator(cw, vconv, 0.0)
    assert np.array_equal(orc.apply_conv(op0, cy), y0) and np.array_equal(orc.apply_conv(op0, cg), g0)
    assert np.array_equal(M0.conv(cw, cy, IGEOM=0), y0)
    vconv1 = np.clip(vconv, 604.0, 636.0)                    # the filters must stay inside the calculation grid
    nfil = np.array([5, 7, 9, 4, 6, 8, 5, 7, 9, 6, 4], dtype="int32")
    vfil = np.zeros((9, 11))
    afil = np.zeros((9, 11))
    for ic in range(11):
        half = rng.uniform(0.6, 2.4)
        vfil[:nfil[ic], ic] = np.linspace(vconv1[ic] - half, vconv1[ic] + half, nfil[ic])
        afil[:nfil[ic], ic] = np.maximum(0.0, 1.0 - np.abs(np.linspace(-1.0, 1.0, nfil[ic])) ** 2) + 1e-3 * (ic % 3 == 0)
    M1 = ans.Measurement_0(NGEOM=1, FWHM=-1.0, NCONV=np.array([11], dtype="int32"))
    M1.VCONV, M1.NFIL, M1.VFIL, M1.AFIL = vconv1.reshape(11, 1), nfil, vfil, afil
    y1, g1 = M1.convg(cw, cy, cg, IGEOM=0)
    op1 = b2plan.conv_operator(cw, vconv1, -1.0, nfil, vfil, afil)
    assert np.array_equal(orc.apply_conv(op1, cy), y1) and np.array_equal(orc.apply_conv(op1, cg), g1)
    assert np.array_equal(M1.conv(cw, cy, IGEOM=0), y1)
    # integrated radiance over the same filters: Measurement_0.integrate_filterg (:2742-2800, :4188-4250)
    M1.V_DOPPLER = 0.0
    yi, gi = M1.integrate_filterg(cw, cy, cg, IGEOM=0)
    opi = b2plan.filter_integral_operator(cw, 11, nfil, vfil, afil)
    assert rel(orc.apply_conv(opi, cy), yi) < 1e-14 and rel(orc.apply_conv(opi, cg), gi) < 1e-12
    out.update(cv_yi=yi, cv_gi=gi)
    out.update(cv_wave=cw, cv_y=cy, cv_grad=cg, cv_vconv=vconv, cv_vconv1=vconv1, cv_nfil=nfil, cv_vfil=vfil, cv_afil=afil,
               cv_y0=y0, cv_g0=g0, cv_y1=y1, cv_g1=g1)
    # ---- map2pro / map2xvec -----------------------------------------------------------------------------
    rng = np.random.default_rng(5)
    dspec = rng.normal(size=(5, c["NPAR"], 9, 1))
    inc = [i for i in range(c["NPAR"]) if np.mean(c["xmap"][:, i, :]) != 0.0]
    d2 = map2pro(dspec, 5, c["NVMR"], c["NDUST"], c["NPRO"], 1, c["NLAYIN"], c["LAYINC"], c["DTE"], c["DAM"], c["DCO"],
                 INCPAR=inc)
    dx = map2xvec(d2, 5, c["NVMR"], c["NDUST"], c["NPRO"], 1, c["NX"], c["xmap"])
    out.update(mp_dspec=dspec, mp_d2=d2, mp_dx=dx, mp_inc=np.array(inc))
    # ---- line by line ---------------------------------------------------------------------------------------
    wn = np.linspace(1000.0, 1004.0, 801)
    lines = syn.make_line_list(60, 1000.0, 1004.0, seed=4, pad=30.0)
    mix = np.array([0.05, 0.95])
    pts = [(200.0, 0.1, 1.3), (296.0, 1.0, 1.0), (120.0, 1e-4, 4.0)]
    res = []
    for (tc, pc, q) in pts:
        o = np.zeros(len(wn))
        add_line_set_monochromatic_absorption(wn, voigt, tc, 296.0, pc, 1.0, q, 0.98, 28.0, mix, lines["broadening"],
                                              lines["nu"], lines["sw"], lines["e_lower"], lines["stim_ref"], o)
        res.append(o)
        assert rel(orc.lbl_absorption(wn, lines, tc, pc, 296.0, 1.0, q, 0.98, 28.0, mix), o) < 1e-13
    out.update(lbl_wn=wn, lbl_pts=np.array(pts), lbl_mix=mix, lbl_out=np.array(res),
               **{"lbl_" + kk: v for kk, v in lines.items()})
    np.savez_compressed(os.path.join(GOLD, "stages.npz"), **out)
    print("wrote stages.npz", os.path.getsize(os.path.join(GOLD, "stages.npz")) // 1024, "KiB")


def golden_lbl_table(ans):
    """Line-by-line tables: Spectroscopy_0.calc_klbl / calc_klblg, the LBL branch of
    ForwardModel_0.calculate_gaseous_line_opacity, Measurement_0.lblconv / lblconvg -> lbl_table.npz."""
    import types
    from archnemesis_dist_b200 import plan as b2plan
    out = {}
    rng = np.random.default_rng(41)
    for tag, ngas, gdt in (("a", 3, np.float32), ("b", 11, np.float64)):
        nw, npg, ntg, nlay = 7, 5, 4, 10
        K = np.exp(rng.uniform(-60.0, -40.0, size=(nw, npg, ntg, ngas)))
        K[rng.uniform(size=K.shape) < 0.08] = 0.0                    # mixed corners -> zero
        K[2, :, :, 0] = 0.0                                          # all four corners zero -> linear branch
        K[4, :, :, 1] = -np.abs(rng.normal(size=(npg, ntg))) * 1e-30  # negative table values -> linear branch
        PRESS = np.exp(np.linspace(np.log(1e-6), np.log(10.0), npg)).astype(gdt)
        TEMP = np.linspace(80.0, 320.0, ntg).astype(gdt)
        press = np.exp(rng.uniform(np.log(2e-6), np.log(5.0), nlay))
        temp = rng.uniform(90.0, 310.0, nlay)
        press[0], press[1], press[2] = 50.0, 1e-9, float(PRESS[2])   # clamped high / low, on a node
        temp[3], temp[4], temp[5], temp[6] = 20.0, 500.0, float(TEMP[0]), float(TEMP[2])   # clamped, first node, a node
        S = ans.Spectroscopy_0(ILBL=2)
        S.K, S.PRESS, S.TEMP, S.NWAVE, S.NG, S.NP, S.NT, S.NGAS = K, PRESS, TEMP, nw, 1, npg, ntg, ngas
        S.DELG = np.array([1.0])
        S.ID, S.ISO = np.arange(1, ngas + 1), np.zeros(ngas, dtype=int)
        k = S.calc_klbl(nlay, press, temp)
        kg, dkdT = S.calc_klblg(nlay, press, temp)
        nvmr = ngas + 2
        slot = rng.permutation(nvmr)[:ngas]
        AMOUNT = np.exp(rng.uniform(np.log(1e20), np.log(1e26), size=(nlay, nvmr)))
        stub = types.SimpleNamespace(
            SpectroscopyX=S, ScatterX=types.SimpleNamespace(NDUST=1),
            LayerX=types.SimpleNamespace(NLAY=nlay, PRESS=press * 101325.0, TEMP=temp, AMOUNT=AMOUNT),
            AtmosphereX=types.SimpleNamespace(NVMR=nvmr, locate_gas=lambda gid, iso: int(slot[int(gid) - 1])))
        tau, _ = ans.ForwardModel_0.calculate_gaseous_line_opacity(stub, return_grad=False)
        taug, dtau = ans.ForwardModel_0.calculate_gaseous_line_opacity(stub, return_grad=True)
        # the layer pressures the reference interpolates at are LayerX.PRESS / 101325, not `press` itself
        pl = stub.LayerX.PRESS / 101325.0
        k, (kg, dkdT) = S.calc_klbl(nlay, pl, temp), S.calc_klblg(nlay, pl, temp)
        assert np.array_equal(orc.calc_klbl(K, PRESS, TEMP, pl, temp), k)
        o = orc.calc_klbl(K, PRESS, TEMP, pl, temp, want_grad=True)
        assert np.array_equal(o[0], kg) and np.array_equal(o[1], dkdT)
        amount = np.stack([AMOUNT[:, g] * 1.0e-4 for g in slot])
        assert np.array_equal(orc.lbl_table_opacity(K, PRESS, TEMP, pl, temp, amount), tau)
        ot, odk = orc.lbl_table_opacity(K, PRESS, TEMP, pl, temp, amount, want_grad=True)
        assert np.array_equal(ot, taug)
        for i, g in enumerate(slot):
            assert np.array_equal(odk[:, :, :, i] * 1.0e-4, dtau[:, :, g, :])
        assert np.array_equal(odk[:, :, :, ngas], dtau[:, :, nvmr, :])
        out.update({"%s_K" % tag: K, "%s_PRESS" % tag: PRESS, "%s_TEMP" % tag: TEMP, "%s_press" % tag: pl,
                    "%s_temp" % tag: temp, "%s_amount" % tag: amount, "%s_k" % tag: k, "%s_kg" % tag: kg,
                    "%s_dkdT" % tag: dkdT, "%s_tau" % tag: tau, "%s_taug" % tag: taug, "%s_dk" % tag: odk})
    # ---- lblconv / lblconvg (Measurement_0.py:2125-2284, kernels :3335-4076) -------------------------------------
    cw = np.linspace(2000.0, 2004.0, 801)
    cy = rng.uniform(0.5, 2.0, 801) * 1e-7
    cg = rng.normal(size=(801, 4)) * 1e-9
    vconv = np.sort(rng.uniform(2000.4, 2003.6, 9))
    vconv[2] = cw[300]
    out.update(lc_wave=cw, lc_y=cy, lc_grad=cg, lc_vconv=vconv)
    for tag, fwhm, ishape in (("sq", 0.11, 0), ("tr", 0.09, 1), ("ga", 0.07, 2), ("ha", 0.10, 3), ("ip", 0.0, 2)):
        M = ans.Measurement_0(NGEOM=1, FWHM=fwhm, ISHAPE=ishape, NCONV=np.array([9], dtype="int32"))
        M.VCONV, M.V_DOPPLER = vconv.reshape(9, 1), 0.0
        yg, gg = M.lblconvg(cw, cy, cg, IGEOM=0)
        op = b2plan.lbl_conv_operator(cw, vconv, fwhm, ishape)
        assert np.array_equal(orc.apply_conv(op, cy), yg), tag
        assert np.array_equal(orc.apply_conv(op, cg), gg), tag
        # the spectrum-only lblconv is plain Python (different exp / pow roundings; an empty Hamming window -> NaN)
        op0 = b2plan.lbl_conv_operator(cw, vconv, fwhm, ishape, grad=False)
        with np.errstate(all="ignore"):
            y0 = M.lblconv(cw, cy, IGEOM=0)
            assert np.array_equal(orc.apply_conv(op0, cy), y0, equal_nan=True), tag
        out.update({"lc_%s_y" % tag: yg, "lc_%s_g" % tag: gg, "lc_%s_y0" % tag: y0, "lc_%s_fwhm" % tag: fwhm,
                    "lc_%s_ishape" % tag: ishape})
    nfil = np.array([5, 7, 9, 4, 6, 8, 5, 7, 9], dtype="int32")
    vfil, afil = np.zeros((9, 9)), np.zeros((9, 9))
    for ic in range(9):
        half = rng.uniform(0.05, 0.3)
        vfil[:nfil[ic], ic] = np.linspace(vconv[ic] - half, vconv[ic] + half, nfil[ic])
        afil[:nfil[ic], ic] = np.maximum(0.0, 1.0 - np.abs(np.linspace(-1.0, 1.0, nfil[ic])) ** 2) + 1e-3 * (ic % 3 == 0)
    M = ans.Measurement_0(NGEOM=1, FWHM=-1.0, NCONV=np.array([9], dtype="int32"))
    M.VCONV, M.V_DOPPLER, M.NFIL, M.VFIL, M.AFIL = vconv.reshape(9, 1), 0.0, nfil, vfil, afil
    yf, gf = M.lblconvg(cw, cy, cg, IGEOM=0)
    op = b2plan.lbl_conv_operator(cw, vconv, -1.0, NFIL=nfil, VFIL=vfil, AFIL=afil)
    assert np.array_equal(orc.apply_conv(op, cy), yf) and np.array_equal(orc.apply_conv(op, cg), gf)
    assert np.array_equal(M.lblconv(cw, cy, IGEOM=0), yf)
    out.update(lc_nfil=nfil, lc_vfil=vfil, lc_afil=afil, lc_fil_y=yf, lc_fil_g=gf)
    np.savez_compressed(os.path.join(GOLD, "lbl_table.npz"), **out)
    print("wrote lbl_table.npz", os.path.getsize(os.path.join(GOLD, "lbl_table.npz")) // 1024, "KiB")


def main():
    os.makedirs(GOLD, exist_ok=True)
    ans = import_reference()
    if "--lbl-table-only" in sys.argv:
        golden_lbl_table(ans)
        return
    golden_stages(ans)
    golden_lbl_table(ans)
    golden_jupiter(ans)


if __name__ == "__main__":
    main()
