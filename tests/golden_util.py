"""Loaders for the golden fixtures (tests/golden/*.npz, written by oracle/make_golden.py from the live reference)."""
import os
import types

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name), allow_pickle=False))


def stage_table():
    """The synthetic table the stage-level goldens were made from (same seeded recipe)."""
    from archnemesis_dist_b200 import synthetic as syn
    return syn.make_fm_case(nwave=5, ng=20, npress=6, ntemp=5, ngas=3, nlay=9, nvmr=4, ndust=1, npro=9, nx=7, seed=17,
                            zero_fraction=0.2)


def jupiter_objects(g):
    """Plain namespaces carrying the attributes the hot path reads, filled from the
    arrays captured from the reference's objects on the Jupiter CIRS deck."""
    ns = types.SimpleNamespace
    K = g["K_f32"].astype(np.float64)       # read_ktable values are exactly float32-representable (Spectroscopy_0.py:2848)
    nw, ng, npr, nt, ngas = K.shape
    ids, isos = g["ATM_ID"], g["ATM_ISO"]

    def locate_gas(gas_id, iso_id):                          # Atmosphere_0.locate_gas (Atmosphere_0.py:1152-1162)
        w = np.where((ids == gas_id) & (isos == iso_id))[0]
        return int(w[0])

    sp = ns(K=K, PRESS=g["PRESS"], TEMP=g["TEMP"], DELG=g["DELG"], G_ORD=g["G_ORD"], WAVE=g["WAVE"], NWAVE=nw, NG=ng,
            NP=npr, NT=nt, NGAS=ngas, ID=g["ID"], ISO=g["ISO"], ILBL=0)
    lay = ns(NLAY=len(g["LAY_PRESS"]), PRESS=g["LAY_PRESS"], TEMP=g["LAY_TEMP"], AMOUNT=g["LAY_AMOUNT"],
             TOTAM=g["LAY_TOTAM"], DTE=g["DTE"], DAM=g["DAM"], DCO=g["DCO"])
    path = ns(NPATH=g["LAYINC"].shape[1], LAYINC=g["LAYINC"], SCALE=g["SCALE"], NLAYIN=g["NLAYIN"], EMTEMP=g["EMTEMP"],
              IMOD=g["IMOD"], SOL_ANG=g["SOL_ANG"], EMISS_ANG=g["EMISS_ANG"])
    atm = ns(NVMR=int(g["NVMR"]), NDUST=int(g["NDUST"]), NP=int(g["NP"]), ID=ids, ISO=isos, locate_gas=locate_gas,
             RADIUS=7.1e7)
    surf = ns(TSURF=float(g["TSURF"]), GASGIANT=True, LOWBC=0, VEM=None, EMISSIVITY=None)
    meas = ns(ISPACE=int(g["ISPACE"]), IFORM=int(g["IFORM"]))
    scat = ns(NDUST=int(g["NDUST"]))
    stel = ns(SOLEXIST=False)
    var = ns(NX=int(g["NX"]), JSURF=int(g["JSURF"]))
    objs = dict(SpectroscopyX=sp, LayerX=lay, PathX=path, AtmosphereX=atm, SurfaceX=surf, MeasurementX=meas,
                ScatterX=scat, StellarX=stel, Variables=var)
    cont = {k: g[k] for k in ("TAUCIA", "dTAUCIA", "TAURAY", "dTAURAY", "TAUDUST1", "TAUCLSCAT", "dTAUDUST1",
                              "dTAUCLSCAT")}
    return objs, cont
