"""Readers of the NEMESIS binary look-up tables (.kta correlated-k, .lta line-by-line) without the per-record loops.

The reference's read_ktable / read_lbltable (archnemesis/Spectroscopy_0.py:2733-2852, :2626-2729) read the block of
float32 records in one np.fromfile and then copy it record by record in Python -- nwave*npress*ntemp (resp.
nwave*npress) iterations, 7-16 s for the tables of one evaluation of the Jupiter CIRS deck, paid again on every
nemesisfm[g] call.  The copy is a reshape: records are ordered wavenumber -> pressure -> temperature -> g.  These
functions return exactly what the reference functions return (same tuple, dtypes and bits: the division by 1e20 is done
in float32 like the reference's, Spectroscopy_0.py:2849), in milliseconds.  `install_readers()` rebinds the two module
names, so Spectroscopy_0.read_tables (:1448-1528) -- which looks them up by bare name -- uses them.
"""
import sys

import numpy as np

K_PACK = 1.0e20        # BINARY_K_ABS_PACK_INTO_FLOAT_FACTOR (Spectroscopy_0.py:67)
_INSTALLED = {}


def read_ktable(filename, wavemin, wavemax):
    """Spectroscopy_0.read_ktable: gasID, isoID, nwave, wave, fwhm, ng, g_ord, del_g, npress, presslevels, ntemp,
    templevels, k_g[nwave,ng,npress,ntemp]."""
    if not filename.endswith('.kta'):
        filename += '.kta'
    with open(filename, 'rb') as f:
        head = np.fromfile(f, dtype='int32', count=10)
        irec0, nwavekta = int(head[0]), int(head[1])
        vmin, delv, fwhm = (head[2:5].view('float32')[i] for i in range(3))
        fwhm = float(fwhm)
        npress, ntemp, ng, gasID, isoID = (int(x) for x in head[5:10])
        vmin = np.round(np.float64(vmin), decimals=7)
        delv = np.round(np.float64(delv), decimals=7)
        g_ord = np.zeros(ng)
        del_g = np.zeros(ng)
        templevels = np.zeros(ntemp)
        presslevels = np.zeros(npress)
        g_ord[:] = np.fromfile(f, dtype='float32', count=ng)
        del_g[:] = np.fromfile(f, dtype='float32', count=ng)
        np.fromfile(f, dtype='float32', count=2)
        presslevels[:] = np.fromfile(f, dtype='float32', count=npress)
        templevels[:] = np.fromfile(f, dtype='float32', count=ntemp)
        if delv > 0.0:
            vmax = delv * (nwavekta - 1) + vmin
            wavetot = np.linspace(vmin, vmax, nwavekta)
        else:
            wavetot = np.zeros([nwavekta])
            wavetot[:] = np.fromfile(f, dtype='float32', count=nwavekta)
        ins = np.where((wavetot >= wavemin) & (wavetot <= wavemax))[0]
        nwave = len(ins)
        wave = np.zeros([nwave])
        wave[:] = wavetot[ins]
        njump = npress * ntemp * ng * ins[0]          # (IndexError for an empty selection, like the reference)
        f.seek(njump * 4 + (irec0 - 1) * 4, 0)
        k_out = np.fromfile(f, dtype='float32', count=ntemp * npress * ng * nwave)
    # records: wavenumber -> pressure -> temperature -> g; float32 / 1e20 stays float32 (numpy scalar promotion),
    # then widens on assignment into the float64 array
    rec = (k_out / K_PACK).reshape(nwave, npress, ntemp, ng)
    k_g = np.zeros([nwave, ng, npress, ntemp], dtype=np.float64)
    k_g[...] = np.transpose(rec, (0, 3, 1, 2))
    return gasID, isoID, nwave, wave, fwhm, ng, g_ord, del_g, npress, presslevels, ntemp, templevels, k_g


def make_read_lbltable(ref_read_lbltable):
    """Spectroscopy_0.read_lbltable with the record loop replaced (header parsed the same way, numpy integer scalars and
    float32 level arrays included)."""

    def read_lbltable(filename, wavemin, wavemax):
        if not filename.endswith('.lta'):
            filename += '.lta'
        with open(filename, 'rb') as f:
            head = np.fromfile(f, dtype='int32', count=8)
            irec0, nwavelta = head[0], head[1]
            vmin, delv = head[2:4].view('float32')
            npress, ntemp, gasID, isoID = head[4], head[5], head[6], head[7]
            vmin = np.round(np.float64(vmin), decimals=7)
            delv = np.round(np.float64(delv), decimals=7)
            presslevels = np.fromfile(f, dtype='float32', count=npress)
            if ntemp > 0:
                templevels = np.fromfile(f, dtype='float32', count=ntemp)
            else:
                templevels = np.zeros((npress, 2))
                for i in range(npress):
                    templevels[i] = np.fromfile(f, dtype='float32', count=-ntemp)
            vmax = vmin + delv * (nwavelta - 1)
            wavelta = np.linspace(vmin, vmax, nwavelta)
            wn_idxs = np.nonzero((wavemin <= wavelta) & (wavelta <= wavemax))[0]
            nwave = len(wn_idxs)
            wave = np.zeros(nwave)
            wave[:] = wavelta[wn_idxs]
            nt = abs(int(ntemp))
            njump = int(npress) * nt * int(wn_idxs[0])
            f.seek(njump * 4 + (int(irec0) - 1) * 4, 0)
            k_out = np.fromfile(f, dtype='float32', count=nt * int(npress) * nwave)
        k = np.zeros([nwave, npress, nt], dtype=np.float64)
        k[...] = (k_out / K_PACK).reshape(nwave, int(npress), nt)
        return npress, ntemp, gasID, isoID, presslevels, templevels, nwave, wave, k

    read_lbltable.__doc__ = (ref_read_lbltable.__doc__ or "") + "\n(archnemesis_dist_b200: vectorised record copy)"
    return read_lbltable


def install_readers():
    """Rebind Spectroscopy_0.read_ktable / read_lbltable (module functions read_tables calls by bare name)."""
    import importlib
    importlib.import_module("archnemesis.Spectroscopy_0")
    mod = sys.modules["archnemesis.Spectroscopy_0"]
    if "read_ktable" not in _INSTALLED:
        _INSTALLED["read_ktable"] = mod.read_ktable
        _INSTALLED["read_lbltable"] = mod.read_lbltable
    fast = read_ktable
    fast.__doc__ = (_INSTALLED["read_ktable"].__doc__ or "") + "\n(archnemesis_dist_b200: vectorised record copy)"
    mod.read_ktable = fast
    mod.read_lbltable = make_read_lbltable(_INSTALLED["read_lbltable"])


def uninstall_readers():
    if "read_ktable" not in _INSTALLED:
        return
    mod = sys.modules["archnemesis.Spectroscopy_0"]
    mod.read_ktable = _INSTALLED.pop("read_ktable")
    mod.read_lbltable = _INSTALLED.pop("read_lbltable")
