"""Run-time line-by-line cross sections behind the reference's own API.

The reference computes them in ``add_line_set_monochromatic_absorption`` (archnemesis/LineData_0.py:279-358: line
strengths, Doppler / Lorentz widths and shifts at (T, p), then a line-outer / grid-inner loop that walks the whole
grid for every line, :228-277).  Its callers, from the top: ``Spectroscopy_0.calc_klbl_online`` /
``calc_klblg_online`` (Spectroscopy_0.py:1922-2143; the ``ILBL = 1`` branch of calculate_gaseous_line_opacity),
``calc_lbltable`` / ``calc_lbltable_chunk`` (:3124-3333), ``LineData_0.add_monochromatic_absorption``
(LineData_0.py:2282) and, per isotopologue, ``LineSetSpecData.add_monochromatic_absorption`` (:822-915), which
looks the numba function up by its bare name in the module.

``install_lbl`` rebinds two names and nothing else:

* ``archnemesis.LineData_0.add_line_set_monochromatic_absorption`` -> the same signature, the pair loop on the device
  (``ansb200_lbl_absorption``);
* ``LineSetSpecData.add_monochromatic_absorption`` -> the reference method's own logic (empty sets, result cache,
  wavenumber mask, the pressure-shift switch) with the line arrays kept RESIDENT on the device, keyed by the
  set's data hash, the mask range and the shift switch: a (p, T) grid or a temperature pair (calc_klblg_online
  evaluates T and T + 5 K) sends only scalars after the first call.

Line shapes the kernel knows (Voigt = SciPy's voigt_profile, Lorentz, Gaussian) run on the device; any other
``lineshape_fn`` (the sub-Lorentzian and empirical CH4-H2 shapes) goes to the reference function unchanged.  Unlike the
reference the ``store`` scratch array is not filled (nothing reads it afterwards).
"""
import sys

import numpy as np

# backend: an object with .available() and .absorption(wn_grid, lines, t, p, q, t_ref, p_ref, abundance, mass, mix,
# s_floor, calc_win, approx_win, shape) -> numpy [NWAVE]; tests swap in an oracle-backed one
_BACKEND = None
_INSTALLED = {}
_RESIDENT_MAX = 8


class DeviceBackend:
    """ansb200_lbl_absorption through archnemesis_dist_b200.lbl; `lines` may hold host arrays or resident tensors."""

    shapes = ("voigt", "lorentz", "gaussian")

    def available(self):
        import torch
        return torch.cuda.is_available()

    def resident(self, lines):
        from . import lbl
        return lbl.resident_lines(lines)

    def absorption(self, wn_grid, lines, t_calc, p_calc, q_ratio, t_ref, p_ref, abundance, mass, mix, s_floor,
                   calc_win, approx_win, shape):
        from . import lbl
        res = lbl.lbl_absorption(wn_grid, lines, [(t_calc, p_calc, q_ratio)], t_ref, p_ref, abundance, mass, mix,
                                 s_floor=s_floor, wn_calc_window=calc_win, wn_approx_window=approx_win, shape=shape)
        return res[0].cpu().numpy()


def backend():
    global _BACKEND
    if _BACKEND is None:
        _BACKEND = DeviceBackend()
    return _BACKEND


def set_backend(b):
    """Replace the device backend (tests); returns the previous one."""
    global _BACKEND
    old, _BACKEND = _BACKEND, b
    return old


def shape_of(lineshape_fn):
    """'voigt' / 'lorentz' / 'gaussian' for the reference's lineshape functions of those names, else None."""
    ls = sys.modules.get("archnemesis.lineshape")
    if ls is None:
        return None
    for name in ("voigt", "lorentz", "gaussian"):
        if lineshape_fn is getattr(ls, name, None):
            return name
    return None


def make_line_set_function(ref_fn):
    """Replacement of the module function add_line_set_monochromatic_absorption (LineData_0.py:279-358)."""

    def add_line_set_monochromatic_absorption(wn_grid, lineshape_fn, t_calc, t_ref, p_calc, p_ref, q_ratio,
                                              isotopic_abundance, isotopic_mass, mol_mix_frac, broadening_params, nu,
                                              sw, e_lower, stimulated_emission_at_t_ref, out, store=None, s_floor=0,
                                              wn_calc_window=25.0, wn_approx_window=75.0):
        shape = shape_of(lineshape_fn)
        b = backend()
        if shape is None or shape not in b.shapes or not b.available():
            return ref_fn(wn_grid, lineshape_fn, t_calc, t_ref, p_calc, p_ref, q_ratio, isotopic_abundance,
                          isotopic_mass, mol_mix_frac, broadening_params, nu, sw, e_lower,
                          stimulated_emission_at_t_ref, out, store, s_floor, wn_calc_window, wn_approx_window)
        if len(nu) == 0 or len(wn_grid) == 0:
            return None
        lines = dict(nu=nu, sw=sw, e_lower=e_lower, stim_ref=stimulated_emission_at_t_ref, broadening=broadening_params)
        out += b.absorption(wn_grid, lines, float(t_calc), float(p_calc), float(q_ratio), float(t_ref), float(p_ref),
                            float(isotopic_abundance), float(isotopic_mass), np.asarray(mol_mix_frac, dtype=np.float64),
                            float(s_floor), float(wn_calc_window), float(wn_approx_window), shape)
        return None

    add_line_set_monochromatic_absorption.__doc__ = (ref_fn.__doc__ or "") + "\n(archnemesis_dist_b200: pair loop on the device)"
    return add_line_set_monochromatic_absorption


def _masked_lines(self, wn_calc_range, include_pressure_shift):
    """The arrays the reference method hands down (LineData_0.py:883-903): rows 0-3 and the broadening block of
    `_data` under the wavenumber mask, pressure shifts zeroed on request."""
    data = self._data
    if wn_calc_range is None:
        sel = slice(None)
    else:
        sel = (wn_calc_range[0] <= data[0]) & (data[0] <= wn_calc_range[1])
    br = data[5:, sel]
    if not include_pressure_shift:
        keep = np.array([0 if ((i + 1) % 3) == 0 else 1 for i in range(br.shape[0])], dtype=int)[:, None]
        br = np.array(br) * keep
    return dict(nu=data[0, sel], sw=data[1, sel], e_lower=data[2, sel], stim_ref=data[3, sel], broadening=br)


def make_line_set_method(ref_method):
    """Replacement of LineSetSpecData.add_monochromatic_absorption (LineData_0.py:822-915) with a resident line list."""
    resident = {}

    def add_monochromatic_absorption(self, wn_grid, lineshape_fn, t_calc, p_calc, partition_function, mol_mix_frac,
                                     isotopic_abundance=1.0, out=None, store=None, s_floor=0, wn_calc_window=25.0,
                                     wn_approx_window=75.0, wn_calc_range=None, include_pressure_shift=True,
                                     use_cache=True):
        shape = shape_of(lineshape_fn)
        b = backend()
        if shape is None or shape not in b.shapes or not b.available():
            return ref_method(self, wn_grid, lineshape_fn, t_calc, p_calc, partition_function, mol_mix_frac,
                              isotopic_abundance=isotopic_abundance, out=out, store=store, s_floor=s_floor,
                              wn_calc_window=wn_calc_window, wn_approx_window=wn_approx_window,
                              wn_calc_range=wn_calc_range, include_pressure_shift=include_pressure_shift,
                              use_cache=use_cache)
        if out is None:
            out = np.zeros_like(wn_grid, dtype=float)
        if not self.has_data:
            return out
        q_ratio = partition_function(self.t_ref) / partition_function(t_calc)
        if use_cache:
            # the reference's result cache, with its own key (LineData_0.py:856-880)
            wn_grid.flags.writeable = False
            mol_mix_frac.flags.writeable = False
            bucket = (self.__class__.__name__, "add_monochromatic_absorption")
            ident = (self.cache_identity(), hash(bytes(wn_grid.data)), id(lineshape_fn), t_calc, p_calc, q_ratio,
                     hash(bytes(mol_mix_frac.data)), isotopic_abundance, s_floor, wn_calc_window, wn_approx_window,
                     wn_calc_range, include_pressure_shift)
            hit = self._result_cache.get(bucket, ident, None)
            if hit is not None:
                out[...] = hit
                return out
        key = (self._data_hash, self._data.shape, None if wn_calc_range is None else tuple(wn_calc_range),
               bool(include_pressure_shift))
        lines = resident.get(key)
        if lines is None:
            host = _masked_lines(self, wn_calc_range, include_pressure_shift)
            lines = b.resident(host) if hasattr(b, "resident") else host
            while len(resident) >= _RESIDENT_MAX:
                resident.pop(next(iter(resident)))
            resident[key] = lines
        else:
            resident[key] = resident.pop(key)        # most recently used last
        n_lines = int(lines["nu"].shape[0])
        if n_lines and len(wn_grid):
            out += b.absorption(wn_grid, lines, float(t_calc), float(p_calc), float(q_ratio), float(self.t_ref),
                                float(self.p_ref), float(isotopic_abundance), float(self._molecular_mass),
                                np.asarray(mol_mix_frac, dtype=np.float64), float(s_floor), float(wn_calc_window),
                                float(wn_approx_window), shape)
        if use_cache:
            self._result_cache.set(bucket, ident, out)
        return out

    add_monochromatic_absorption.__doc__ = (ref_method.__doc__ or "") + \
        "\n(archnemesis_dist_b200: pair loop on the device, line arrays resident)"
    add_monochromatic_absorption._b200_resident = resident
    return add_monochromatic_absorption


def _linedata_module():
    # (the package attribute archnemesis.LineData_0 is the CLASS: the star import in archnemesis/__init__.py shadows
    # the submodule, as for ForwardModel_0)
    import importlib
    importlib.import_module("archnemesis.LineData_0")
    return sys.modules["archnemesis.LineData_0"]


def install_lbl():
    """Rebind the two names (see the module docstring).  Idempotent; undone by uninstall_lbl()."""
    ld = _linedata_module()
    if "fn" not in _INSTALLED:
        _INSTALLED["fn"] = ld.add_line_set_monochromatic_absorption
        _INSTALLED["method"] = ld.LineSetSpecData.add_monochromatic_absorption
    ld.add_line_set_monochromatic_absorption = make_line_set_function(_INSTALLED["fn"])
    ld.LineSetSpecData.add_monochromatic_absorption = make_line_set_method(_INSTALLED["method"])


def uninstall_lbl():
    if "fn" not in _INSTALLED:
        return
    ld = _linedata_module()
    ld.add_line_set_monochromatic_absorption = _INSTALLED.pop("fn")
    ld.LineSetSpecData.add_monochromatic_absorption = _INSTALLED.pop("method")
