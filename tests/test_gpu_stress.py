"""GPU: short randomised sweep of the overlap kernels against the oracle (tools/stress_overlap.py): random shapes,
44 orders of magnitude between gases (static orders, exact ties), scrambled g-ordering, non-positive and denormal
gases, near-equal gases (packed-key collisions), float32 / float64 quadrature weights.  The parallel rebins must
agree with the reference-order oracle to rounding, the sequential rebin bit for bit."""
import numpy as np
import pytest

from tools import stress_overlap as so

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [2024, 7])
def test_overlap_random_sweep(seed):
    rng = np.random.default_rng(seed)
    for case in range(18):
        k, dkdT, amount, dg, desc = so.make_case(rng, case)
        r = so.run_case(k, dkdT, amount, dg)
        assert r["seq_exact"], desc
        assert r["finite"], desc
        # float64 quadrature weights are not exactly summable: the order of the cumulative sum shows at 1e-14
        assert r["tau"] < 2e-13 and r["tau_nograd"] < 2e-13, (desc, r)
        assert r["dk"] < 1e-11, (desc, r)


@pytest.mark.parametrize("seed", [5, 99])
def test_radiance_projection_random_sweep(seed):
    """tools/stress_radiance.py: thermal / transmission, 1..33 nadir or ragged limb paths (one path per CTA, the
    staged multi-path kernel and the warp-per-path transmission kernel), with and without gradients, surface,
    dust / Rayleigh terms, wavelength space, followed by the projection onto the state vector."""
    from tools import stress_radiance as sr
    rng = np.random.default_rng(seed)
    for case in range(16):
        desc, e_s, e_d, e_x = sr.run_case(rng)
        assert e_s < 1e-11, (desc, e_s)
        assert e_d < 1e-10 and e_x < 1e-10, (desc, e_d, e_x)


def test_lbl_random_sweep():
    """tools/stress_lbl.py: pressures 1e-9..100 atm, temperatures 60..900 K, grids of 0.3..100 cm-1, windows, strength
    floor, one or two broadeners, against the oracle's SciPy-Voigt restatement of add_line_set_monochromatic_absorption."""
    from tools import stress_lbl as sl
    rng = np.random.default_rng(3)
    for case in range(8):
        desc, e = sl.run_case(rng)
        assert e < 1e-11, (desc, e)
