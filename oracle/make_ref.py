"""TEST INFRASTRUCTURE ONLY -- recipe that stages the unmodified reference for the GPU box.

The reference is pure Python + numba (no build step), but ``/root/reference`` does not exist on the GPU box.  This
recipe mirrors, byte for byte, the package ``/root/reference/archnemesis`` and the one test deck the drop-in tests run
(``tests/files/Jupiter_CIRS_nadir_thermal_emission``) into ``oracle/_ref/`` -- git-ignored, so no reference source
enters the history, but not gpurun-ignored, so it travels like a built ``.so``.  ``oracle/ref_import.py`` falls back to
this mirror when ``/root/reference`` is absent.  It is used

  * by ``tests/test_gpu_reference_dropin.py``: ``install()`` + the CUDA engine against the stock reference classes on
    a real B200 (the seam the CPU suite can only exercise with the oracle-backed engine), and
  * by ``bench.py``'s ``cpu_baseline_reference`` entry: the reference's own numba ``k_overlapg`` + interpolation on a
    wavenumber slice of the timed case (``kind: "reference"``).

``__graft_entry__.build()`` runs it whenever ``/root/reference`` is present.  Nothing in the product imports it.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
DECKS = ["Jupiter_CIRS_nadir_thermal_emission"]
# data the hot path never opens (stellar spectra, partition functions are read lazily by other subsystems)
SKIP_DIRS = {"__pycache__"}


def _mirror(src, dst):
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(root, src)
        out = os.path.join(dst, rel) if rel != "." else dst
        os.makedirs(out, exist_ok=True)
        for f in files:
            if f.endswith((".pyc", ".nbi", ".nbc")):
                continue
            a, b = os.path.join(root, f), os.path.join(out, f)
            if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
                shutil.copyfile(a, b)
                n += 1
    return n


def make(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "archnemesis")):
        if verbose:
            print("oracle/make_ref: %s absent, keeping %s as it is" % (SRC, DST))
        return False
    n = _mirror(os.path.join(SRC, "archnemesis"), os.path.join(DST, "archnemesis"))
    for d in DECKS:
        n += _mirror(os.path.join(SRC, "tests", "files", d), os.path.join(DST, "tests", "files", d))
    if verbose:
        print("oracle/make_ref: %d files refreshed under %s" % (n, DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if make(verbose=True) else 1)
