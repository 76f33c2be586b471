"""In-tree build of libansb200.so: explicit nvcc for sm_100a, one object per .cu, in parallel.

``python -m archnemesis_dist_b200.build`` (or ``__graft_entry__.build()``) leaves
``archnemesis_dist_b200/libansb200.so`` next to the package; the .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libansb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newest_header():
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "ansb200.h")]
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    log = obj[:-2] + ".ptxas.log"
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), _newest_header()):
        return obj, False
    cmd = [NVCC] + ARCH + CFLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stderr)
        raise RuntimeError("nvcc failed for %s" % src)
    if verbose:
        print("compiled", os.path.basename(src))
    return obj, True


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if force:
        for o in glob.glob(os.path.join(OBJ, "*.o")):
            os.remove(o)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    if any(ch for _, ch in results) or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        subprocess.check_call(cmd)
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
