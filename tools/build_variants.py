#!/usr/bin/env python
"""Build experiment variants of libansb200.so into scratch/variants/<name>.so:
   python tools/build_variants.py name1="-DOV_MINB=1" name2="-DOV_NWARPS=4 -DOV_MINB=4" ...
and time them on the GPU with tools/time_variants.py."""
import concurrent.futures, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "archnemesis_dist_b200", "csrc")
OUT = os.path.join(ROOT, "scratch", "variants")
NVCC = "/usr/local/cuda/bin/nvcc"
BASE = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]

def build(name, flags):
    d = os.path.join(OUT, name + "_obj"); os.makedirs(d, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    def cc(src):
        o = os.path.join(d, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([NVCC] + BASE + ["-Xptxas", "-v"] + flags.split() + ["-c", src, "-o", o], capture_output=True, text=True)
        if r.returncode: sys.stderr.write(r.stderr); raise SystemExit(1)
        open(o[:-2] + ".log", "w").write(r.stderr)
        return o
    with concurrent.futures.ThreadPoolExecutor(8) as ex: objs = list(ex.map(cc, srcs))
    subprocess.check_call([NVCC] + BASE[:2] + ["-shared", "-o", os.path.join(OUT, name + ".so")] + objs + ["-cudart", "static"])
    print("built", name, flags)

for a in sys.argv[1:]:
    n, f = a.split("=", 1); build(n, f)
