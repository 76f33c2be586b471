#!/usr/bin/env python
"""Time the fused gas-opacity kernel for every scratch/variants/*.so (one subprocess per library)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path.insert(0, %r)
from archnemesis_dist_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
import runpy
sys.argv = ["time_overlap.py"] + sys.argv[2:]
runpy.run_path(os.path.join(%r, "tools", "time_overlap.py"), run_name="__main__")
''' % (ROOT, ROOT)
for so in sorted(glob.glob(os.path.join(ROOT, "scratch", "variants", "*.so"))):
    r = subprocess.run([sys.executable, "-c", CODE, so] + sys.argv[1:], capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if "parallel" in l]
    print(os.path.basename(so), " | ".join(l.split("grad=")[1] for l in lines) if lines else r.stderr[-400:])
