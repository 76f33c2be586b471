"""Dump the SASS of an ncu report in address order with executed counts and the CUDA line it belongs to:
python tools/ncu_sass.py report.ncu-rep lo hi   (only instructions whose source line is in [lo,hi] of koverlap_impl.cuh)"""
import csv, subprocess, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; line = None; items = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": continue
    if len(r) > 8:
        if r[0] != "":
            try: line = int(r[0])
            except ValueError: line = None
            continue
        if r[2].startswith("0x") and fname == "koverlap_impl.cuh" and line is not None and lo <= line <= hi:
            items.append((int(r[2], 16), line, r[3].strip(), int(r[7]), int(r[6])))
items.sort()
base = items[0][0] if items else 0
for a, l, sass, inst, smp in items:
    print("%6x L%-4d exec %9d smp %5d  %s" % (a - base, l, inst, smp, sass))
