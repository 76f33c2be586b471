"""Phase digest of the fast overlap kernel from an `ncu --page source --csv --print-source cuda,sass` export.

    python tools/ncu_phases.py prof_cs.csv NFOLDS NSORTED

Lines of koverlap_fast.cu are grouped by the `// @phase name` markers in the source; inlined intrinsics are
attributed to their call site.  Prints executed warp instructions, stall samples and shared-memory wavefronts
per phase, per fold.
"""
import csv
import os
import re
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "archnemesis_dist_b200", "csrc", "koverlap_fast.cu")


def phases_from_source():
    marks = []
    for n, line in enumerate(open(SRC), 1):
        m = re.search(r"//\s*@phase\s+(\S+)", line)
        if m:
            marks.append((n, m.group(1)))
    return marks


def main():
    path, nfold, nsorted = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
    marks = phases_from_source()

    def phase_of(f, ln):
        if f != "koverlap_fast.cu":
            return "kinterp" if f == "kinterp.cuh" else "hdr:" + f
        name = "top"
        for n, p in marks:
            if n <= ln:
                name = p
        return name
    rows = list(csv.reader(open(path, newline="")))
    cur = None
    f = None
    idx = None
    by_addr = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            f = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            idx = {h: i for i, h in enumerate(r)}
            continue
        if r[0] == "Function Name":
            continue
        if r[0] not in ("-", ""):
            try:
                cur = (f, int(r[0]))
            except ValueError:
                cur = None
            continue
        if cur is None:
            continue
        addr = r[idx["Address"]]

        def num(name):
            try:
                return float(r[idx[name]])
            except (ValueError, KeyError, IndexError):
                return 0.0
        rec = (cur, num("Instructions Executed"), num("# Samples"), num("L1 Wavefronts Shared"),
               num("stall_short_sb"), num("stall_wait"), num("stall_no_inst"), num("stall_long_sb"), num("stall_barrier"),
               num("stall_math"), num("stall_not_selected"))
        if addr not in by_addr or (".cu" in cur[0] and ".cu" not in by_addr[addr][0][0]):
            by_addr[addr] = rec
    agg = defaultdict(lambda: [0.0] * 10)
    for rec in by_addr.values():
        p = phase_of(*rec[0])
        for i in range(10):
            agg[p][i] += rec[1 + i]
    ti = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values())
    tw = sum(v[2] for v in agg.values())
    print("total: %.0f warp instructions (%.0f per fold), %.0f samples, %.0f shared wavefronts (%.0f per fold)" %
          (ti, ti / nfold, ts, tw, tw / nfold))
    print("%-14s %6s %8s %8s %6s %6s %7s | stall samples: %6s %6s %6s %6s %6s %6s %6s" %
          ("phase", "inst%", "/fold", "/sorted", "smpl%", "wf%", "smp/ki", "ssb", "wait", "noinst", "lsb", "bar", "math", "notsel"))
    for p, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-14s %6.2f %8.1f %8.1f %6.2f %6.2f %7.2f | %20.0f %6.0f %6.0f %6.0f %6.0f %6.0f %6.0f" %
              (p[:14], 100 * v[0] / ti, v[0] / nfold, v[0] / nsorted, 100 * v[1] / ts, 100 * v[2] / tw if tw else 0,
               1000 * v[1] / v[0] if v[0] else 0, v[3], v[4], v[5], v[6], v[7], v[8], v[9]))


if __name__ == "__main__":
    main()
