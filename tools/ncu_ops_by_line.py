"""Executed warp instructions by (opcode, CUDA line) from an `ncu --page source --csv --print-source cuda,sass` export:
    python tools/ncu_ops_by_line.py prof_cs.csv [top_n] [opcode-prefix ...]"""
import csv
import re
import sys
from collections import Counter


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    prefixes = tuple(sys.argv[3:])
    seen, cur, tot = set(), None, 0
    by_op, by_op_line = Counter(), Counter()
    for r in rows:
        if len(r) < 8 or r[0] == "Line No":
            continue
        if r[0].isdigit():
            cur = int(r[0])
            continue
        a = r[2]
        if not a.startswith("0x") or a in seen:
            continue
        seen.add(a)
        try:
            n = int(float(r[7]))
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
        if not m:
            continue
        op = m.group(2)
        tot += n
        by_op[op.split(".")[0]] += n
        if not prefixes or op.startswith(prefixes):
            by_op_line[(op, cur)] += n
    print("executed warp instructions", tot)
    print("  ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in by_op.most_common(24)))
    for (op, line), n in by_op_line.most_common(top):
        print("%-24s line %5s %6.2f%%" % (op, line, 100.0 * n / tot))


if __name__ == "__main__":
    main()
