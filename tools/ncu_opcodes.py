"""Executed warp instructions by opcode from an `ncu --page source --csv --print-source cuda,sass` export."""
import csv
import re
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    hdr, idx = None, {}
    ops = defaultdict(int)
    tot = 0
    seen = set()
    for r in rows:
        if not r:
            continue
        if r[0] == "Line No":
            hdr = r
            idx = {h: i for i, h in enumerate(hdr)}
            continue
        if hdr is None or "Address" not in idx or len(r) <= idx["Address"]:
            continue
        a = r[idx["Address"]]
        if not a or a in seen:
            continue
        seen.add(a)
        src = [i for i, h in enumerate(hdr) if h == "Source"]
        sass = r[src[-1]]
        try:
            n = int(float(r[idx["# Warp Instructions Executed"]]))
        except (KeyError, ValueError):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
        op = m.group(2).split(".")[0] if m else sass[:10]
        ops[op] += n
        tot += n
    print("executed warp instructions", tot)
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
        print("%-12s %6.2f%%" % (k, 100 * v / tot))


if __name__ == "__main__":
    main()
