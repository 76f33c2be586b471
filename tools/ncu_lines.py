"""Per-CUDA-line digest of an `ncu --page source --csv --print-source cuda,sass` export.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > prof_cs.csv
    python tools/ncu_lines.py prof_cs.csv [top_n]

For every source line: executed warp instructions, stall samples, shared-memory wavefronts (and the excess over
the ideal count), and the number of SASS instructions generated for it (static size / executed size: what the
instruction cache has to hold).
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path, newline="")))
    cur_file = "?"
    hdr = None
    lines = {}
    sass_static = defaultdict(int)
    sass_hot = defaultdict(int)
    cur_line = None
    max_exec = 0
    raw = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            idx = {h: i for i, h in enumerate(hdr)}
            # two "Source" columns: 1 = cuda text, 3 = sass text
            continue
        if r[0] == "Function Name" or hdr is None:
            continue
        if r[0] != "-" and r[0] != "":
            # a CUDA line row (aggregated)
            try:
                ln = int(r[0])
            except ValueError:
                continue
            cur_line = (cur_file, ln)

            def num(name):
                try:
                    return float(r[idx[name]])
                except (ValueError, KeyError, IndexError):
                    return 0.0
            lines[cur_line] = dict(text=r[1].strip(), inst=num("Instructions Executed"), samples=num("# Samples"),
                                   wf=num("L1 Wavefronts Shared"), wfx=num("L1 Wavefronts Shared Excessive"),
                                   noinst=num("stall_no_inst"), ssb=num("stall_short_sb"), wait=num("stall_wait"),
                                   bar=num("stall_barrier"), lsb=num("stall_long_sb"))
        else:
            # a SASS row under the current CUDA line
            try:
                ex = float(r[idx["Instructions Executed"]])
            except (ValueError, KeyError, IndexError):
                ex = 0.0
            raw.append((cur_line, ex, r[idx["Address"]] if "Address" in idx else len(raw), r))
            max_exec = max(max_exec, ex)
    # An instruction inlined from a header is listed under the header line AND under its call sites: keep one
    # listing per SASS address, preferring the line in the .cu file (the call site), and rebuild the per-line sums
    prefer = sys.argv[3] if len(sys.argv) > 3 else ".cu"
    by_addr = {}
    for ln, ex, addr, r in raw:
        if addr not in by_addr or (prefer in ln[0] and prefer not in by_addr[addr][0][0]):
            by_addr[addr] = (ln, ex, r)
    for v in lines.values():
        v["inst"] = v["samples"] = v["wf"] = v["wfx"] = v["noinst"] = v["ssb"] = v["wait"] = v["bar"] = v["lsb"] = 0.0
    def fnum(r, name):
        try:
            return float(r[idx[name]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    raw2 = []
    for addr, (ln, ex, r) in by_addr.items():
        raw2.append((ln, ex))
        v = lines[ln]
        v["inst"] += ex
        v["samples"] += fnum(r, "# Samples")
        v["wf"] += fnum(r, "L1 Wavefronts Shared")
        v["wfx"] += fnum(r, "L1 Wavefronts Shared Excessive")
        v["noinst"] += fnum(r, "stall_no_inst"); v["ssb"] += fnum(r, "stall_short_sb"); v["wait"] += fnum(r, "stall_wait")
        v["bar"] += fnum(r, "stall_barrier"); v["lsb"] += fnum(r, "stall_long_sb")
    for ln, ex in raw2:
        sass_static[ln] += 1
        if ex >= 0.02 * max_exec:
            sass_hot[ln] += 1
    tot_i = sum(v["inst"] for v in lines.values()) or 1.0
    tot_s = sum(v["samples"] for v in lines.values()) or 1.0
    tot_w = sum(v["wf"] for v in lines.values()) or 1.0
    print("total executed warp instructions %.0f, samples %.0f, shared wavefronts %.0f (excess %.0f)" %
          (tot_i, tot_s, tot_w, sum(v["wfx"] for v in lines.values())))
    print("static SASS %d instr (%.1f KB); executed >= 2%% of max: %d instr (%.1f KB)" %
          (sum(sass_static.values()), sum(sass_static.values()) * 16 / 1024.0, sum(sass_hot.values()),
           sum(sass_hot.values()) * 16 / 1024.0))
    print("%-22s %5s %6s %6s %6s %6s %5s %5s  %s" % ("file", "line", "inst%", "smpl%", "wf%", "wfx%", "sass", "hot", "source"))
    for key, v in sorted(lines.items(), key=lambda kv: -kv[1]["inst"])[:top]:
        print("%-22s %5d %6.2f %6.2f %6.2f %6.1f %5d %5d  %s" %
              (key[0][:22], key[1], 100 * v["inst"] / tot_i, 100 * v["samples"] / tot_s, 100 * v["wf"] / tot_w,
               100 * v["wfx"] / v["wf"] if v["wf"] else 0.0, sass_static[key], sass_hot[key], v["text"][:90]))
    print("\nstall samples by reason (sum over lines): no_inst %.0f short_sb %.0f wait %.0f barrier %.0f long_sb %.0f" %
          tuple(sum(v[k] for v in lines.values()) for k in ("noinst", "ssb", "wait", "bar", "lsb")))
    # hot code by file region
    print("\nhot SASS by 25-line block:")
    blocks = defaultdict(int)
    for (f, ln), c in sass_hot.items():
        blocks[(f, ln // 25 * 25)] += c
    for (f, b), c in sorted(blocks.items(), key=lambda kv: -kv[1])[:20]:
        print("  %-22s lines %4d-%4d  %5d instr %6.1f KB" % (f, b, b + 24, c, c * 16 / 1024.0))


if __name__ == "__main__":
    main()
