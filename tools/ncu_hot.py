"""Hot-code footprint: SASS instructions executed at least FRAC x the hottest count, bytes per source line range.
python tools/ncu_hot.py report.ncu-rep [frac=0.02]"""
import csv, subprocess, sys
rep = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
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
        if r[2].startswith("0x"):
            items.append((int(r[2], 16), fname, line, int(r[7])))
# the same SASS address appears once; total static size:
addrs = {}
for a, f, l, n in items: addrs[a] = (f, l, n)
mx = max(n for _, _, n in addrs.values())
hot = {a: v for a, v in addrs.items() if v[2] >= frac * mx}
print("static SASS %d instr (%.1f KB); executed >= %.0f%% of max: %d instr (%.1f KB); executed at all: %d (%.1f KB)" % (
    len(addrs), len(addrs) * 16 / 1024, frac * 100, len(hot), len(hot) * 16 / 1024,
    sum(1 for v in addrs.values() if v[2] > 0), sum(1 for v in addrs.values() if v[2] > 0) * 16 / 1024))
by = {}
for a, (f, l, n) in hot.items():
    k = (f, (l // 25) * 25 if l else 0)
    by[k] = by.get(k, 0) + 1
for k, c in sorted(by.items(), key=lambda kv: -kv[1])[:25]:
    print("%-22s lines %4d-%4d  %5d instr %6.1f KB" % (k[0], k[1], k[1] + 24, c, c * 16 / 1024))
