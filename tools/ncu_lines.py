"""Per-source-line summary of an ncu report: python tools/ncu_lines.py report.ncu-rep [top]
Uses `ncu --page source --print-source cuda,sass`; prints executed warp instructions and stall
samples per CUDA source line (needs -lineinfo)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; agg = {}; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            line = int(r[0]); samples = int(r[6]); inst = int(r[7])
        except ValueError:
            continue
        k = (fname, line)
        a = agg.setdefault(k, [0, 0, r[1]])
        a[0] += inst; a[1] += samples
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total inst %d samples %d" % (ti, ts))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-20s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (k[0], k[1], 100 * a[0] / ti, 100 * a[1] / ts, a[2].strip()[:90]))
if len(sys.argv) > 3:
    # ranges "name:lo-hi,..." over koverlap_impl.cuh
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        i = sum(a[0] for k, a in agg.items() if k[0] == "koverlap_impl.cuh" and lo <= k[1] <= hi)
        s_ = sum(a[1] for k, a in agg.items() if k[0] == "koverlap_impl.cuh" and lo <= k[1] <= hi)
        print("%-12s inst %5.1f%% samples %5.1f%%" % (name, 100 * i / ti, 100 * s_ / ts))
    i = sum(a[0] for k, a in agg.items() if k[0] != "koverlap_impl.cuh"); s_ = sum(a[1] for k, a in agg.items() if k[0] != "koverlap_impl.cuh")
    print("%-12s inst %5.1f%% samples %5.1f%%" % ("other files", 100 * i / ti, 100 * s_ / ts))
