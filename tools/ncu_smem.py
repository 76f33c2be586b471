"""Shared-memory wavefronts per CUDA source line: python tools/ncu_smem.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; agg = {}; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        iw = hdr.index("L1 Wavefronts Shared"); ie = hdr.index("L1 Wavefronts Shared Excessive"); ii = hdr.index("L1 Wavefronts Shared Ideal")
        continue
    if hdr and len(r) > 8 and r[0] not in ("", "Line No"):
        try: line = int(r[0]); w = int(r[iw]); e = int(r[ie]); inst = int(r[7])
        except ValueError: continue
        a = agg.setdefault((fname, line), [0, 0, 0, r[1]]); a[0] += w; a[1] += e; a[2] += inst
tw = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
print("total shared wavefronts %d excessive %d (%.1f%%)" % (tw, te, 100.0 * te / max(tw, 1)))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-20s %5d  wf %5.1f%%  excess %5.1f%% of line  %s" % (k[0], k[1], 100 * a[0] / tw, 100 * a[1] / max(a[0], 1), a[3].strip()[:80]))
