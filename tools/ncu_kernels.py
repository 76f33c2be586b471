"""Key metrics of every kernel in one ncu report (`ncu --set full`): python tools/ncu_kernels.py report.ncu-rep"""
import csv, subprocess, sys
KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
for v in rows[2:]:
    d = dict(zip(h, v))
    u = dict(zip(h, units))
    print("== %s" % d.get("Kernel Name", "?"))
    for k in KEYS:
        if k in d:
            print("%-76s %s %s" % (k, d[k], u.get(k, "")))
    stalls = sorted(((float(d[k] or 0), k) for k in h if "issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k),
                    reverse=True)
    for val, k in stalls[:6]:
        print("%-76s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", " (per issue)"), val))
    print()
