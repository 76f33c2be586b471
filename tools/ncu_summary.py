"""Key metrics of one or more ncu reports side by side: python tools/ncu_summary.py a.ncu-rep b.ncu-rep ..."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shfl.sum" , "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
cols = []
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, v = rows[0], rows[-1]
    d = dict(zip(h, v))
    stalls = {k: d[k] for k in h if "issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k}
    cols.append((d, stalls))
for k in KEYS:
    print("%-72s" % k, "  ".join("%14s" % c[0].get(k, "-") for c in cols))
names = sorted(cols[0][1], key=lambda k: -float(cols[-1][1].get(k, 0) or 0))
for k in names[:12]:
    print("%-72s" % k.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""),
          "  ".join("%14s" % c[1].get(k, "-") for c in cols))
