"""Key metrics of `ncu --page raw --csv` exports, one column per file.  usage: ncu_raw.py a_raw.csv b_raw.csv ..."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum"]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["long_scoreboard", "no_instruction", "wait", "short_scoreboard", "membar", "lg_throttle", "math_pipe_throttle", "barrier",
          "branch_resolving", "not_selected", "dispatch_stall", "mio_throttle", "imc_miss", "drain", "sleeping", "tex_throttle", "selected"]
cols = []
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    cols.append((f, dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))))
print(" " * 58 + "".join(f"{c[0].split('/')[-1][:22]:>24s}" for c in cols))
for k in ["Kernel Name"] + KEYS + [STALL % s for s in STALLS]:
    if any(k in c[1] for c in cols):
        print(f"{k.replace('smsp__average_warps_issue_stalled_','stall ').replace('_per_issue_active.ratio',''):58s}" +
              "".join(f"{c[1].get(k, '-')[:22]:>24s}" for c in cols) + "  " + cols[0][2].get(k, ""))
