"""Summarise ncu outputs into small text files for profiles/ (the .ncu-rep files stay in gpurun_out/).
usage: ncu_summary.py launches <launches.csv> | full <report.ncu-rep>"""
import collections, csv, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        us = v if u.startswith("us") else (v * 1e3 if u.startswith("ms") else v / 1e3)
        a = agg.setdefault(row["Kernel Name"].split("(")[0][-48:], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':50s} {'launches':>8s} {'total_us':>12s} {'mean_us':>10s} {'share':>7s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:50s} {a[0]:8d} {a[1]:12.1f} {a[1]/a[0]:10.1f} {a[1]/tot:7.3f}")
    print(f"{'TOTAL':50s} {sum(a[0] for a in agg.values()):8d} {tot:12.1f}")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active"]

def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=== kernel:", r[hdr.index("Kernel Name")][:100])
        for w in WANT:
            if w in hdr:
                print(f"  {w:72s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
        st = [(h, r[i]) for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
        print("  top stall reasons (warps stalled per issue-active cycle):")
        for h, v in sorted(st, key=lambda x: -float(x[1].replace(",", "") or 0))[:6]:
            print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v}")

if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
