"""Key metrics of `ncu --page raw --csv` exports (one kernel launch per file), one column per file.
usage: ncu_raw.py a_raw.csv b_raw.csv ...            table on stdout
       ncu_raw.py --json queries name=a_raw.csv ...  per-launch figures for profiles/roofline_traffic.json"""
import csv, json, sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["long_scoreboard", "no_instruction", "wait", "short_scoreboard", "membar", "lg_throttle", "math_pipe_throttle", "barrier",
          "branch_resolving", "not_selected", "dispatch_stall", "mio_throttle", "imc_miss", "drain", "selected"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}


def load(path):
    rows = list(csv.reader(open(path)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def val(d, u, k):
    """metric in base units (bytes, microseconds, plain numbers)"""
    if k not in d or d[k] in ("", "n/a"):
        return None
    return float(d[k].replace(",", "")) * SCALE.get(u.get(k, ""), 1.0)


if sys.argv[1] == "--json":
    nq = float(sys.argv[2])
    out = {}
    for spec in sys.argv[3:]:
        name, path = spec.split("=", 1)
        d, u = load(path)
        inst = val(d, u, "smsp__inst_executed.sum")
        issue = val(d, u, "smsp__issue_active.avg.pct_of_peak_sustained_active") or val(d, u, "smsp__issue_active.avg.pct")
        out[name] = {"dram_bytes_per_launch": int(val(d, u, "dram__bytes_read.sum") + val(d, u, "dram__bytes_write.sum")),
                     "l2_bytes_per_launch": int(32 * val(d, u, "lts__t_sectors.sum")), "inst_per_query": round(inst / nq, 1),
                     "issue_active_pct": None if issue is None else round(issue, 1),
                     "l1_hit_pct": round(val(d, u, "l1tex__t_sector_hit_rate.pct"), 1), "l2_hit_pct": round(val(d, u, "lts__t_sector_hit_rate.pct"), 1),
                     "warps_active_pct": round(val(d, u, "sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
                     "ncu_duration_us_cold_cache": round(val(d, u, "gpu__time_duration.sum"), 1), "capture": path.split("/")[-1]}
    print(json.dumps({"k_project": out, "note": "ncu --set full --clock-control none, one launch per regime, caches cold and launches "
                                               "serialised under the profiler (tools/ncu_kproj.sh); bytes and instructions per launch"}, indent=1))
    sys.exit(0)

if sys.argv[1] == "--kernels":
    # every launch of a multi-kernel capture: duration, DRAM / L2 bytes, achieved DRAM GB/s, issue utilisation, occupancy
    print(f"{'kernel':44s} {'n':>3s} {'us':>8s} {'dram MB':>8s} {'L2 MB':>8s} {'dram GB/s':>9s} {'issue%':>6s} {'warps%':>6s} {'L2hit%':>6s} {'regs':>4s} {'grid':>6s}")
    for path in sys.argv[2:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], dict(zip(rows[0], rows[1]))
        agg = {}
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:44]
            us = val(d, units, "gpu__time_duration.sum")
            dr = (val(d, units, "dram__bytes_read.sum") or 0) + (val(d, units, "dram__bytes_write.sum") or 0)
            a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, d.get("launch__registers_per_thread", ""), d.get("launch__grid_size", "")])
            a[0] += 1; a[1] += us; a[2] += dr; a[3] += 32 * (val(d, units, "lts__t_sectors.sum") or 0)
            a[4] += val(d, units, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0
            a[5] += val(d, units, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0
            a[6] += val(d, units, "lts__t_sector_hit_rate.pct") or 0
        tot = sum(a[1] for a in agg.values())
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            n = a[0]
            print(f"{name:44s} {n:3d} {a[1]:8.1f} {a[2] / 1e6:8.1f} {a[3] / 1e6:8.1f} {a[2] / max(a[1], 1e-9) / 1e3:9.1f} {a[4] / n:6.1f} {a[5] / n:6.1f} {a[6] / n:6.1f} {a[7]:>4s} {a[8]:>6s}")
        print(f"{'TOTAL ' + path.split('/')[-1]:44s} {'':3s} {tot:8.1f}")
    sys.exit(0)

cols = [(f,) + load(f) for f in sys.argv[1:]]
print(" " * 50 + "".join(f"{c[0].split('/')[-1].replace('_raw.csv', '')[:20]:>22s}" for c in cols))
for k in KEYS + [STALL % s for s in STALLS]:
    vals = [val(c[1], c[2], k) for c in cols]
    if any(v is not None for v in vals):
        name = k.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", "")
        print(f"{name[:50]:50s}" + "".join(f"{'-' if v is None else format(v, '.4g'):>22s}" for v in vals))
