"""BASELINE config 4 (HDL-64 frame vs a 5 M-point map: 160 MB of sorted points + normals, larger than L2):
index build and registration times, per-launch projection times.  Used for the HBM-bound view of the roofline."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_vs_dense_map()
ctx = plo.Context(0)
idx, reg = [], []
for i in range(5):
    ctx.set_target(pair.target); ctx.set_source(pair.source); T, rs = ctx.register()
    t = ctx.last_timings(); idx.append(t["ms_index_build"]); reg.append(t["ms_register"])
ctx.set_profiling(True)
ctx.set_target(pair.target); ctx.set_source(pair.source); T, rs = ctx.register()
each = ctx.last_project_times()
ctx.set_profiling(False)
n_t, n_s = int(pair.target.shape[0]), int(pair.source.shape[0])
pairs = int(rs["pairs"])
alg = pairs * 504 + (n_s - pairs) * 48
out = dict(workload="cfg-4: HDL-64 frame vs 5 M-point map", map_points=n_t, source_points=n_s, iters=int(rs["iters"]), pairs=pairs,
           ms_index_build=float(np.median(idx[1:])), index_build_GBps_algorithmic=52.0 * n_t / (np.median(idx[1:]) * 1e-3) / 1e9,
           ms_register=float(np.median(reg[1:])), ms_project_each=[round(float(x), 4) for x in each],
           k_project_GBps_algorithmic=[round(alg / (float(x) * 1e-3) / 1e9, 1) for x in each])
print(json.dumps(out))
