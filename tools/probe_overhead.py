"""Per-iteration overhead of the resident loop: tiny clouds, convergence thresholds at zero (every registration runs
all `iterations`), graph loop against the enqueue-all loop.  usage: python tools/probe_overhead.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_pair(max_source=2000, max_target=20000)
for iters in (10, 30):
    for no_graph in (0, 1):
        ctx = plo.Context(0, plo.default_params(iterations=iters, delta_dist_threshold=0.0, delta_angle_threshold=0.0))
        ctx.set_tuning("no_graph", no_graph)
        ms = []
        for i in range(12):
            ctx.set_target(pair.target); ctx.set_source(pair.source)
            T, rs = ctx.register()
            if i >= 4: ms.append(ctx.last_timings()["ms_register"])
        print(f"iterations {iters} no_graph {no_graph}: status {rs['status_name']} iters {rs['iters']} register {np.median(ms):.4f} ms -> {1e3 * np.median(ms) / iters:.2f} us per iteration", flush=True)
