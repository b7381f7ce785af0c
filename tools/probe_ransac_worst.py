"""Worst case of the RANSAC front (the 95 % exit never fires): ransac_max_iterations hypotheses over the pairs of the
north-star frame, timed around plo_solve_ransac (stepped call, synchronous).  usage: python tools/probe_ransac_worst.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_vs_map()
for maxit, pct in ((5000, 0.95), (296, 1.0), (5000, 1.0)):
    ctx = plo.Context(0)
    ctx.set_params(plo.default_params(solver=2, ransac_final=1, ransac_max_iterations=maxit, ransac_min_inliers_percentage=pct,
                                      ransac_distance_threshold=0.05 if pct == 1.0 else 0.8))
    ctx.set_target(pair.target); ctx.set_source(pair.source)
    ctx.project(np.eye(4))
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); delta, info = ctx.solve_ransac(); ts.append(1e3 * (time.perf_counter() - t0))
    print(f"max_iterations {maxit} min_inliers {pct}: hypotheses {info['hypotheses']} inliers {info['inliers']} | solve_ransac ms {min(ts):.2f}", flush=True)
