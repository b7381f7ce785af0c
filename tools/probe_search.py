"""Traversal statistics + one projection (used for ncu captures)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo

mp = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pair = plo.synth.workloads.hdl64_vs_map(map_points=mp)
ctx = plo.Context(0)
ctx.set_target(pair.target)
ctx.set_source(pair.source)
st = ctx.project(np.eye(4), hooks=True)
ss = ctx.search_stats()
print("pairs", st["n_pairs"], "counters", st["counters"])
for i, nm in enumerate(["leaves", "nodes", "inserts"]):
    v = ss[:, i]
    print(f"{nm:8s} mean {v.mean():8.2f} p50 {np.percentile(v,50):6.0f} p90 {np.percentile(v,90):6.0f} p99 {np.percentile(v,99):6.0f} max {v.max()}")
T, rs = ctx.register()
print("register", rs["iters"], rs["status_name"], ctx.last_timings())
for T_ in (np.eye(4), T):
    print("k_project ms", ctx.time_project_kernel(T_, reps))
fb = (ss[:, 2] < 0).mean()
print("fallback fraction at iteration 0:", fb, " collected mean (non-fallback):", ss[ss[:, 2] >= 0, 2].mean() if (ss[:, 2] >= 0).any() else None)
ctx.set_target(pair.target); ctx.set_source(pair.source)
import time
for _ in range(3):
    ctx.set_target(pair.target); ctx.set_source(pair.source); T, rs = ctx.register(); print(ctx.last_timings())
ctx.project(T, hooks=True); ss2 = ctx.search_stats()
print("converged-pose stats: leaves %.2f nodes %.2f cand %.2f fallback %.4f" % (ss2[:,0].mean(), ss2[:,1].mean(), ss2[ss2[:,2]>=0,2].mean(), (ss2[:,2]<0).mean()))
