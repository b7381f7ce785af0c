"""Tile-kernel traversal statistics at the north-star size: cold projection, warm projections, kernel time."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo

mp = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pair = plo.synth.workloads.hdl64_vs_map(map_points=mp)
ctx = plo.Context(0)
ctx.set_target(pair.target)
ctx.set_source(pair.source)

def show(tag):
    ss = ctx.search_stats()
    c = ss[:, 2] % 100000
    o = ss[:, 2] // 100000
    print(f"{tag}: leaves/tile {ss[:,0].mean():.1f} (p90 {np.percentile(ss[:,0],90):.0f} max {ss[:,0].max()}) nodes/tile {ss[:,1].mean():.1f} "
          f"cand/lane {c.mean():.1f} (p90 {np.percentile(c,90):.0f} max {c.max()}) overflow flushes/lane {o.mean():.3f}")

st = ctx.project(np.eye(4), hooks=True); show("cold @I")
st = ctx.project(np.eye(4), hooks=True); show("warm same pose")
T, rs = ctx.register()
print("register", rs["iters"], rs["status_name"], ctx.last_timings())
ctx.project(T, hooks=True); show("after register, @T")
ctx.project(T, hooks=True); show("warm @T")
for T_ in (np.eye(4), T):
    print("k_project ms", ctx.time_project_kernel(T_, 3))
