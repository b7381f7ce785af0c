"""Diagnostics of k_project on the north-star workload: traversal statistics per projection of one
registration (hooks instantiation), then plain registrations (for an ncu capture with PLO_NO_GRAPH=1)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_vs_map()
ctx = plo.Context(0)
if "stats" in sys.argv:
    ctx.set_target(pair.target); ctx.set_source(pair.source)
    T = np.eye(4)
    for it in range(7):
        ctx.project(T, hooks=True)
        s = ctx.search_stats()
        hit = (s[:, 2] % 100000) >= 50000
        cand = s[:, 2] % 50000
        print(f"it{it} cache hits {hit.mean():.3f} leaves {s[:,0].mean():.2f} nodes {s[:,1].mean():.2f} cand {cand.mean():.1f} shrinks {(s[:,2]//100000).mean():.3f} "
              f"p90 leaves {np.percentile(s[:,0],90):.0f} cand p90 {np.percentile(cand,90):.0f}", flush=True)
        d = ctx.solve_wls()
        d = d[0] if isinstance(d, tuple) else d
        T = np.asarray(d).reshape(4, 4) @ T
for _ in range(2):
    ctx.set_target(pair.target); ctx.set_source(pair.source); T, rs = ctx.register()
print(rs["iters"], ctx.last_timings())
