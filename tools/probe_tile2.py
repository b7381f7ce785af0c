import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_vs_map(map_points=1_000_000)
ctx = plo.Context(0)
ctx.set_target(pair.target)
for it in (1, 2, 3, 4, 7):
    ctx.set_params(plo.default_params(iterations=it))
    for rep in range(2):
        ctx.set_source(pair.source)
        T, rs = ctx.register()
    print(it, rs["iters"], rs["status_name"], ctx.last_timings())
ctx.set_source(pair.source)
st = ctx.project(np.eye(4), hooks=True)
ss = ctx.search_stats()
print("col2 raw:", ss[:8, 2], ss[:, 2].max(), "col0", ss[:8, 0])
