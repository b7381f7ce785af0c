import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_vs_map(map_points=int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000)
ctx = plo.Context(0)
for _ in range(2):
    ctx.set_target(pair.target); ctx.set_source(pair.source); T, rs = ctx.register()
print(rs["iters"], ctx.last_timings())
