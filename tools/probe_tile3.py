"""Few tile-kernel launches for ncu: cold, warm, warm."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_vs_map(map_points=1_000_000)
ctx = plo.Context(0)
ctx.set_target(pair.target)
ctx.set_source(pair.source)
for _ in range(3):
    st = ctx.project(np.eye(4))
print(st["n_pairs"])
