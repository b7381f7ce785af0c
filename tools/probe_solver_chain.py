"""One registration with the config.json default solver chain (RANSAC -> DRPM) at the north-star size (for ncu)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import plo_b200 as plo
final = int(sys.argv[1]) if len(sys.argv) > 1 else 2
solver = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pair = plo.synth.workloads.hdl64_vs_map(map_points=1_000_000)
ctx = plo.Context(0, plo.default_params(solver=solver, ransac_final=final))
ctx.set_target(pair.target); ctx.set_source(pair.source)
T, st = ctx.register()
print(st["iters"], st["status_name"], ctx.last_timings())
