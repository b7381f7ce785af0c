"""PCA-normal mode (get_normals.enabled = false) on the cfg-1 pair: timing of set_target + first projection, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_pair()
ctx = plo.Context(0, plo.default_params(is_get_normals=0))
for _ in range(2):
    ctx.set_target(pair.target); ctx.set_source(pair.source)
    T, rs = ctx.register()
print(rs["iters"], rs["pairs"], ctx.last_timings())
