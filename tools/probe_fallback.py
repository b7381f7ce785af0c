import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_vs_map(map_points=1_000_000)
ctx = plo.Context(0)
ctx.set_target(pair.target); ctx.set_source(pair.source)
ctx.project(np.eye(4), hooks=True)
ss = ctx.search_stats(); nb = ctx.neighbors()
c = ss[:, 2]
ovf = (c < 0) & (c > -1000000); und = c <= -1000000; ok = c >= 0
print("overflow", ovf.mean(), "undercount", und.mean(), "ok", ok.mean())
kth = np.sqrt(nb["nn_d2"][:, 19])
rng = np.linalg.norm(pair.source[:, 0:3], axis=1)
for name, msk in (("ok", ok), ("overflow", ovf), ("under", und)):
    if msk.any():
        print(name, "kth dist p10/50/90", np.percentile(kth[msk & np.isfinite(kth)], [10, 50, 90]) if (msk & np.isfinite(kth)).any() else None,
              "range p10/50/90", np.percentile(rng[msk], [10, 50, 90]), "z mean", pair.source[msk, 2].mean())
