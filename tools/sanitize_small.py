"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck): every kernel runs once."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_pair(max_source=1500, max_target=9000)
tgt = pair.target.copy(); tgt[5, 0] = np.nan
for kw in ({}, {"is_get_normals": 0, "weight_mode": 1}):
    ctx = plo.Context(0, plo.default_params(**kw))
    ctx.set_target(tgt); ctx.set_source(pair.source)
    ctx.project(np.eye(4), hooks=True); ctx.neighbors(); ctx.pairs(); ctx.solve_wls(); ctx.target_normals()
    T, st = ctx.register()
    Tb, sb = ctx.register_batch([pair.source, pair.source[:100]], [tgt, tgt])
    print(kw, st["iters"], st["status_name"], sb[1]["status_name"])
    ctx.close()
# the other solver chains: trimmed LS, RANSAC (several hypotheses, exit out of reach) -> LS / weighted LS / DRPM
for kw in ({"solver": 1}, {"solver": 2, "ransac_final": 0}, {"solver": 2, "ransac_final": 1, "ransac_min_inliers_percentage": 1.0, "ransac_max_iterations": 40},
           {"solver": 2, "ransac_final": 2}):
    ctx = plo.Context(0, plo.default_params(**kw))
    ctx.set_target(tgt); ctx.set_source(pair.source)
    T, st = ctx.register()
    print(kw, st["iters"], st["status_name"])
    ctx.close()
print("sanitize case done")
