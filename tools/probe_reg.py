import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_pair(max_source=5000, max_target=40000)
for kw in ({}, {"solver": 1}, {"solver": 2, "ransac_final": 1}, {"solver": 2, "ransac_final": 2}):
    ctx = plo.Context(0, plo.default_params(**kw))
    ctx.set_target(pair.target); ctx.set_source(pair.source)
    st = ctx.project(np.eye(4)); print(kw, "project", st["n_pairs"], st["counters"])
    T, rs = ctx.register(); print("  register", rs["status_name"], rs["iters"], rs["pairs"], rs["rank"], np.round(T[:3,3],4))
    os.environ["PLO_NO_GRAPH"]="1"
