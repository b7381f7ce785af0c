import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
import numpy as np, plo_b200 as plo, oracle_ctypes as oc
# pathological: hundreds of coincident / equidistant points around the query
rng = np.random.default_rng(1)
tgt = np.zeros((3000, 12), np.float32); tgt[:, 6] = 1
tgt[:300, 0:3] = [1, 1, 1]                       # 300 duplicates
th = rng.uniform(0, 2*np.pi, 700)
tgt[300:1000, 0] = 1 + 0.5*np.cos(th); tgt[300:1000, 1] = 1 + 0.5*np.sin(th); tgt[300:1000, 2] = 1   # ring of 700 near-equidistant
tgt[1000:, 0:3] = rng.uniform(-3, 5, size=(2000, 3))
src = np.zeros((4, 12), np.float32); src[:, 6] = 1
src[0, 0:3] = [1, 1, 1]; src[1, 0:3] = [1, 1, 1.001]; src[2, 0:3] = [1.2, 1, 1]; src[3, 0:3] = [0, 0, 0]
for kw in ({}, {"search_number": 32}, {"search_number": 3}):
    ctx = plo.Context(0, plo.default_params(**kw)); orc = oc.Oracle(oc.default_params(**kw))
    for o in (ctx, orc): o.set_target(tgt); o.set_source(src)
    ctx.project(np.eye(4), hooks=True); g = ctx.neighbors(); ss = ctx.search_stats()
    o = orc.project(np.eye(4), hooks=True)
    print(kw, "idx equal", np.array_equal(g["nn_idx"], o["nn_idx"]), "d2 equal", np.array_equal(g["nn_d2"], o["nn_d2"]), "nn1", np.array_equal(g["nn1_idx"], o["nn1_idx"]), "stats", ss.tolist())
