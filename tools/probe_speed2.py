import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
t = np.load("/tmp/pair_t.npy"); s = np.load("/tmp/pair_s.npy")
ctx = plo.Context(0)
ctx.set_target(t); ctx.set_source(s)
T, rs = ctx.register()
regs = []
for _ in range(5):
    ctx.set_target(t); ctx.set_source(s); T, rs = ctx.register(); regs.append(ctx.last_timings()["ms_register"])
print(sys.argv[1:], "iters", rs["iters"], "register ms", np.round(regs, 3), "k_project@converged ms", round(ctx.time_project_kernel(T, 10), 4))
