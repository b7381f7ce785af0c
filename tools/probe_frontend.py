"""Front-end (plo_frontend) timing on one HDL-64 scan: device-resident input, wall clock around 20 runs."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_pair()
pts = np.ascontiguousarray(pair.source[:, 0:3])
ctx = plo.Context(0)
dev = torch.from_numpy(pts).cuda()
for _ in range(3):
    st = ctx.frontend(dev, fetch=False)
torch.cuda.synchronize()
l0 = ctx.launch_count
t = time.perf_counter()
for _ in range(20):
    st = ctx.frontend(dev, fetch=False)       # stats read-back = one sync per run
dt = (time.perf_counter() - t) / 20
print("points", pts.shape[0], "->", st["n"], "gpu ms/scan", dt * 1e3, "launches/scan", (ctx.launch_count - l0) / 20)
