"""Phase breakdown of k_register_loop from globaltimer stamps (needs a -DPLO_LOOP_TIMING build: PLO_LIB=...)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_vs_map()
ctx = plo.Context(0)
for i in range(4):
    ctx.set_target(pair.target); ctx.set_source(pair.source); T, rs = ctx.register()
buf = np.zeros(128, np.uint64)
ctx.L.plo_debug_loop_stamps.argtypes = [C.c_void_p, C.c_void_p]
ctx.L.plo_debug_loop_stamps(ctx.h, buf.ctypes.data_as(C.c_void_p))
st = buf.reshape(16, 8)[: rs["iters"], :7].astype(np.int64)
names = ["project", "barrier1", "reduce", "barrier2", "sum partials", "solve"]
print("iterations", rs["iters"], "register ms", ctx.last_timings()["ms_register"])
for it in range(rs["iters"]):
    d = np.diff(st[it]) / 1e3
    gap = (st[it + 1, 0] - st[it, 6]) / 1e3 if it + 1 < rs["iters"] else 0.0
    print(f"it {it}: " + "  ".join(f"{n} {v:7.1f}" for n, v in zip(names, d)) + f"  | to next {gap:5.1f} us")

try:
    sb = np.zeros(8, np.uint64)
    ctx.L.plo_debug_solve_stamps.argtypes = [C.c_void_p]
    ctx.L.plo_debug_solve_stamps(sb.ctypes.data_as(C.c_void_p))
    d = np.diff(sb[:6].astype(np.int64)) / 1e3
    print("last solve (us): state copy %.1f  ldlt %.1f  (bookkeeping %.1f)  rodrigues+polar %.1f  compose+tests %.1f" % tuple(d))
except AttributeError:
    pass
