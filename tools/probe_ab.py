"""A/B timing of one build (PLO_LIB selects it) on the north-star workload: registration time (CUDA graph,
events), per-launch k_project time over the 7 iterations (profiling mode) and a pose / pair-count
fingerprint that must not change between builds.  usage: PLO_LIB=... python tools/probe_ab.py [tag]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
cache = "/tmp/plo_pair.npz"
if os.path.exists(cache):
    z = np.load(cache); target, source = z["t"], z["s"]
else:
    pair = plo.synth.workloads.hdl64_vs_map(); target, source = pair.target, pair.source
    np.savez(cache, t=target, s=source)
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ctx = plo.Context(0)
for kv in os.environ.get("PLO_TUNE", "").split(","):   # e.g. PLO_TUNE=force_warm=1,group=16
    if "=" in kv: ctx.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
regs, idxs = [], []
for i in range(8):
    ctx.set_target(target); ctx.set_source(source)
    flush.zero_(); torch.cuda.synchronize()
    T, rs = ctx.register()
    if i >= 2: regs.append(ctx.last_timings()["ms_register"]); idxs.append(ctx.last_timings()["ms_index_build"])
ctx.set_profiling(True)
kp = []
for i in range(3):
    ctx.set_target(target); ctx.set_source(source); T, rs = ctx.register(); kp.append(ctx.last_kernel_timings()["ms_project_mean"]); each = ctx.last_project_times()
ctx.set_profiling(False)
misses = ctx.last_tile_misses()
steady = ctx.time_project_kernel(T, 10)
fp = float(np.abs(T).sum())
print(f"[{tag}] iters {rs['iters']} pairs {rs['pairs']} fp {fp:.12f} | index build ms {np.median(idxs):.4f} | register ms median {np.median(regs):.4f} min {np.min(regs):.4f} | "
      f"k_project mean/launch {np.mean(kp):.4f} | steady (converged pose) {steady:.4f} | per launch {np.round(each, 3).tolist()} | tile misses {misses.tolist()}", flush=True)
