"""Config-2 / 5 shaped pair (VLP-32C frame against the previous frame, ~52 k points each): where a registration's time
goes -- index build, loop, per-projection times -- and the per-pair time of a batch call.  usage: python tools/probe_seq.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
fs = plo.synth.workloads.generate_sequences([5000], 9)[0]
frames = [fs.frame(k) for k in range(9)]
import torch
pinned = [torch.from_numpy(f).pin_memory() for f in frames]
ctx = plo.Context(0)
idx, reg = [], []
for rep in range(3):
    for k in range(1, 9):
        ctx.set_target(pinned[k - 1]); ctx.set_source(pinned[k]); T, rs = ctx.register()
        tm = ctx.last_timings()
        if rep: idx.append(tm["ms_index_build"]); reg.append(tm["ms_register"])
print(f"points {frames[0].shape[0]}: index build {np.median(idx):.4f} ms, loop {np.median(reg):.4f} ms ({rs['iters']} iterations last)")
ctx.set_profiling(True)
ctx.set_target(pinned[3]); ctx.set_source(pinned[4]); T, rs = ctx.register()
print("per projection", np.round(ctx.last_project_times(), 4).tolist(), "misses", ctx.last_tile_misses().tolist())
ctx.set_profiling(False)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    Tb, sb = ctx.register_batch(pinned[1:], pinned[:-1])
    dt = time.perf_counter() - t
print(f"batch of 8 pairs: {1e3 * dt / 8:.4f} ms per pair, iterations {[s['iters'] for s in sb]}")
