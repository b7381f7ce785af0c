"""Measurements for the SURVEY §8f "next" rows at the north-star size (B200): solver variants inside the resident
loop, local-map push, front-end.  Device-resident inputs; wall clock around synchronous calls, median of 10."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import plo_b200 as plo

def med(f, n=10, warm=2):
    for _ in range(warm): f()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return 1e3 * float(np.median(ts))

pair = plo.synth.workloads.hdl64_vs_map(map_points=1_000_000)
d_t = torch.from_numpy(pair.target).cuda(); d_s = torch.from_numpy(pair.source).cuda()
ctx = plo.Context(0)
ctx.set_target(d_t)
print(f"workload: {pair.source.shape[0]} source points vs {pair.target.shape[0]}-point map")
for name, kw in (("weighted LS", {}), ("trimmed LS", dict(solver=1)), ("RANSAC -> Weighted LS", dict(solver=2, ransac_final=1)),
                 ("RANSAC -> DRPM", dict(solver=2, ransac_final=2))):
    ctx.set_params(plo.default_params(**kw))
    out = {}
    def run():
        ctx.set_source(d_s); out["r"] = ctx.register()
    ms = med(run)
    T, st = out["r"]
    print(f"register [{name:22s}] {ms:7.3f} ms  iters {st['iters']}  status {st['status_name']}  pairs {st['pairs']}  "
          f"device loop {ctx.last_timings()['ms_register']:.3f} ms -> {ctx.last_timings()['ms_register'] / max(st['iters'], 1):.3f} ms/iter")
ctx.set_params(plo.default_params())
# local map: 8-frame queue of 132 k-point frames (1.06 M points), push = transform of 7 kept frames + append + index rebuild
frame = torch.from_numpy(np.ascontiguousarray(pair.source)).cuda()
T = plo.synth.scenes.pose_matrix([0.8, 0.05, 0.01], yaw_deg=1.5, pitch_deg=0.2)
ctx2 = plo.Context(0)
for _ in range(8): ctx2.map_push(frame, T, max_queue=8, transform_normals=True)
ms = med(lambda: (ctx2.map_push(frame, T, max_queue=8, transform_normals=True), ctx2.synchronize() if hasattr(ctx2, "synchronize") else ctx2.map_info()))
fr, pts = ctx2.map_info()
print(f"map_push: queue {fr} frames / {pts} points: {ms:.3f} ms per push (index build alone {ctx2.last_timings()['ms_index_build']:.3f} ms); "
      f"set_target of the same {pts} points from the host costs the upload of {pts * 48 / 1e6:.0f} MB instead of {frame.shape[0] * 48 / 1e6:.1f} MB")
# front-end
raw = torch.from_numpy(np.ascontiguousarray(pair.source[:, 0:3])).cuda()
out = {}
def fe(): out["s"] = ctx.frontend(raw, fetch=False)
ms = med(fe)
print(f"frontend: {raw.shape[0]} raw points -> {out['s']['n']} filtered, {out['s']['candidates']} presampled: {ms:.3f} ms per scan (16 launches, one sync)")
