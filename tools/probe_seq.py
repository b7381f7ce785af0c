"""BASELINE config 2 in miniature: a VLP-32C-shaped synthetic sequence through LaserOdometry (resident loop), wall
clock per frame against the device time of its index build and registration."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seq = plo.synth.workloads.Sequence(seed=2001, n_frames=n)
frames = [seq.frame(k) for k in range(n)]
for rep in range(2):
    odo = plo.LaserOdometry(resident=True)
    idx, reg, wall = [], [], []
    for f in frames:
        t0 = time.perf_counter()
        odo.process_frame(f)
        wall.append(time.perf_counter() - t0)
        t = odo.ctx.last_timings(); idx.append(t["ms_index_build"]); reg.append(t["ms_register"])
iters = [s["iters"] for s in odo.frame_stats[1:]]
print(json.dumps(dict(frames=n, points=int(np.mean([f.shape[0] for f in frames])), wall_ms_per_frame=1e3 * float(np.mean(wall[2:])),
                      fps=1.0 / float(np.mean(wall[2:])), ms_index_build=float(np.mean(idx[2:])), ms_register=float(np.mean(reg[2:])),
                      iters_mean=float(np.mean(iters)))))
