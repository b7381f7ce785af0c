import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import plo_b200 as plo
pair = plo.synth.workloads.hdl64_vs_map(map_points=1_000_000)
np.save("/tmp/pair_t.npy", pair.target); np.save("/tmp/pair_s.npy", pair.source)
