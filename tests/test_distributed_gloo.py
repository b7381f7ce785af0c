"""Multi-process host logic of the batched mode on CPU (gloo, world_size 2): units are sharded
unit -> rank (u mod G), every rank ends with the full pose table, the table is identical to a
single-process run.  The registration itself is replaced by a deterministic fake (no GPU here);
the real thing is covered by the -m gpu tests and bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import plo_b200 as plo
from plo_b200 import distributed as D


class FakeSeq:
    def __init__(self, seed, n_frames):
        self.seed, self.n_frames = seed, n_frames

    def frame(self, k):
        return np.full((4, 12), self.seed * 100 + k, np.float32)


class FakeCtx:
    """register_batch returns a pose that encodes (source id, target id)"""

    def register_batch(self, sources, targets):
        T = np.tile(np.eye(4), (len(sources), 1, 1))
        stats = []
        for i, (s, t) in enumerate(zip(sources, targets)):
            T[i, 0, 3] = float(s[0, 0])
            T[i, 1, 3] = float(t[0, 0])
            ang = 0.01 * (i + 1)
            T[i, :2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
            stats.append(dict(iters=3 + i, pairs=1000 + i, rms=0.01 * i, status=1))
        return T, stats


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seqs = [FakeSeq(s, 3 + s % 3) for s in range(5)]
    trajs, table = D.register_sequences_sharded(FakeCtx(), seqs)
    q.put((rank, [t.copy() for t in trajs], table.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_units_covers_everything_once():
    for n, w in ((64, 8), (5, 2), (3, 4), (0, 2)):
        seen = sorted(u for r in range(w) for u in D.shard_units(n, r, w))
        assert seen == list(range(n))
        assert D.shard_units(n, 0, 1) == list(range(n))


def test_chain_poses_matches_reference_composition():
    rng = np.random.default_rng(0)
    rel = np.stack([plo.synth.scenes.pose_matrix(rng.normal(size=3), yaw_deg=rng.normal()) for _ in range(6)])
    rel[0] = np.eye(4)
    out = D.chain_poses(rel)
    cur = np.eye(4)
    for i in range(6):
        cur = cur @ rel[i]          # nowPose = prevLaserPose * rPose, src/laser_odometry.cpp:652
        assert np.array_equal(out[i], cur)


def test_sharded_gather_world2_equals_single_process():
    seqs = [FakeSeq(s, 3 + s % 3) for s in range(5)]
    ref_trajs, ref_table = D.register_sequences_sharded(FakeCtx(), seqs)      # no process group: world 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, trajs, table in res:
        assert np.array_equal(table, ref_table)                               # bitwise identical on every rank
        for a, b in zip(trajs, ref_trajs):
            assert np.array_equal(a, b)
    # rows carry iters / pairs / rms / status of each frame pair
    assert ref_table[1, 16] == 3 and ref_table[1, 19] == 1
