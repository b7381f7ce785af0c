"""Batched multi-GPU mode: registration units (frame pairs / sequences) are
independent — every frame starts from the identity (src/laser_odometry.cpp:484-485) against the
previous frame in its own coordinates (:116-136) — so they are sharded across ranks with no
data-path collective.  The only exchange is one all-gather of per-unit poses + stats at the end
(`torch.distributed`: NCCL over NVLink on GPUs, gloo in the CPU tests).  One process per GPU.
"""
from __future__ import annotations

import numpy as np

RESULT_WIDTH = 20   # 16 pose + iters + pairs + rms + status


def shard_units(n_units: int, rank: int, world: int) -> list[int]:
    """Unit u runs on rank u mod world (sequence s -> GPU s mod G)."""
    return list(range(rank, n_units, world))


def pack_result(T: np.ndarray, stats: dict | None) -> np.ndarray:
    row = np.zeros(RESULT_WIDTH, np.float64)
    row[:16] = np.asarray(T, np.float64).reshape(16)
    if stats is not None:
        row[16:] = [stats["iters"], stats["pairs"], stats["rms"], stats["status"]]
    return row


def gather_results(local_rows: np.ndarray, local_units: list[int], n_units: int, device=None, slots: int | None = None):
    """All-gather [units_local x RESULT_WIDTH] fp64 blocks and place them by unit id.
    Returns the [n_units x RESULT_WIDTH] table on every rank.  `slots`: rows reserved per rank when the caller
    knows the largest shard (skips the sizing all-reduce: ONE collective in total)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    table = np.zeros((n_units, RESULT_WIDTH), np.float64)
    if world == 1:
        table[local_units] = local_rows
        return table
    # shards may be ragged (sequences of different lengths): size the slots by the largest one
    if slots is None:
        cnt = torch.tensor([len(local_units)], dtype=torch.int64, device=device)
        dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
        per = max(int(cnt.item()), 1)
    else:
        per = max(int(slots), len(local_units), 1)
    buf = torch.zeros((per, RESULT_WIDTH + 1), dtype=torch.float64, device=device)
    if len(local_units):
        buf[:len(local_units), :RESULT_WIDTH] = torch.as_tensor(np.asarray(local_rows), dtype=torch.float64, device=device)
        buf[:len(local_units), RESULT_WIDTH] = torch.as_tensor(np.asarray(local_units, np.float64) + 1.0, device=device)
    out = torch.empty((world * per, RESULT_WIDTH + 1), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy()
    ids = out[:, RESULT_WIDTH].astype(np.int64) - 1
    valid = ids >= 0
    table[ids[valid]] = out[valid, :RESULT_WIDTH]
    return table


def chain_poses(rel: np.ndarray) -> np.ndarray:
    """Global poses from per-frame relative ones: nowPose = prevLaserPose * rPose
    (src/laser_odometry.cpp:649-655); rel[0] is the identity of the seeding frame."""
    out = np.empty_like(rel)
    cur = np.eye(4)
    for i in range(rel.shape[0]):
        cur = cur @ rel[i]
        out[i] = cur
    return out


def register_sequences_sharded(ctx, sequences, device=None, timing: dict | None = None):
    """cfg-5: `sequences` is a list of objects with n_frames / frame(k); each rank registers the
    frame pairs of its own sequences (one plo_register_batch per sequence) and all ranks end with
    every sequence's global trajectory.  Returns (list of [n_frames, 4, 4] arrays, result table).
    `timing` (optional dict) receives `register_s` (this rank's registrations, host clock around the
    batch calls, which synchronise) and `gather_ms` (the one collective)."""
    import time

    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    offsets = np.cumsum([0] + [s.n_frames for s in sequences])
    n_units = int(offsets[-1])
    rows, units = [], []
    t0 = time.perf_counter()
    for si in shard_units(len(sequences), rank, world):
        seq = sequences[si]
        frames = [seq.frame(k) for k in range(seq.n_frames)]
        T, st = ctx.register_batch(frames[1:], frames[:-1]) if seq.n_frames > 1 else (np.zeros((0, 4, 4)), [])
        rows.append(pack_result(np.eye(4), None))
        units.append(int(offsets[si]))
        for k in range(1, seq.n_frames):
            rows.append(pack_result(T[k - 1], st[k - 1]))
            units.append(int(offsets[si]) + k)
    t1 = time.perf_counter()
    # every rank knows the largest shard (sequence lengths are global knowledge): ONE collective
    slots = max(sum(sequences[si].n_frames for si in shard_units(len(sequences), r, world)) for r in range(world))
    table = gather_results(np.asarray(rows).reshape(-1, RESULT_WIDTH), units, n_units, device=device, slots=slots)
    if timing is not None:
        timing["register_s"] = t1 - t0
        timing["gather_ms"] = 1e3 * (time.perf_counter() - t1)
    return [chain_poses(table[offsets[i]:offsets[i + 1], :16].reshape(-1, 4, 4)) for i in range(len(sequences))], table
