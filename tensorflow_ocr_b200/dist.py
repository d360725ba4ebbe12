"""Batch sharding across GPUs.

The head shards by image, exactly like the reference's in-graph towers
(multigpu_train.py:111-125 ``tf.split`` + one tower per GPU): every rank runs the
fused head on its contiguous ``B/G`` slice with shard-local normalisers, so the
gradients need no communication.  The only collective is one all-reduce of the
loss scalars for reporting (the reference prints the last tower's loss only,
multigpu_train.py:125,171 — quirk Q18).  Decode needs no communication.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib

__all__ = ["shard_bounds", "shard_batch", "allreduce_loss_stats", "LossStatsReducer"]


def shard_bounds(B: int, world_size: int, rank: int):
    """Contiguous split identical to tf.split(axis=0) into equal parts; the remainder
    (when B % world_size != 0) goes to the first ranks."""
    base, rem = divmod(B, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(arrays: dict, world_size: int, rank: int) -> dict:
    B = next(iter(arrays.values())).shape[0]
    lo, hi = shard_bounds(B, world_size, rank)
    return {k: v[lo:hi] for k, v in arrays.items()}


def allreduce_loss_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum-all-reduce the first PLH_STATS_FLOATS loss scalars (<= 256 B, latency bound) and
    return the cross-shard MEAN of the loss terms (tower average) together with the
    summed counts.  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    v = stats[:_lib.STATS_FLOATS].clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        w = dist.get_world_size(group)
        mean_idx = [_lib.ST_TOTAL, _lib.ST_L_PIX, _lib.ST_LINK_TOTAL] + list(range(_lib.ST_L_LINK, _lib.ST_L_LINK + 8))
        v[mean_idx] = v[mean_idx] / w
    return v


_MEAN_IDX = [_lib.ST_TOTAL, _lib.ST_L_PIX, _lib.ST_LINK_TOTAL] + list(range(_lib.ST_L_LINK, _lib.ST_L_LINK + 8))


class LossStatsReducer(object):
    """The per-step NCCL all-reduce of the loss scalars, kept OFF the critical path.

    `submit(stats)` snapshots the 64 scalars and launches the all-reduce on a private side
    stream that waits for the producing stream; the producing stream never waits for the
    collective.  `result()` blocks the host and returns the tower-mean loss terms and the
    summed counts (what multigpu_train.py would print if it reported every tower, quirk Q18).
    `event` can be waited on by a stream that is about to overwrite `stats`.
    """

    def __init__(self, device, group=None, stream=None):
        self.group = group
        # several reducers can share one side stream (every extra CUDA stream competes for the device's
        # few hardware work queues with the streams that carry the head's two branches)
        self.stream = stream if stream is not None else torch.cuda.Stream(device)
        self.buf = torch.zeros(_lib.STATS_FLOATS, dtype=torch.float32, device=device)
        self.event = torch.cuda.Event()
        self.pending = False

    def submit(self, stats: torch.Tensor):
        cur = torch.cuda.current_stream(stats.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.buf.copy_(stats[:_lib.STATS_FLOATS], non_blocking=True)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
            self.event.record(self.stream)
        self.pending = True

    def result(self) -> torch.Tensor:
        self.event.synchronize()
        v = self.buf.clone()
        w = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        v[_MEAN_IDX] = v[_MEAN_IDX] / w
        return v
