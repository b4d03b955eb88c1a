"""Range-sharded MSM across the GPUs of one box (one process per GPU, torch.distributed).

ark-ec's multi_scalar_mul is a plain sum over (base, scalar) pairs, so it shards by index range
with no data-path collective: rank r owns pairs [lo_r, hi_r), runs the full bucket pipeline on
them and produces ONE affine point.  The only exchange is the handful of partial points
(104 B each for BLS12-381 G1), all-gathered over NCCL/NVLink (gloo on CPU in the tests) and
summed on every rank -- the "final GroupProjective additions" of a sharded caller.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of [0, n): the first n % world ranks get one extra pair."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def record_words(coord_words: int) -> int:
    """Result record: x, y (coord_words each) + one flag word."""
    return 2 * coord_words + 1


def gather_records(local_record, world: int, group=None):
    """All-gather one fixed-size record per rank.  `local_record` is a 1-D torch tensor (int64) on the
    backend's device; returns a (world, len) tensor in rank order."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_record.reshape(1, -1).clone()
    out = torch.empty((world, local_record.numel()), dtype=local_record.dtype, device=local_record.device)
    dist.all_gather_into_tensor(out.view(-1), local_record.contiguous(), group=group)
    return out


def sharded_msm(n: int, rank: int, world: int, local_msm: Callable[[int, int], "object"],
                sum_records: Callable[["object"], "object"], group=None):
    """local_msm(lo, hi) -> this rank's partial record (torch int64 tensor);
    sum_records((world, len) tensor) -> the final record.  Every rank returns the same result."""
    lo, hi = shard_range(n, rank, world)
    part = local_msm(lo, hi)
    allp = gather_records(part, world, group)
    return sum_records(allp)
