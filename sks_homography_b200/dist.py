"""Multi-GPU driver: one process per GPU, torch.distributed for the plumbing.

Streaming solvers shard the batch into contiguous index ranges and need no
collective at all; fused ACA-RANSAC shards the hypothesis ids
of every image pair and merges the per-rank best keys with ONE integer
max-all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  The winning
model is recomputed from its id on every rank, so no homography travels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._lib import lib


def world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int | None = None, world_size: int | None = None, group=None
                ) -> tuple[int, int]:
    """(begin, count) of this rank's contiguous shard of n units."""
    r, w = world(group)
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return lib().shard_range(n, rank, world_size)


def merge_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """In-place max-combine of packed (count<<32 | ~hyp) keys across ranks.

    The keys are < 2^63 (counts are < 2^31), so the signed int64 max NCCL / gloo
    implement orders them exactly like the unsigned keys.
    """
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    return keys


def ransac_aca(corr: torch.Tensor, n_hyp: int, seed: int, thr2: float,
               samples: torch.Tensor | None = None, group=None, want_mask: bool = False):
    """Hypothesis-sharded fused ACA-RANSAC.  Every rank holds all pairs' matches
    `corr` [P, n_pts, 4]; returns (H_best [P,9], inlier_count [P], hyp_id [P], mask)."""
    from . import api
    begin, count = shard_range(n_hyp, group=group)
    keys = api.ransac_keys(corr, n_hyp, seed, thr2, samples, begin, count)
    merge_keys(keys, group)
    H, cnt, mask = api.ransac_finalize(corr, n_hyp, seed, thr2, keys, samples, want_mask)
    _, hyp = api.decode_keys(keys)
    return H, cnt, hyp, mask
