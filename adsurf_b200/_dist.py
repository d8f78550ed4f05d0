"""Sharding of the Monte-Carlo ensemble across ranks (one process per GPU, torch.distributed).

Every (model, period, velocity) sample is independent and gradients reduce only within a model, so
the model axis S is split into contiguous blocks, one per rank.  No collective is on
the data path of a determinant sweep.  The inversion loop exchanges only
  * the Vs rows of the whole ensemble, once per iteration and only when damp_horizontal > 0 (that
    Tikhonov term couples neighbouring models, reference _model.py:471-473), and
  * the per-model losses and iterates, once at the end (best-iterate pick, _model.py:651).
Backend: NCCL over NVLink on GPUs, gloo on the CPU (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_sharded():
    """True when the ensemble should be split over the ranks (ADSURF_NO_SHARD=1 keeps a rank's run whole, e.g. to
    compare a sharded run with the unsharded one inside the same job)"""
    import os
    if os.environ.get("ADSURF_NO_SHARD") == "1":
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world():
    if is_sharded():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(S, rank=None, world=None):
    """contiguous [lo, hi) block of rank `rank`; the first S % world ranks get one extra model"""
    if rank is None or world is None:
        rank, world = rank_world()
    base, extra = divmod(S, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def min_shard_rows(S, world=None):
    """rows of the smallest shard"""
    if world is None:
        world = rank_world()[1]
    return S // world


def all_gather_rows(local, S, dim=0):
    """Concatenate the per-rank blocks along `dim` into the full [.., S, ..] tensor on every rank.

    Blocks may differ by one row; they are padded to the largest block for the collective."""
    if not is_sharded():
        return local
    rank, world = rank_world()
    sizes = [shard_bounds(S, r, world)[1] - shard_bounds(S, r, world)[0] for r in range(world)]
    big = max(sizes)
    x = local.movedim(dim, 0).contiguous()
    if min(sizes) == big:  # equal shards: one collective straight into the result
        full = torch.empty((world * big,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(full, x)
        return full.movedim(0, dim)
    if x.shape[0] < big:
        pad = torch.zeros((big - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat((x, pad), dim=0)
    out = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(out, x)
    full = torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
    return full.movedim(0, dim)
