"""Tile sharding and the one collective of this path: an all-reduce of the C x C confusion matrix.

Tiles (and mosaics) are independent units, so every rank processes its own contiguous index range with no data-path
exchange; integer confusion counts are summed once per report point (exact, independent of the number of ranks).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n units for `rank` of `world`; the last rank takes the remainder (SURVEY.md 8(d) cfg 3)."""
    per = n // world
    lo = rank * per
    hi = n if rank == world - 1 else lo + per
    return lo, hi


def shard_by_key(keys, rank, world):
    """Indices whose key hashes to this rank -- used to keep all tiles of one image on one GPU for the big-mask path
    (segmentation_test.py:160: image_idx = name.split('_')[0])."""
    uniq = sorted(set(keys))
    owner = {k: i % world for i, k in enumerate(uniq)}
    return [i for i, k in enumerate(keys) if owner[k] == rank]


def all_reduce_confusion(conf, group=None):
    """In-place SUM of an int64 [C,C] tensor over the process group (NCCL on GPUs, gloo on CPU); no-op when not distributed."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(conf, op=dist.ReduceOp.SUM, group=group)
    return conf
