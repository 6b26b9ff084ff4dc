"""Multi-GPU sharding of the hot path.

Every window and every image is independent until the per-image purge, so the path shards with no
collective on the data path: batch mode assigns images round-robin to ranks, a single huge batch of
windows is cut into contiguous, tile-aligned ranges (concatenating rank results in rank order reproduces
the reference order).  Only the per-image detection lists -- dozens of rows -- are gathered, as Python
objects on the host (``torch.distributed.all_gather_object``; gloo or nccl process groups both work).
One process per GPU (``torchrun``); nothing here launches kernels.
"""
from __future__ import annotations

TILE = 128


def image_shard(n_images, rank, world):
    """Indices of the images rank `rank` processes (round robin)."""
    return list(range(rank, n_images, world))


def window_shard(n_windows, rank, world, align=TILE):
    """Contiguous [start, stop) of the windows of rank `rank`; boundaries are multiples of `align`."""
    tiles = (n_windows + align - 1) // align
    per = (tiles + world - 1) // world
    start = min(n_windows, rank * per * align)
    stop = min(n_windows, (rank + 1) * per * align)
    return start, stop


def gather_detections(local, image_ids, n_images, group=None):
    """local[i] belongs to image image_ids[i]; returns the list of all n_images results on every rank
    (single process: just reorders)."""
    import torch.distributed as dist
    pairs = list(zip(image_ids, local))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        buckets = [None] * dist.get_world_size(group)
        dist.all_gather_object(buckets, pairs, group=group)
        pairs = [p for b in buckets for p in b]
    out = [None] * n_images
    for k, d in pairs:
        out[k] = d
    if any(o is None for o in out):
        raise RuntimeError("detections of %d images are missing after the gather" % sum(o is None for o in out))
    return out
