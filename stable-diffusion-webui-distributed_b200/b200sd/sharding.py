"""Batch-index sharding of one request over ranks (one process per GPU) and the single end-of-sampling gather.

Image k of a request depends only on (prompt, seed + k, weights) — the reference exploits exactly this when it offsets
each job's seed by the images owned by earlier jobs (scripts/distributed.py:297-305).  So rank r takes the global
image indices [start_r, start_r + count_r) with seeds seed + index, runs all sampler steps with no exchange, and the
decoded uint8 images are all-gathered once (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_counts(total: int, world: int) -> List[int]:
    """even split, remainder one-by-one to the first ranks — what World.optimize_jobs produces for equal-speed workers
    (reference world.py:482-504; golden: 8 images on 3 workers -> 3, 3, 2)."""
    base, rem = divmod(total, world)
    return [base + (1 if r < rem else 0) for r in range(world)]


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    counts = shard_counts(total, world)
    return sum(counts[:rank]), counts[rank]


def all_gather_images(local: torch.Tensor, counts: List[int], group=None) -> torch.Tensor:
    """local [count_r, H, W, 3] uint8 -> [sum(counts), H, W, 3] in global image order on every rank.
    Uneven shards are padded to max(counts) for the collective and trimmed afterwards."""
    world = len(counts)
    if world == 1:
        return local
    cmax = max(counts)
    pad = local
    if local.shape[0] < cmax:
        pad = torch.cat([local, local.new_zeros((cmax - local.shape[0], *local.shape[1:]))])
    out = local.new_empty((world * cmax, *local.shape[1:]))
    if local.is_cuda:
        dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    else:
        dist.all_gather(list(out.chunk(world)), pad.contiguous(), group=group)
    if all(c == cmax for c in counts):
        return out
    return torch.cat([out[r * cmax: r * cmax + c] for r, c in enumerate(counts)])
