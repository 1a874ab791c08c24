"""Collective plumbing of the contrastive step (replaces x_clip/distributed.py:14-56).

The reference gathers a python LIST of padded fp32 tensors after a size exchange with two host
syncs (distributed.py:18-23, x_clip.py:764) and a second gather for the extra latents
(x_clip.py:767-768).  Here every rank contributes equal-sized shards (asserted), ALL latent sets
travel in ONE contiguous bf16 all-gather, and the only other exchange is a [4, b] fp32 gather of
(lse, positive) pairs so each rank can form the global loss and its backward weights.
Pure torch.distributed (NCCL on GPUs, gloo in the CPU tests); no arithmetic lives here.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when no multi-rank process group exists."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_rows(shards: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Each shard is this rank's [b, D] block of one latent set.  Returns, per set, the
    [W*b, D] matrix whose rows r*b..(r+1)*b-1 are rank r's - one collective for all sets."""
    rank, w = world()
    if w == 1:
        return [s.contiguous() for s in shards]
    b, D = shards[0].shape
    for s in shards:
        assert tuple(s.shape) == (b, D), "all ranks must hold equal local batches of every latent set"
    stacked = torch.stack(list(shards)).contiguous()                       # [k, b, D]
    out = torch.empty(w * stacked.numel(), device=stacked.device, dtype=stacked.dtype)
    dist.all_gather_into_tensor(out, stacked.view(-1))                     # flat: backend agnostic
    out = out.view(w, len(shards), b, D)
    return [out[:, j].reshape(w * b, D).contiguous() for j in range(len(shards))]


def gather_stats(stats: torch.Tensor) -> torch.Tensor:
    """stats [k, b] per rank -> [k, W*b] in rank order."""
    rank, w = world()
    if w == 1:
        return stats
    k, b = stats.shape
    out = torch.empty(w * k * b, device=stats.device, dtype=stats.dtype)
    dist.all_gather_into_tensor(out, stats.contiguous().view(-1))
    return out.view(w, k, b).permute(1, 0, 2).reshape(k, w * b).contiguous()


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if world()[1] > 1:
        dist.all_reduce(t)
    return t


def reduce_scatter_rows(t: torch.Tensor) -> torch.Tensor:
    """t [W*r, D] holds this rank's partial sums for every rank's rows; returns the summed
    [r, D] block of this rank (FILIP: gradients of the gathered image-token latents)."""
    rank, w = world()
    if w == 1:
        return t
    rows = t.shape[0] // w
    t = t.contiguous()
    if dist.get_backend() == "gloo":          # gloo has no reduce_scatter: all-reduce + slice
        dist.all_reduce(t)
        return t[rank * rows:(rank + 1) * rows].clone()
    out = torch.empty((rows,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    dist.reduce_scatter_tensor(out, t)
    return out
