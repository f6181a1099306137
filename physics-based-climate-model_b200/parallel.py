"""Host-side data-parallel plumbing (device agnostic: NCCL on the GPUs, gloo in the CPU tests).

The path shards by sample: every rank owns a full model replica and a disjoint slice of the
training windows; the ONLY exchange in a step is the all-reduce of the flat gradient buffer, and the only
other cross-rank quantity is the (associative) metric partial sums."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_windows(n_windows: int, rank: int, world: int, seed: int = 42, epoch: int = 0, shuffle: bool = True,
                  drop_last: bool = False) -> torch.Tensor:
    """Indices of the training windows this rank visits in `epoch` — the partition Lightning's auto-injected
    DistributedSampler would produce for main_final.py's DataLoader (shuffle with seed+epoch, pad by wrapping
    so every rank gets the same count, then stride by world size)."""
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        idx = torch.randperm(n_windows, generator=g)
    else:
        idx = torch.arange(n_windows)
    if drop_last:
        per = n_windows // world
        idx = idx[: per * world]
    else:
        per = (n_windows + world - 1) // world
        pad = per * world - n_windows
        if pad:
            idx = torch.cat([idx, idx[:pad]])
    return idx[rank::world]


def allreduce_flat_grads(flat_grad: torch.Tensor, n_reduced: int, group=None) -> float:
    """Sum the first `n_reduced` entries of the flat gradient buffer over all ranks (in place) and return
    the scale (1/world) the optimizer must apply — the mean-gradient semantics of DDP.  Entries past
    n_reduced (parameters the forward never uses, e.g. AttUNetConvLSTM.post_conv) are left alone."""
    _, world = world_info(group)
    if world > 1 and n_reduced > 0:
        dist.all_reduce(flat_grad[:n_reduced], op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def allreduce_metric_partials(partial: torch.Tensor, t_local: int, group=None):
    """Metric partial sums are associative over time shards: sum them (and the row counts) over ranks."""
    _, world = world_info(group)
    t = torch.tensor([float(t_local)], dtype=torch.float64, device=partial.device)
    if world > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return partial, int(t.item())
