"""One-process-per-GPU sharding of the two hot paths.

* Loss: the batch shards across ranks; every rank runs the reference semantics on
  its shard (negatives and the pointwise means are per local batch, exactly what
  DDP would give the reference).  The only exchange is the all-reduce of the
  trainable-head gradients after backward — ``allreduce_mean_``.
* KNN: query rows shard; the pooled-feature database is all-gathered once and every
  rank ranks its own rows against the full database.  Results are disjoint row
  ranges, so the host side is a concatenation.

Pure ``torch.distributed`` plumbing (NCCL on GPUs, gloo in the CPU tests); the
split arithmetic is backend-independent and is what the gloo tests exercise.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int):
    """Contiguous near-equal row ranges: the first n % world ranks get one extra row."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[torch.Tensor], world_size: int, rank: int):
    """Slice the leading (batch) dimension of every tensor for this rank."""
    lo, hi = shard_bounds(tensors[0].shape[0], world_size, rank)
    return [t[lo:hi] for t in tensors]


def allreduce_mean_(grads: Sequence[torch.Tensor], group=None) -> None:
    """In-place mean all-reduce of gradient tensors, flattened into one bucket so a
    step costs a single collective (the head gradients are ~3 MB: latency-bound)."""
    grads = [g for g in grads if g is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def allgather_rows(local_rows: torch.Tensor, total_rows: int, group=None) -> torch.Tensor:
    """All-gather row shards laid out by ``shard_bounds`` into the full [N,F] matrix."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_rows
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total_rows, world, r) for r in range(world)]
    out = torch.empty((total_rows,) + tuple(local_rows.shape[1:]), device=local_rows.device, dtype=local_rows.dtype)
    if len({hi - lo for lo, hi in sizes}) == 1:
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    # ragged shards: collectives want equal sizes, so pad every shard to the largest and compact
    rows = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros((rows,) + tuple(local_rows.shape[1:]), device=local_rows.device, dtype=local_rows.dtype)
    padded[:local_rows.shape[0]] = local_rows
    gathered = torch.empty((world * rows,) + tuple(local_rows.shape[1:]), device=local_rows.device,
                           dtype=local_rows.dtype)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    for r, (lo, hi) in enumerate(sizes):
        out[lo:hi] = gathered[r * rows:r * rows + (hi - lo)]
    return out


def sharded_knn(local_feats: torch.Tensor, total_rows: int, k: int,
                topk_fn: Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor], group=None,
                gather_result: bool = False):
    """Query-row-sharded KNN build.  ``local_feats`` are this rank's rows of the
    normalised feature matrix; returns this rank's int64 [rows,k] block (or, with
    ``gather_result``, the concatenated [N,k] on every rank)."""
    db = allgather_rows(local_feats, total_rows, group)
    idx = topk_fn(local_feats, db, k)
    if gather_result and dist.is_initialized() and dist.get_world_size(group) > 1:
        return allgather_rows(idx, total_rows, group)
    return idx
