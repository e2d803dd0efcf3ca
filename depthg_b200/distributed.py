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


def knn_shard_bounds(n: int, world_size: int, rank: int):
    """Row ranges of the KNN build: every rank owns ceil(n / world) rows and the LAST ranks run short, so the equal-size
    all-gather of the shards (padded at the very end only) IS the database, in place, without a compaction pass."""
    per = -(-n // world_size)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


_comm_streams = {}


def sharded_knn_build(local_feats: torch.Tensor, total_rows: int, k: int, group=None, return_stats: bool = False):
    """The query-row-sharded KNN build with the all-gather overlapped: ``local_feats`` are rows
    ``knn_shard_bounds(total_rows, world, rank)`` of the normalised feature matrix.  The all-gather of the database
    runs on a side stream while this rank already ranks its rows against the rows it holds
    (``KnnShard.begin``); ``finish`` continues over the remote rows.  Returns this rank's int64 [rows,k] block."""
    from .precompute_knns import KnnShard, knn_topk
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = knn_topk(local_feats, local_feats, k, return_stats=return_stats)
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = knn_shard_bounds(total_rows, world, rank)
    if local_feats.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} owns rows [{lo},{hi}) of {total_rows}, got {local_feats.shape[0]} rows")
    per = -(-total_rows // world)
    dev = local_feats.device
    local_feats = local_feats.contiguous()
    src = local_feats
    if hi - lo != per:                       # a short last shard: pad the all-gather input (the pad lands beyond row N)
        src = torch.zeros((per, local_feats.shape[1]), device=dev, dtype=local_feats.dtype)
        src[:hi - lo] = local_feats
    gathered = torch.empty((world * per, local_feats.shape[1]), device=dev, dtype=local_feats.dtype)
    cur = torch.cuda.current_stream(dev)
    comm = _comm_streams.get(dev.index)
    if comm is None:
        comm = _comm_streams[dev.index] = torch.cuda.Stream(device=dev)
    comm.wait_stream(cur)
    with torch.cuda.stream(comm):
        dist.all_gather_into_tensor(gathered, src, group=group)
    # Two phases pay when the local pass is short enough to hide under the all-gather: it runs on the SMs the
    # communication kernels leave free (112 of 148, one CTA per SM and 128 query rows per CTA), 256 database rows per
    # tile at ~14 us a tile.  With few ranks the local rows are a large part of the database and one pass is faster.
    nblocks = -(-(hi - lo) // 128)
    nseg_local = max(1, min(112 // max(nblocks, 1), 8))
    overlap = nblocks <= 112 and -(-(hi - lo) // 256) / nseg_local <= 26
    if overlap and hi > lo:
        shard = KnnShard(local_feats, lo, total_rows, k).begin()
        cur.wait_stream(comm)
        return shard.finish(gathered[:total_rows], return_stats=return_stats)
    cur.wait_stream(comm)
    if hi == lo:
        idx = torch.empty((0, k), device=dev, dtype=torch.int64)
        return (idx, {"pipeline_error": 0, "fallback_rows": 0}) if return_stats else idx
    return knn_topk(local_feats, gathered[:total_rows], k, return_stats=return_stats)
