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

import os
import time
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
_symm_states = {}


class _SymmKnnState:
    """Per-(N, F, k, group) state of the peer-copy exchange: the KNN workspace and the gathered fp32 database live in
    symmetric memory (torch.distributed._symmetric_memory: every rank can address every rank's buffers over NVLink),
    so a rank PUSHES its bf16 panel rows and fp32 rows into its peers' buffers with plain device-to-device copies -
    copy engines, no SMs, nothing for the tensor pass to compete with - instead of running an all-gather kernel.
    (Pushes, not pulls: posted writes stream over NVLink, reads pay the round trip; measured 2x apart.)"""

    def __init__(self, N, F, k, group, dev):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        from .precompute_knns import knn_panel_layout
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        world, rank = self.world, self.rank
        self.per = -(-N // world)
        self.N, self.F, self.k, self.dev = N, F, k, dev
        self.ws_bytes = _lib.lib().dg_knn_workspace_bytes(self.per, N, F, k)
        self.ws = symm.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.gathered = symm.empty((N, F), dtype=torch.float32, device=dev)
        gname = (group or dist.group.WORLD).group_name
        self.h_ws = symm.rendezvous(self.ws, gname)
        self.h_g = symm.rendezvous(self.gathered, gname)
        hi_off, lo_off, row_bytes = knn_panel_layout(N, F)
        self.npanels = _lib.lib().dg_knn_panel_count()      # 1: the fast pass reads (and the peers need) the hi panel only
        self.comm = [torch.cuda.Stream(device=dev) for _ in range(max(world - 1, 1))]
        self.ev_split, self.ev_ready = torch.cuda.Event(), torch.cuda.Event()
        self.ev_panels, self.ev_rows = torch.cuda.Event(), torch.cuda.Event()
        # the two copy batches (pointers never change): A = this rank's panel rows and its squared-norm maximum into
        # every peer's workspace, B = its fp32 rows into every peer's copy of the database; one stream per peer
        lo, hi = knn_shard_bounds(N, world, rank)
        wsp, gp = self.ws.data_ptr(), self.gathered.data_ptr()
        a, b = [], []
        for step in range(1, world):
            p = (rank + step) % world
            st = self.comm[step - 1].cuda_stream
            pw = self.h_ws.get_buffer(p, (self.ws_bytes,), torch.uint8, 0).data_ptr()
            pg = self.h_g.get_buffer(p, (N, F), torch.float32, 0).data_ptr()
            if hi > lo:
                for off in (hi_off, lo_off)[:self.npanels]:
                    o = off + lo * row_bytes
                    a.append((pw + o, wsp + o, (hi - lo) * row_bytes, st))
                b.append((pg + lo * F * 4, gp + lo * F * 4, (hi - lo) * F * 4, st))
            a.append((pw + 128 + 4 * rank, wsp + 8, 4, st))     # header slot 32 + rank of the peer <- this rank's maximum

        def pack(items):
            n = len(items)
            return (n, (C.c_void_p * n)(*[i[0] for i in items]), (C.c_void_p * n)(*[i[1] for i in items]),
                    (C.c_size_t * n)(*[i[2] for i in items]), (C.c_void_p * n)(*[i[3] for i in items]))

        self.batch_a, self.batch_b = pack(a), pack(b)
        self.trace = None

    def build(self, local, return_stats):
        from . import _lib
        from ._lib import check
        from .precompute_knns import KnnShard
        N, k, world, rank = self.N, self.k, self.world, self.rank
        lo, hi = knn_shard_bounds(N, world, rank)
        nq = hi - lo
        cur = torch.cuda.current_stream(self.dev)
        trace = self.trace = [] if os.environ.get("DEPTHG_KNN_TRACE") else None

        def mark(name, stream):                   # DEPTHG_KNN_TRACE=1: (name, host seconds, event) per milestone
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                trace.append((name, time.perf_counter(), ev))

        mark("start", cur)
        shard = None
        lib = _lib.lib()
        if nq:
            self.gathered[lo:hi].copy_(local)
            shard = KnnShard(local, lo, N, k, ws=self.ws).begin(_lib.KNN_SPLIT_LOCAL)
        else:
            self.ws[:256].zero_()
        self.ev_split.record(cur)
        mark("split_local", cur)
        c0 = self.comm[0]
        c0.wait_event(self.ev_split)
        with torch.cuda.stream(c0):
            # every rank is through its previous build (stream order) and has its panel rows ready: peers' buffers
            # may be written
            self.h_ws.barrier(channel=0)
            self.ev_ready.record(c0)
            mark("barrier0", c0)
        if shard is not None:
            shard.begin(_lib.KNN_PASS_LOCAL | _lib.KNN_ALL_SMS)      # under the copies: they need no SMs
        mark("pass_local", cur)
        for c in self.comm[1:]:
            c.wait_event(self.ev_ready)
        check(lib.dg_memcpy_batch(*self.batch_a), "dg_memcpy_batch")
        for c in self.comm[1:]:
            c0.wait_stream(c)
        with torch.cuda.stream(c0):
            self.h_ws.barrier(channel=1)          # every rank's panel rows have landed everywhere
            self.ev_panels.record(c0)
            mark("panels", c0)
        if shard is not None:                     # enqueue the remote pass before the host spends time on the rest
            cur.wait_event(self.ev_panels)
            shard.finish(None, phases=_lib.KNN_PASS_REMOTE)
            mark("pass_remote", cur)
        for c in self.comm[1:]:
            c.wait_event(self.ev_panels)          # (keeps the fp32 copies behind ALL panel copies on the links)
        check(lib.dg_memcpy_batch(*self.batch_b), "dg_memcpy_batch")   # fp32 rows: only the re-rank reads them
        for c in self.comm[1:]:
            c0.wait_stream(c)
        with torch.cuda.stream(c0):
            self.h_ws.barrier(channel=2)          # ... and so have the fp32 rows
            self.ev_rows.record(c0)
            mark("rows", c0)
        if shard is None:
            idx = torch.empty((0, k), device=self.dev, dtype=torch.int64)
            return (idx, {"pipeline_error": 0, "fallback_rows": 0}) if return_stats else idx
        cur.wait_event(self.ev_rows)
        out = shard.finish(self.gathered, return_stats=False, phases=_lib.KNN_RERANK, npeer_max=world)
        mark("rerank", cur)
        if return_stats:
            head = self.ws[:12].view(torch.int32).cpu()
            return out, {"pipeline_error": int(head[0]), "fallback_rows": int(head[1])}
        return out


def _symm_knn_state(N, F, k, group, dev):
    """The cached exchange state, or None when symmetric memory cannot be set up here (all ranks agree)."""
    from . import _lib
    key = (N, F, k, id(group), dev.index, _lib.lib().dg_knn_panel_count())
    if key not in _symm_states:
        st = None
        try:
            st = _SymmKnnState(N, F, k, group, dev)
            ok = 1.0
        except Exception as e:  # noqa: BLE001  (no NVLink peer access / no symmetric allocator: use the collective)
            import warnings
            warnings.warn(f"depthg_b200: symmetric-memory KNN exchange unavailable ({type(e).__name__}: {e}); "
                          "falling back to the NCCL all-gather")
            ok = 0.0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        _symm_states[key] = st if flag.item() == 1.0 else None
    return _symm_states[key]


def sharded_knn_build(local_feats: torch.Tensor, total_rows: int, k: int, group=None, return_stats: bool = False,
                      exchange: str = "auto"):
    """The query-row-sharded KNN build with the exchange overlapped: ``local_feats`` are rows
    ``knn_shard_bounds(total_rows, world, rank)`` of the normalised feature matrix.  Returns this rank's int64 [rows,k]
    block.  ``exchange``: "peer" = panel rows and fp32 rows pulled from the peers' symmetric memory with copy engines
    (no SMs) while this rank ranks its rows against the rows it holds; "nccl" = ``all_gather_into_tensor`` on a side
    stream; "auto" = peer copies when the shard is small enough for two phases to pay and symmetric memory is there."""
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
    # Two phases pay when the local pass is short enough to hide under the exchange: 128 query rows per CTA, one CTA
    # per SM, 256 database rows per tile at ~14 us a tile.  With few ranks the local rows are a large part of the
    # database and one pass over everything is faster.
    nblocks_max = -(-per // 128)
    nseg_local = max(1, min(112 // max(nblocks_max, 1), 8))
    two_phase = nblocks_max <= 112 and -(-per // 256) / nseg_local <= 26
    if exchange not in ("auto", "peer", "nccl"):
        raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
    if exchange == "peer" or (exchange == "auto" and two_phase):
        st = _symm_knn_state(total_rows, local_feats.shape[1], k, group, dev)
        if st is not None:
            return st.build(local_feats, return_stats)
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
    if two_phase and hi > lo:
        shard = KnnShard(local_feats, lo, total_rows, k).begin()
        cur.wait_stream(comm)
        return shard.finish(gathered[:total_rows], return_stats=return_stats)
    cur.wait_stream(comm)
    if hi == lo:
        idx = torch.empty((0, k), device=dev, dtype=torch.int64)
        return (idx, {"pipeline_error": 0, "fallback_rows": 0}) if return_stats else idx
    return knn_topk(local_feats, gathered[:total_rows], k, return_stats=return_stats)
