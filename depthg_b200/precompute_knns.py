"""KNN build of the reference's ``precompute_knns.py`` on sm_100a kernels.

The reference (/root/reference/src/precompute_knns.py:94-115) pools + normalises
backbone features, then on the CPU computes a [N/64, N] similarity block per chunk
and takes ``topk(…, 30)``.  Here the similarity GEMM and the running top-k are one
kernel (``dg_knn_topk``); the similarity matrix is never materialised.  Query rows
shard across GPUs; the database is all-gathered (``depthg_b200.distributed``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, require_cuda_f32, stream_ptr

TOPK = 30          # src/precompute_knns.py:108
N_BATCHES = 64     # src/precompute_knns.py:51 (chunking of the reference loop; the kernel streams instead)


def knn_topk(queries: torch.Tensor, db: torch.Tensor, k: int = TOPK, return_sims: bool = False,
             return_stats: bool = False):
    """Top-k database rows by fp32 dot product for every query row.

    queries [Nq,F], db [N,F] CUDA fp32 (unit-norm rows for cosine similarity).
    Returns int64 [Nq,k] sorted by descending similarity (and the similarities).  ``return_stats`` appends a dict
    with the tensor-core path's diagnostics (synchronises): ``fallback_rows`` = query rows whose candidate list
    failed the certificate and were recomputed by the exact fp32 kernel, ``pipeline_error`` (0 = ok)."""
    require_cuda_f32(queries, "queries")
    require_cuda_f32(db, "db")
    if queries.dim() != 2 or db.dim() != 2 or queries.shape[1] != db.shape[1]:
        raise ValueError(f"queries {tuple(queries.shape)} / db {tuple(db.shape)} must be [Nq,F] and [N,F]")
    if not 1 <= k <= 32:
        raise ValueError(f"k={k} must be in [1,32] (the reference uses 30)")
    if k > db.shape[0]:
        raise ValueError(f"k={k} exceeds the database size {db.shape[0]}")
    queries, db = queries.contiguous(), db.contiguous()
    Nq, F = queries.shape
    N = db.shape[0]
    if Nq == 0:
        return torch.empty((0, k), device=db.device, dtype=torch.int64)
    idx = torch.empty((Nq, k), device=db.device, dtype=torch.int64)
    sims = torch.empty((Nq, k), device=db.device, dtype=torch.float32) if return_sims else None
    lib = _lib.lib()
    ws_bytes = lib.dg_knn_workspace_bytes(Nq, N, F, k)
    ws = torch.empty(ws_bytes, device=db.device, dtype=torch.uint8)
    if queries.device != db.device:
        raise ValueError(f"queries on {queries.device}, db on {db.device}")
    with torch.cuda.device(db.device):
        check(lib.dg_knn_topk(ptr(queries), ptr(db), Nq, N, F, k, ptr(idx), ptr(sims), ptr(ws), ws_bytes,
                              stream_ptr(db.device.index)), "dg_knn_topk")
    out = (idx, sims) if return_sims else (idx,)
    if return_stats:
        head = ws[:12].view(torch.int32).cpu()
        out = out + ({"pipeline_error": int(head[0]), "fallback_rows": int(head[1])},)
    return out if len(out) > 1 else out[0]


class KnnShard:
    """Phased build of one query-row shard (``dg_knn_shard_begin`` / ``dg_knn_shard_finish``): the shard's rows are
    database rows [row_lo, row_lo + Nq).  ``begin`` needs only the local rows and is what a rank runs while the
    exchange of the database is in flight; ``finish`` takes the gathered database.  Same result as
    ``knn_topk(local, db, k)``.  ``ws``: an externally owned workspace (e.g. symmetric memory the peers copy panel rows
    out of); ``phases`` as in include/depthg_b200.h."""

    def __init__(self, local: torch.Tensor, row_lo: int, total_rows: int, k: int = TOPK, ws: torch.Tensor = None):
        require_cuda_f32(local, "local")
        if local.dim() != 2:
            raise ValueError(f"local must be [Nq,F], got {tuple(local.shape)}")
        if not 1 <= k <= 32 or k > total_rows:
            raise ValueError(f"k={k} must be in [1,32] and at most the database size {total_rows}")
        if row_lo < 0 or row_lo + local.shape[0] > total_rows:
            raise ValueError(f"rows [{row_lo},{row_lo + local.shape[0]}) outside the {total_rows}-row database")
        self.local, self.row_lo, self.N, self.k = local.contiguous(), int(row_lo), int(total_rows), int(k)
        self.Nq, self.F = self.local.shape
        self.ws, self.ws_bytes = ws, 0
        if self.Nq:
            self.ws_bytes = _lib.lib().dg_knn_workspace_bytes(self.Nq, self.N, self.F, self.k)
            if ws is None:
                self.ws = torch.empty(self.ws_bytes, device=local.device, dtype=torch.uint8)
            elif ws.numel() < self.ws_bytes or ws.dtype != torch.uint8 or ws.device != local.device:
                raise ValueError("ws must be a uint8 tensor of at least dg_knn_workspace_bytes on the rows' device")

    def begin(self, phases: int = _lib.KNN_BEGIN_ALL):
        if self.Nq:
            dev = self.local.device
            with torch.cuda.device(dev):
                check(_lib.lib().dg_knn_shard_begin(ptr(self.local), self.Nq, self.row_lo, self.N, self.F, self.k,
                                                    ptr(self.ws), self.ws_bytes, phases, stream_ptr(dev.index)),
                      "dg_knn_shard_begin")
        return self

    def finish(self, db: torch.Tensor, return_stats: bool = False, phases: int = _lib.KNN_FINISH_ALL,
               npeer_max: int = 0, idx: torch.Tensor = None):
        dev = self.local.device
        if db is not None:
            require_cuda_f32(db, "db")
            if db.dim() != 2 or db.shape[0] != self.N or db.shape[1] != self.F or not db.is_contiguous():
                raise ValueError(f"db must be a contiguous [{self.N},{self.F}] tensor, got {tuple(db.shape)}")
            if db.device != dev:
                raise ValueError(f"db on {db.device}, local rows on {dev}")
        if idx is None and (phases & _lib.KNN_RERANK):
            idx = torch.empty((self.Nq, self.k), device=dev, dtype=torch.int64)
        if self.Nq:
            with torch.cuda.device(dev):
                check(_lib.lib().dg_knn_shard_finish(ptr(db), self.Nq, self.row_lo, self.N, self.F, self.k, ptr(idx),
                                                     None, ptr(self.ws), self.ws_bytes, phases, npeer_max,
                                                     stream_ptr(dev.index)), "dg_knn_shard_finish")
        if return_stats:
            head = self.ws[:12].view(torch.int32).cpu() if self.Nq else [0, 0]
            return idx, {"pipeline_error": int(head[0]), "fallback_rows": int(head[1])}
        return idx


def knn_panel_layout(total_rows: int, F: int):
    """(hi_offset, lo_offset, row_bytes) of the bf16 panels inside a KNN workspace (dg_knn_panel_layout)."""
    import ctypes as C
    hi, lo, rb = C.c_size_t(), C.c_size_t(), C.c_size_t()
    check(_lib.lib().dg_knn_panel_layout(total_rows, F, C.byref(hi), C.byref(lo), C.byref(rb)), "dg_knn_panel_layout")
    return hi.value, lo.value, rb.value


def build_knn_index(normed_feats: torch.Tensor, k: int = TOPK, n_batches: int = N_BATCHES) -> torch.Tensor:
    """The whole loop of src/precompute_knns.py:99-113 as one call: int64 [N,k].
    ``n_batches`` is accepted for signature parity; chunking does not change the
    result and the kernel streams the database tile by tile instead."""
    del n_batches
    return knn_topk(normed_feats, normed_feats, k)


def pool_normalize(feat_maps: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """get_feats' ``F.normalize(model(img).mean([2,3]), dim=1)`` (src/precompute_knns.py:19)
    as one kernel.  feat_maps [N,C,H,W] (any strides) -> [N,C]."""
    require_cuda_f32(feat_maps, "feat_maps")
    if feat_maps.dim() != 4:
        raise ValueError("feat_maps must be [N,C,H,W]")
    N, Cdim, H, W = feat_maps.shape
    out = torch.empty((N, Cdim), device=feat_maps.device, dtype=torch.float32)
    with torch.cuda.device(feat_maps.device):
        check(_lib.lib().dg_pool_normalize(ptr(feat_maps), _lib.i64_array(feat_maps.stride()), N, Cdim, H, W, eps,
                                           ptr(out), stream_ptr(feat_maps.device.index)), "dg_pool_normalize")
    return out


PRECOMPUTE_RES = 392   # the script hard-codes res = 392 (src/precompute_knns.py:50); the dataset looks up cfg.res


def nns_filename(model_type, dataset_name, image_set, crop_type, res) -> str:
    """File name contract of src/precompute_knns.py:71-72 (producer) / src/data.py:1056-1057 (consumer).  Note the
    producer formats its hard-coded ``res`` (392) while the consumer formats ``cfg.res``: the two meet only when the
    training config says res = 392, exactly as in the reference."""
    return "nns_{}_{}_{}_{}_{}.npz".format(model_type, dataset_name, image_set, crop_type, res)


def save_nns(path: str, nearest_neighbors: torch.Tensor) -> None:
    """np.savez_compressed(file, nns=int64[N,k]) (src/precompute_knns.py:115): the one array ``ContrastiveSegDataset``
    reads back (``np.load(file)["nns"]``, src/data.py:1062-1063; row ``ind``, columns 1..num_neighbors, :1079)."""
    nn = nearest_neighbors.to("cpu", torch.int64)
    if nn.dim() != 2:
        raise ValueError(f"nearest_neighbors must be [N,k], got {tuple(nn.shape)}")
    np.savez_compressed(path, nns=nn.numpy())


def load_nns(path: str) -> np.ndarray:
    """The consumer side (src/data.py:1062-1064): ``np.load(feature_cache_file)["nns"]``."""
    return np.load(path)["nns"]


def precompute_and_save(feat_maps_or_feats: torch.Tensor, data_dir: str, model_type, dataset_name, image_set,
                        crop_type, res=PRECOMPUTE_RES, k: int = TOPK) -> str:
    """One (crop_type, image_set, dataset) iteration of the script's loop (src/precompute_knns.py:66-116) with the
    feature matrix already extracted: pool + normalise if given [N,C,H,W] maps (``get_feats``, :15-21), build the
    index on the GPU, write ``<data_dir>/nns/nns_<model>_<dataset>_<set>_<crop>_<res>.npz``.  Returns the path."""
    import os
    feats = pool_normalize(feat_maps_or_feats) if feat_maps_or_feats.dim() == 4 else feat_maps_or_feats
    nn = build_knn_index(feats, k)
    os.makedirs(os.path.join(data_dir, "nns"), exist_ok=True)
    path = os.path.join(data_dir, "nns", nns_filename(model_type, dataset_name, image_set, crop_type, res))
    save_nns(path, nn)
    return path
