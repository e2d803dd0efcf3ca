"""depthg_b200 — B200-native (sm_100a) kernels for the DepthG hot path.

Public surface (mirrors the names of the reference's src/modules.py and
src/precompute_knns.py):

    from depthg_b200.modules import ContrastiveCorrelationLoss, norm, sample, \
        tensor_correlation, farthest_point_sampling_depth, super_perm
    from depthg_b200.precompute_knns import build_knn_index, knn_topk, pool_normalize
    from depthg_b200.probes import ClusterLookup, linear_probe_loss
"""
from . import _lib  # noqa: F401
from .modules import (ContrastiveCorrelationLoss, DepthContrastiveCorrelationLoss, farthest_point_sampling_depth, fps_index_sets, norm, sample,  # noqa: F401
                      sample_norm, super_perm, tensor_correlation, depth_correlation)
from .precompute_knns import build_knn_index, knn_topk, pool_normalize, save_nns  # noqa: F401
from .probes import ClusterLookup, linear_probe_loss  # noqa: F401

__all__ = ["ContrastiveCorrelationLoss", "DepthContrastiveCorrelationLoss", "farthest_point_sampling_depth", "fps_index_sets", "norm", "sample",
           "sample_norm", "super_perm", "tensor_correlation", "depth_correlation", "build_knn_index", "knn_topk", "pool_normalize",
           "save_nns", "ClusterLookup", "linear_probe_loss"]
