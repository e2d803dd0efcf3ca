"""The trainer's probe losses on the detached code map, on sm_100a kernels.

Reference call sites (``LitUnsupervisedSegmenter.training_step``)::

    detached_code = torch.clone(code.detach())                       # src/train_segmentation.py:426
    linear_logits = self.linear_probe(detached_code)                 # :429   nn.Conv2d(dim, n_classes, 1x1)
    linear_logits = F.interpolate(linear_logits, label.shape[-2:], mode='bilinear', align_corners=False)
    linear_logits = linear_logits.permute(0, 2, 3, 1).reshape(-1, self.n_classes)
    linear_loss = self.linear_probe_loss_fn(linear_logits[mask], flat_label[mask]).mean()   # :435
    cluster_loss, cluster_probs = self.cluster_probe(detached_code, None)                   # :441

``linear_probe_loss`` is the first five lines as one fused pass (``dg_linear_probe_ce``);
``ClusterLookup`` mirrors src/modules.py:646-675 (``dg_cluster_probe``).  The code map is
detached in the reference, so neither op returns a gradient for it, and both refuse an
input that still requires grad rather than silently dropping that gradient.  As everywhere
in this package there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, require_cuda_f32, stream_ptr

MAX_CLASSES = 32


def _workspace(code: torch.Tensor, K: int):
    B, D, h, w = code.shape
    n = _lib.lib().dg_probe_workspace_bytes(B, h, w, D, K)
    return torch.empty(n, device=code.device, dtype=torch.uint8), n


def _check_code(code: torch.Tensor, name: str):
    require_cuda_f32(code, name)
    if code.dim() != 4:
        raise ValueError(f"{name} must be [B,D,h,w]")
    if code.requires_grad:
        raise ValueError(f"{name} still requires grad: the reference feeds the probes code.detach() "
                         f"(src/train_segmentation.py:426) and these kernels return no gradient for it")


class _LinearProbeCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, code, weight, bias, label):
        B, D, h, w = code.shape
        K = weight.shape[0]
        want_grad = weight.requires_grad or (bias is not None and bias.requires_grad)
        w2 = weight.detach().reshape(K, D).contiguous()
        b1 = None if bias is None else bias.detach().contiguous()
        loss = torch.empty((), device=code.device, dtype=torch.float32)
        dw = torch.empty((K, D), device=code.device, dtype=torch.float32) if want_grad else None
        db = torch.empty((K,), device=code.device, dtype=torch.float32) if want_grad and bias is not None else None
        ws, ws_bytes = _workspace(code, K)
        check(_lib.lib().dg_linear_probe_ce(ptr(code), _lib.i64_array(code.stride()), B, D, h, w, ptr(w2), ptr(b1), K,
                                            ptr(label), _lib.i64_array(label.stride()), label.shape[1], label.shape[2],
                                            ptr(loss), ptr(dw), ptr(db), ptr(ws), ws_bytes,
                                            stream_ptr(code.device.index)), "dg_linear_probe_ce")
        ctx.save_for_backward(dw, db)
        ctx.wshape = weight.shape
        ctx.set_materialize_grads(False)
        return loss

    @staticmethod
    def backward(ctx, g):
        dw, db = ctx.saved_tensors
        if g is None:
            return None, None, None, None
        return (None, None if dw is None else (dw * g).reshape(ctx.wshape), None if db is None else db * g, None)


def linear_probe_loss(code: torch.Tensor, weight: torch.Tensor, bias, label: torch.Tensor) -> torch.Tensor:
    """Mean cross-entropy of the bilinearly upsampled linear-probe logits against ``label``
    (src/train_segmentation.py:419-437), fused: the [B,K,Hl,Wl] logits are never materialised.

    code   [B,D,h,w] CUDA fp32, detached (any strides);
    weight [K,D,1,1] or [K,D] and bias [K] or None: ``nn.Conv2d(dim, n_classes, (1,1))`` parameters;
    label  [B,Hl,Wl] int64; entries outside [0,K) are masked out as in the reference (:421-423)."""
    _check_code(code, "code")
    require_cuda_f32(weight, "weight")
    if bias is not None:
        require_cuda_f32(bias, "bias")
    if not isinstance(label, torch.Tensor) or not label.is_cuda or label.dtype != torch.int64:
        raise TypeError("label must be a CUDA int64 tensor")
    if label.dim() == 4 and label.shape[1] == 1:
        label = label[:, 0]
    if label.dim() != 3 or label.shape[0] != code.shape[0]:
        raise ValueError(f"label {tuple(label.shape)} must be [B,Hl,Wl] with B={code.shape[0]}")
    K = weight.shape[0]
    if weight.numel() != K * code.shape[1]:
        raise ValueError(f"weight {tuple(weight.shape)} does not match code dim {code.shape[1]}")
    if K > MAX_CLASSES:
        raise ValueError(f"{K} classes > {MAX_CLASSES} not supported")
    return _LinearProbeCEFn.apply(code, weight, bias, label)


class _ClusterProbeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, code, clusters):
        B, D, h, w = code.shape
        N = clusters.shape[0]
        c = clusters.detach().contiguous()
        loss = torch.empty((), device=code.device, dtype=torch.float32)
        probs = torch.empty((B, h, w, N), device=code.device, dtype=torch.float32)
        dc = torch.empty_like(c) if clusters.requires_grad else None
        ws, ws_bytes = _workspace(code, N)
        check(_lib.lib().dg_cluster_probe(ptr(code), _lib.i64_array(code.stride()), B, D, h, w, ptr(c), N, 0, 0.0,
                                          ptr(loss), ptr(probs), ptr(dc), ptr(ws), ws_bytes,
                                          stream_ptr(code.device.index)), "dg_cluster_probe")
        ctx.save_for_backward(dc)
        probs = probs.permute(0, 3, 1, 2)
        ctx.mark_non_differentiable(probs)
        ctx.set_materialize_grads(False)   # no [B,N,h,w] zero tensor for the non-differentiable probs
        return loss, probs

    @staticmethod
    def backward(ctx, g, _gp):
        (dc,) = ctx.saved_tensors
        return None, None if (dc is None or g is None) else dc * g


class ClusterLookup(nn.Module):
    """src/modules.py:646-675 with the same parameter (``clusters`` [n_classes, dim]), call
    signature and return values.  ``alpha=None`` (the training call) is differentiable w.r.t.
    ``clusters``; the soft-assignment calls used by evaluation (``alpha`` given, optionally
    ``log_probs=True``) are forward-only."""

    def __init__(self, dim: int, n_classes: int):
        super().__init__()
        if n_classes > MAX_CLASSES:
            raise ValueError(f"{n_classes} clusters > {MAX_CLASSES} not supported")
        self.n_classes = n_classes
        self.dim = dim
        self.clusters = torch.nn.Parameter(torch.randn(n_classes, dim))

    def reset_parameters(self):
        with torch.no_grad():
            self.clusters.copy_(torch.randn(self.n_classes, self.dim))

    def forward(self, x, alpha, log_probs=False):
        _check_code(x, "x")
        require_cuda_f32(self.clusters, "clusters")
        if x.shape[1] != self.clusters.shape[1]:
            raise ValueError(f"x has {x.shape[1]} channels, clusters have {self.clusters.shape[1]}")
        if alpha is None:
            if log_probs:   # the reference would evaluate inner_products * None
                raise TypeError("log_probs=True needs alpha")
            return _ClusterProbeFn.apply(x, self.clusters)
        if torch.is_grad_enabled() and self.clusters.requires_grad and not log_probs:
            raise NotImplementedError("the soft-assignment loss (alpha given) is forward-only here; the reference "
                                      "trainer calls the probe with alpha=None. Wrap the call in torch.no_grad()")
        B, D, h, w = x.shape
        N = self.n_classes
        c = self.clusters.detach().contiguous()
        out = torch.empty((B, h, w, N), device=x.device, dtype=torch.float32)
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        ws, ws_bytes = _workspace(x, N)
        check(_lib.lib().dg_cluster_probe(ptr(x), _lib.i64_array(x.stride()), B, D, h, w, ptr(c), N,
                                          2 if log_probs else 1, float(alpha), ptr(loss), ptr(out), None, ptr(ws),
                                          ws_bytes, stream_ptr(x.device.index)), "dg_cluster_probe")
        out = out.permute(0, 3, 1, 2)
        return out if log_probs else (loss, out)
