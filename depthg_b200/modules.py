"""Reference-facing Python surface of the DepthG hot path, backed by sm_100a kernels.

Mirrors the names the reference's trainer imports from ``src/modules.py``
(``ContrastiveCorrelationLoss``, ``norm``, ``sample``, ``tensor_correlation``,
``farthest_point_sampling_depth``, ``super_perm``; /root/reference/src/modules.py:
789-825, 999-1037, 1184-1188, 1221-1367) so ``train_segmentation.py`` can switch
with one import line.  Every function launches hand-written CUDA
kernels through the C ABI of ``include/depthg_b200.h``; nothing here computes on
the CPU and there is no PyTorch fallback for the kernels.
"""
from __future__ import annotations

import ctypes as C
import struct

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, require_cuda_f32, stream_ptr

# 2*tan(90/2) evaluated in fp32 with the angle in RADIANS, as the reference does
# (depth2points is called with fov=90, src/modules.py:988-989, :1016).
FOV_FACTOR = struct.unpack("<f", struct.pack("<I", 0x404F54CB))[0]
FAR_PLANE = 5.0
NORM_EPS = 1e-10


# --------------------------------------------------------------------------- sampling
def _fps(depth: torch.Tensor, depth_b, H: int, W: int, S: int, affine: bool, want_idx: bool):
    require_cuda_f32(depth, "depth")
    if depth.dim() != 4 or depth.shape[1] != 1:
        raise ValueError(f"depth must be [B,1,Hd,Wd], got {tuple(depth.shape)}")
    B, _, Hd, Wd = depth.shape
    if S * S > H * W:
        raise ValueError(f"feature_samples**2 = {S * S} exceeds the {H}x{W} feature grid")
    depth = depth.contiguous()
    nimg = B
    if depth_b is not None:
        require_cuda_f32(depth_b, "depth_pos")
        if depth_b.shape != depth.shape:
            raise ValueError("depth and depth_pos must have the same shape")
        depth_b = depth_b.contiguous()
        nimg = 2 * B
    coords = torch.empty((nimg, S, S, 2), device=depth.device, dtype=torch.float32)
    idx = torch.empty((nimg, S * S), device=depth.device, dtype=torch.int32) if want_idx else None
    check(_lib.lib().dg_fps_coords(ptr(depth), ptr(depth_b), B, Hd, Wd, H, W, S, FOV_FACTOR, FAR_PLANE,
                                   1 if affine else 0, ptr(coords), ptr(idx), stream_ptr()), "dg_fps_coords")
    return coords, idx


def farthest_point_sampling_depth(t: torch.Tensor, depth: torch.Tensor, n_samples: int, include_feats: bool = False,
                                  gpu: bool = False, batched: bool = False) -> torch.Tensor:
    """Drop-in for src/modules.py:999-1037: [B,S,S,2] coordinates in [0,1), raster
    order.  ``include_feats`` / ``gpu`` / ``batched`` are accepted and ignored exactly
    as the reference ignores them (the batched variant is equivalent for square maps)."""
    coords, _ = _fps(depth, None, t.shape[-2], t.shape[-1], int(n_samples), affine=False, want_idx=False)
    return coords


def fps_index_sets(depth: torch.Tensor, H: int, W: int, n_samples: int) -> torch.Tensor:
    """Raster-sorted flat indices [B,S*S] (int32) of the selected grid points."""
    _, idx = _fps(depth, None, H, W, int(n_samples), affine=False, want_idx=True)
    return idx


def super_perm(size: int, device) -> torch.Tensor:
    """src/modules.py:1184-1188: randperm with fixed points bumped by one, mod size.
    ``randperm`` is the same torch call (same RNG stream for a shared seed); the bump is
    written as an add of the fixed-point mask instead of the reference's boolean-mask
    ``perm[perm == arange] += 1`` — identical values, but no nonzero() and therefore no
    device->host sync."""
    perm = torch.randperm(size, device=device, dtype=torch.long)
    return (perm + (perm == torch.arange(size, device=device))) % size


def super_perms(n: int, size: int, device) -> torch.Tensor:
    """``n`` successive super_perm draws as one [n,size] tensor (one fix-up for all)."""
    perm = torch.stack([torch.randperm(size, device=device, dtype=torch.long) for _ in range(n)])
    return (perm + (perm == torch.arange(size, device=device))) % size


def _strides(t: torch.Tensor):
    return _lib.i64_array(t.stride())


def _gather(t, coords, S, set_coord, set_slot, perm, eps, Prows, ld, out, rnorm, meanvec, fmt=_lib.PANEL_F32,
            out_lo=None, outT_hi=None, outT_lo=None):
    B, Cdim, H, W = t.shape
    check(_lib.lib().dg_gather_norm(ptr(t), _strides(t), B, Cdim, H, W, ptr(coords), S, len(set_coord),
                                    _lib.i32_array(set_coord), _lib.i32_array(set_slot), ptr(perm), eps, Prows, ld,
                                    fmt, ptr(out), ptr(out_lo), ptr(outT_hi), ptr(outT_lo), ptr(rnorm), ptr(meanvec),
                                    stream_ptr()), "dg_gather_norm")


def corr_kernel_choice(P: int, D: int) -> str:
    """Which correlation kernel a shape gets: the tcgen05 kernel needs S*S <= 128 and dim <= 128;
    everything else runs the generic CUDA-core kernel.  DEPTHG_B200_CORR=simt forces the generic one."""
    import os
    if os.environ.get("DEPTHG_B200_CORR", "") == "simt":
        return "simt"
    return "umma" if (P <= 128 and _lib.panel_ld(D) <= 128) else "simt"


def _check_coords(coords, B):
    require_cuda_f32(coords, "coords")
    if coords.dim() != 4 or coords.shape[0] != B or coords.shape[-1] != 2 or coords.shape[1] != coords.shape[2]:
        raise ValueError(f"coords must be [B,S,S,2], got {tuple(coords.shape)}")
    return coords.shape[1]


def _sample_panel(t: torch.Tensor, coords: torch.Tensor, eps: float):
    require_cuda_f32(t, "t")
    B, Cdim, H, W = t.shape
    S = _check_coords(coords, B)
    P, Prows, ld = S * S, _lib.panel_rows(S * S), _lib.panel_ld(Cdim)
    out = torch.empty((1, B, Prows, ld), device=t.device, dtype=torch.float32)
    rn = torch.empty((1, B, Prows), device=t.device, dtype=torch.float32)
    _gather(t, coords.contiguous().view(1, B, P, 2), S, [0], [0], None, eps, Prows, ld, out, rn, None)
    return out[0, :, :P, :Cdim], rn[0, :, :P], S


def sample(t: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """Drop-in for src/modules.py:822-825 (forward only): bilinear gather with the
    reference's S-axis swap.  Returns [B,C,S,S]."""
    panel, rn, S = _sample_panel(t, coords, NORM_EPS)
    B, _, Cdim = panel.shape
    # the kernel normalises; undo it with the stored 1/max(||x||,eps) to return raw samples
    raw = panel / rn.unsqueeze(-1).clamp_min(1e-30)
    return raw.permute(0, 2, 1).reshape(B, Cdim, S, S)


def sample_norm(t: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """norm(sample(t, coords)) in one kernel — the form the loss consumes."""
    panel, _, S = _sample_panel(t, coords, NORM_EPS)
    B, _, Cdim = panel.shape
    return panel.permute(0, 2, 1).reshape(B, Cdim, S, S)


def norm(t: torch.Tensor) -> torch.Tensor:
    """Drop-in for src/modules.py:789-790: L2-normalise over dim 1, eps 1e-10
    (forward only; the trainer imports the name but the hot path never calls it
    outside the loss).  Runs the gather kernel at the exact grid points."""
    require_cuda_f32(t, "t")
    if t.dim() != 4:
        raise ValueError("norm expects [B,C,H,W]")
    B, Cdim, H, W = t.shape
    if H != W:
        raise ValueError("norm: only square maps are supported by the panel kernel")
    ys = torch.arange(H, device=t.device, dtype=torch.float32)
    g = (ys / (H - 1) * 2 - 1) if H > 1 else torch.zeros(1, device=t.device)
    # coords[b,a,c] is read at output (h=c, w=a): x <- coords[...,0], y <- coords[...,1]
    cx = g.view(H, 1).expand(H, H)   # index a -> x = a
    cy = g.view(1, H).expand(H, H)   # index c -> y = c
    coords = torch.stack([cx, cy], -1).unsqueeze(0).expand(B, H, H, 2).contiguous()
    return sample_norm(t, coords)


def tensor_correlation(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Drop-in for src/modules.py:797-809 (forward only): einsum('nchw,ncij->nhwij')
    through the correlation kernel (pair loss disabled, dense cd output)."""
    require_cuda_f32(a, "a")
    require_cuda_f32(b, "b")
    n, c, h, w = a.shape
    if b.shape != a.shape or h != w:
        raise ValueError("tensor_correlation: operands must share one square [N,C,S,S] shape")
    P, Prows, ld = h * w, _lib.panel_rows(h * w), _lib.panel_ld(c)
    if ld > 128:
        raise ValueError("tensor_correlation: the standalone form supports C <= 128 (the loss kernel has no such limit)")
    dev = a.device
    cn = torch.zeros((2, n, Prows, ld), device=dev, dtype=torch.float32)
    cn[0, :, :P, :c] = a.permute(0, 2, 3, 1).reshape(n, P, c)
    cn[1, :, :P, :c] = b.permute(0, 2, 3, 1).reshape(n, P, c)
    fn = torch.zeros((2, n, Prows, 32), device=dev, dtype=torch.float32)
    out8 = torch.empty(8, device=dev, dtype=torch.float32)
    dC1 = torch.empty((3, n, Prows, ld), device=dev, dtype=torch.float32)
    dC2 = torch.empty_like(dC1)
    cd = torch.empty((2, n, P, P), device=dev, dtype=torch.float32)
    ws_bytes = _lib.lib().dg_corr_loss_workspace_bytes(2, n, P)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    pan = _lib.make_panels(_lib.PANEL_F32, fn, None, cn, None, None, None)
    check(_lib.lib().dg_corr_loss(C.byref(pan), None, None, 2, n, P, Prows, 32, 32, c, ld,
                                  _lib.f32_array([0.0, 0.0]), _lib.i32_array([_lib.GROUP_INTRA, _lib.GROUP_INTER]),
                                  0.0, 0, ptr(out8), ptr(dC1), ptr(dC2), ptr(cd), None, None, None, ptr(ws), ws_bytes,
                                  stream_ptr()), "dg_corr_loss")
    return cd[1].reshape(n, h, w, h, w)


# --------------------------------------------------------------------------- the loss
class _CorrLossFn(torch.autograd.Function):
    """forward: FPS/gather/normalise -> fused correlation loss (values + unit
    gradients); backward: weight the unit gradients by the upstream scalars and
    scatter them back through normalise + bilinear gather into the code tensors."""

    debug_fd = False        # tests: also dump the raw feature correlations of the tcgen05 kernel
    last_fd = None
    last_unit_grads = None

    @staticmethod
    def forward(ctx, feats, feats_pos, code, code_pos, depth, coords, perms, S, shifts, depth_shift, flags,
                materialize):
        lib = _lib.lib()
        B, Cdim, H, W = feats.shape
        D = code.shape[1]
        dev = feats.device
        nneg = 0 if perms is None else perms.shape[0]
        npairs = 2 + nneg
        P, Prows = S * S, _lib.panel_rows(S * S)
        ldf, ldc = _lib.panel_ld(Cdim), _lib.panel_ld(D)
        pointwise = bool(flags & _lib.FLAG_POINTWISE)
        has_depth = depth is not None

        f32 = dict(device=dev, dtype=torch.float32)
        bf16 = dict(device=dev, dtype=torch.bfloat16)
        kind = corr_kernel_choice(P, D)
        if kind == "umma":
            Prows = 128   # the tcgen05 kernel works on whole 128 x 128 tiles
        frn = torch.empty((npairs, B, Prows), **f32)
        fmean = torch.empty((npairs, B, ldf), **f32) if pointwise else None
        crn = torch.empty((npairs, B, Prows), **f32)

        # sets gathered from the "own" tensors: slot 0 at coords1, one slot per negative at coords2
        own_coord = [0] + [1] * nneg
        own_slot = [0] + list(range(2, 2 + nneg))
        if nneg:
            ident = torch.arange(B, device=dev, dtype=torch.long).unsqueeze(0)
            perm_all = torch.cat([ident, perms.to(torch.long)], 0).contiguous()
        else:
            perm_all = None
        if kind == "umma":
            f_hi, f_lo = torch.empty((npairs, B, Prows, ldf), **bf16), torch.empty((npairs, B, Prows, ldf), **bf16)
            c_hi, c_lo = torch.empty((npairs, B, Prows, ldc), **f32), torch.empty((npairs, B, Prows, ldc), **f32)
            ct_hi, ct_lo = torch.empty((npairs, B, 128, 128), **bf16), torch.empty((npairs, B, 128, 128), **bf16)
            fk = dict(fmt=_lib.PANEL_FEATS_SPLIT, out_lo=f_lo)
            ck = dict(fmt=_lib.PANEL_CODE_SPLIT, out_lo=c_lo, outT_hi=ct_hi, outT_lo=ct_lo)
            pan = _lib.make_panels(_lib.PANEL_CODE_SPLIT, f_hi, f_lo, c_hi, c_lo, ct_hi, ct_lo)
        else:
            f_hi, c_hi = torch.empty((npairs, B, Prows, ldf), **f32), torch.empty((npairs, B, Prows, ldc), **f32)
            c_lo = None
            fk, ck = {}, {}
            pan = _lib.make_panels(_lib.PANEL_F32, f_hi, None, c_hi, None, None, None)
        _gather(feats, coords, S, own_coord, own_slot, perm_all, NORM_EPS, Prows, ldf, f_hi, frn, fmean, **fk)
        _gather(feats_pos, coords, S, [1], [1], None, NORM_EPS, Prows, ldf, f_hi, frn, fmean, **fk)
        _gather(code, coords, S, own_coord, own_slot, perm_all, NORM_EPS, Prows, ldc, c_hi, crn, None, **ck)
        _gather(code_pos, coords, S, [1], [1], None, NORM_EPS, Prows, ldc, c_hi, crn, None, **ck)

        dsign = None
        if has_depth:
            dsign = torch.empty((B, Prows), **f32)
            _, _, Hd, Wd = depth.shape
            check(lib.dg_depth_sign(ptr(depth), B, Hd, Wd, S, NORM_EPS, Prows, ptr(dsign), stream_ptr()),
                  "dg_depth_sign")

        groups = [_lib.GROUP_INTRA, _lib.GROUP_INTER] + [_lib.GROUP_NEG] * nneg
        out8 = torch.empty(8, **f32)
        dC1 = torch.empty((npairs + 1, B, Prows, ldc), **f32)
        dC2 = torch.empty((npairs + 1, B, Prows, ldc), **f32)
        cd_out = torch.empty((npairs, B, P, P), **f32) if materialize else None
        loss_out = torch.empty((npairs, B, P, P), **f32) if materialize else None
        dd_out = torch.empty((B, P, P), **f32) if (materialize and has_depth) else None
        ws_bytes = lib.dg_corr_loss_workspace_bytes(npairs, B, P)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        fd_dbg = None
        if _CorrLossFn.debug_fd and kind == "umma":
            fd_dbg = torch.zeros((npairs, B, 128, 128), **f32)
            _CorrLossFn.last_fd = fd_dbg
        check(lib.dg_corr_loss(C.byref(pan), ptr(fmean), ptr(dsign), npairs, B, P, Prows, Cdim, ldf, D, ldc,
                               _lib.f32_array(shifts), _lib.i32_array(groups), float(depth_shift), int(flags),
                               ptr(out8), ptr(dC1), ptr(dC2), ptr(cd_out), ptr(loss_out), ptr(dd_out), ptr(fd_dbg),
                               ptr(ws), ws_bytes, stream_ptr()), "dg_corr_loss")
        _CorrLossFn.last_unit_grads = (dC1, dC2)

        ctx.save_for_backward(coords, perm_all, c_hi, c_lo, crn, dC1, dC2)
        ctx.meta = (S, Prows, ldc, npairs, nneg, groups, has_depth, code.shape, code_pos.shape,
                    code.stride(), code_pos.stride())
        ctx.code_like = (code, code_pos)  # only for zeros_like layout; no extra memory
        outs = [out8[0], out8[2], out8[4], out8[6], out8.detach()]
        dense = [t for t in (cd_out, loss_out, dd_out)]
        ctx.mark_non_differentiable(outs[4], *[t for t in dense if t is not None])
        return (*outs, *dense)

    @staticmethod
    def backward(ctx, g_intra, g_inter, g_neg, g_depth, *unused):
        coords, perm_all, cpan, cpan_lo, crn, dC1, dC2 = ctx.saved_tensors
        S, Prows, ldc, npairs, nneg, groups, has_depth, shp, shp_pos, _, _ = ctx.meta
        code, code_pos = ctx.code_like
        dev = cpan.device
        zero = torch.zeros((), device=dev, dtype=torch.float32)
        gw = torch.stack([g if g is not None else zero for g in (g_intra, g_inter, g_neg, g_depth)]).float()
        scales = [1.0, 1.0] + [1.0 / max(nneg, 1)] * nneg
        lib = _lib.lib()
        B, D, H, W = shp
        need = ctx.needs_input_grad
        d_code = d_code_pos = None
        common = (NORM_EPS, Prows, ldc, ptr(cpan), ptr(cpan_lo), ptr(crn), ptr(dC1), ptr(dC2), npairs,
                  _lib.i32_array(groups),
                  _lib.f32_array(scales), 1 if has_depth else 0, ptr(gw), stream_ptr())
        if need[2]:
            d_code = torch.zeros_like(code)
            own_coord = [0] + [1] * nneg
            own_slot = [0] + list(range(2, 2 + nneg))
            check(lib.dg_gather_norm_bwd(ptr(d_code), _strides(d_code), B, D, H, W, ptr(coords), S, len(own_coord),
                                         _lib.i32_array(own_coord), _lib.i32_array(own_slot), ptr(perm_all), *common),
                  "dg_gather_norm_bwd")
        if need[3]:
            d_code_pos = torch.zeros_like(code_pos)
            check(lib.dg_gather_norm_bwd(ptr(d_code_pos), _strides(d_code_pos), B, D, H, W, ptr(coords), S, 1,
                                         _lib.i32_array([1]), _lib.i32_array([1]), None, *common),
                  "dg_gather_norm_bwd")
        return (None, None, d_code, d_code_pos) + (None,) * 8


class ContrastiveCorrelationLoss(nn.Module):
    """Drop-in for the reference's ``ContrastiveCorrelationLoss`` (src/modules.py:1221-1367).

    Same constructor, same ``forward`` signature, same cfg keys (re-read on every
    call, so the trainer's in-flight mutations of ``cfg.feature_samples`` /
    ``depth_sampling`` / ``depth_feat_shift`` take effect), same 8-/6-tuple, no
    parameters or buffers (state_dict stays empty).  One deliberate relaxation:
    unless ``materialize_cd`` is set, the ``*_cd`` entries and ``neg_inter_loss``
    are returned as 0-dim means instead of [.,S,S,S,S] tensors — the trainer only
    ever calls ``.mean()`` on them (src/train_segmentation.py:303-323), and never
    writing them is where the HBM traffic goes away.  Set ``materialize_cd=True``
    (e.g. on histogram steps) to get the reference's full tensors.
    """

    def __init__(self, cfg, materialize_cd: bool = False):
        super().__init__()
        self.cfg = cfg
        self.materialize_cd = materialize_cd
        # test hooks (CPU and CUDA RNG streams differ): same contract as the oracle's
        self.perm_fn = super_perm
        self.rand_fn = lambda shape, device: torch.rand(shape, device=device)
        self.last_coords = None

    def _flags(self):
        cfg = self.cfg
        return ((_lib.FLAG_POINTWISE if cfg.pointwise else 0) | (_lib.FLAG_ZERO_CLAMP if cfg.zero_clamp else 0) |
                (_lib.FLAG_STABALIZE if cfg.stabalize else 0))

    def forward(self, orig_feats, orig_feats_pos, orig_salience, orig_salience_pos, orig_code, orig_code_pos,
                depth=None, depth_pos=None):
        cfg = self.cfg
        for name, t in (("orig_feats", orig_feats), ("orig_feats_pos", orig_feats_pos), ("orig_code", orig_code),
                        ("orig_code_pos", orig_code_pos)):
            require_cuda_f32(t, name)
            if t.dim() != 4:
                raise ValueError(f"{name} must be [B,C,H,W]")
        B, Cdim, H, W = orig_feats.shape
        if orig_feats_pos.shape != orig_feats.shape or orig_code_pos.shape != orig_code.shape or \
                orig_code.shape[0] != B or orig_code.shape[2:] != orig_feats.shape[2:]:
            raise ValueError("feature / code tensors have inconsistent shapes")
        S = int(cfg.feature_samples)
        nneg = int(cfg.neg_samples)
        if S * S > H * W:
            raise ValueError(f"feature_samples**2 = {S * S} exceeds the {H}x{W} feature grid")
        if nneg > 0 and B < 2:
            raise ValueError("neg_samples > 0 needs a batch of at least 2 (super_perm would pair an image with itself)")
        if 2 + nneg > _lib.DG_MAX_PAIRS:
            raise ValueError(f"neg_samples={nneg} exceeds the supported {_lib.DG_MAX_PAIRS - 2}")
        if cfg.use_salience:
            raise NotImplementedError("use_salience sampling (sample_nonzero_locations, src/modules.py:1191-1204) "
                                      "is outside the accelerated path")
        dev = orig_feats.device
        if cfg.depth_sampling == "fps":
            if depth is None or depth_pos is None:
                raise ValueError("depth_sampling='fps' needs depth and depth_pos")
            coords, _ = _fps(depth, depth_pos, H, W, S, affine=True, want_idx=False)
            coords = coords.view(2, B, S, S, 2)
        elif cfg.depth_sampling in ("simple", "fps_depth_feat"):
            raise NotImplementedError(f"depth_sampling={cfg.depth_sampling!r} (simple_depth_informed_sampling / "
                                      "include_feats, src/modules.py:828-883, :1313-1317) is outside the accelerated path")
        else:
            shape = [B, S, S, 2]
            c1 = self.rand_fn(shape, dev) * 2 - 1
            c2 = self.rand_fn(shape, dev) * 2 - 1
            coords = torch.stack([c1, c2]).float().contiguous()
        self.last_coords = coords
        if not nneg:
            perms = None
        elif self.perm_fn is super_perm:
            perms = super_perms(nneg, B, dev)
        else:
            perms = torch.stack([self.perm_fn(B, dev) for _ in range(nneg)])

        depth_term = bool(cfg.depth_feat_correlation_loss)
        if depth_term:
            if depth is None:
                raise ValueError("depth_feat_correlation_loss=True needs depth")
            require_cuda_f32(depth, "depth")
            depth = depth.contiguous()
        shifts = [float(cfg.pos_intra_shift), float(cfg.pos_inter_shift)] + [float(cfg.neg_inter_shift)] * nneg
        res = _CorrLossFn.apply(orig_feats, orig_feats_pos, orig_code, orig_code_pos, depth if depth_term else None,
                                coords.view(2, B, S * S, 2), perms, S, shifts,
                                float(cfg.depth_feat_shift) if depth_term else 0.0, self._flags(),
                                bool(self.materialize_cd))
        intra, inter, neg, dloss, out8, cd_out, loss_out, dd_out = res
        if self.materialize_cd:
            five = (B, S, S, S, S)
            intra_cd, inter_cd = cd_out[0].view(five), cd_out[1].view(five)
            neg_cd = cd_out[2:].reshape(nneg * B, S, S, S, S)
            # value: the dense unreduced tensor; gradient: routed through its mean (exact for the
            # trainer's neg_inter_loss.mean(), src/train_segmentation.py:303)
            neg_dense = loss_out[2:].reshape(nneg * B, S, S, S, S)
            neg_loss = neg_dense + (neg - neg.detach())
            dd = dd_out.view(five) if depth_term else None
        else:
            intra_cd, inter_cd, neg_cd, dd = out8[1], out8[3], out8[5], out8[7]
            neg_loss = neg
        head = (intra, intra_cd, inter, inter_cd, neg_loss, neg_cd)
        if depth_term:
            return head + (dloss, dd)
        return head
