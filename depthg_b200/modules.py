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
import os
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
    with torch.cuda.device(depth.device):
        check(_lib.lib().dg_fps_coords(ptr(depth), ptr(depth_b), B, Hd, Wd, H, W, S, FOV_FACTOR, FAR_PLANE,
                                       1 if affine else 0, ptr(coords), ptr(idx), stream_ptr(depth.device.index)),
              "dg_fps_coords")
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


class _GraphedSuperPerms:
    """``super_perms`` captured once into a CUDA graph per (n, size, device).

    The ~25 small library kernels behind ``neg_samples`` x ``torch.randperm`` cost ~170 us of host
    time per step; replaying them as one graph costs ~10 us and — because PyTorch's CUDA generator is
    graph-safe (seed/offset are read from device memory at replay and the offset is advanced by the
    graph's total consumption) — produces the SAME permutations as the eager calls under the same seed
    (asserted in tests/test_gpu_parity.py).  Any failure to capture falls back to the eager calls."""

    _cache = {}

    _side = {}
    _events = {}

    @classmethod
    def draw(cls, n: int, size: int, device) -> torch.Tensor:
        device = torch.device(device)
        key = (n, size, device.index)
        entry = cls._cache.get(key)
        if entry is None:
            entry = cls._capture(n, size, device)
            cls._cache[key] = entry
        if entry is False or torch.cuda.is_current_stream_capturing():
            return super_perms(n, size, device)
        graph, out = entry
        graph.replay()
        return out.clone()   # the graph's output buffer is overwritten by the next replay

    @classmethod
    def draw_async(cls, n: int, size: int, device, also_wait=None):
        """Replay on a side stream so the ~25 tiny sampler kernels overlap FPS on the main stream.
        Returns (perms, event); the consumer stream must wait for the event before reading perms.
        ``also_wait``: a CUDA event the returned event additionally stands for (the side stream waits for it before
        recording), i.e. foreign work - a gradient all-reduce - that may run underneath FPS but must be over before
        the SM-filling gathers start (the library waits for the event right after launching FPS)."""
        device = torch.device(device)
        main = torch.cuda.current_stream(device)
        if cls._cache.get((n, size, device.index)) in (None, False) or torch.cuda.is_current_stream_capturing():
            return cls.draw(n, size, device), None      # (the caller then hands also_wait to the library itself)
        side = cls._side.get(device.index)
        if side is None:
            side = cls._side[device.index] = torch.cuda.Stream(device=device)
        ev = cls._events.get(device.index)
        if ev is None:   # one event per device: a pending wait keeps the record it was issued against
            ev = cls._events[device.index] = torch.cuda.Event()
        with torch.cuda.stream(side):
            perms = cls.draw(n, size, device)
            if also_wait is not None:
                side.wait_event(also_wait)
            ev.record(side)
        perms.record_stream(main)   # allocated on the side stream, consumed on the main one
        return perms, ev

    @staticmethod
    def _capture(n, size, device):
        gen = torch.cuda.default_generators[device.index]
        state = gen.get_state()
        try:
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(2):          # warm up allocator / cub temp storage outside the capture
                    super_perms(n, size, device)
            torch.cuda.current_stream(device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            # thread_local: CUDA calls of OTHER threads (DataLoader pin_memory thread, NCCL watchdog) neither fail
            # nor invalidate this capture
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out = super_perms(n, size, device)
            return graph, out
        except Exception:                   # noqa: BLE001 - capture is an optimisation only
            return False
        finally:
            gen.set_state(state)            # neither warm-up nor capture (failed or not) may consume the user's RNG stream


_side_streams = {}


def _side_stream(device):
    """One side stream per device for work that overlaps the caller's stream (prefetch_sampling)."""
    st = _side_streams.get(device.index)
    if st is None:
        st = _side_streams[device.index] = torch.cuda.Stream(device=device)
    return st


def _reserve_perm_stream(size: int, device):
    """(seed, offset) of the Philox stream the one-launch sampler uses for this draw; advances torch's CUDA generator
    past it, so ``torch.manual_seed`` makes runs reproducible and successive draws differ."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + 4 * ((size + 3) // 4))
    return seed, offset


def fused_super_perms(n: int, size: int, device) -> torch.Tensor:
    """All ``n`` negative-pair permutations in ONE kernel (dg_super_perms): same distribution as
    ``super_perm`` but a Philox stream of its own, keyed by torch's CUDA generator (seed, offset) so
    ``torch.manual_seed`` still makes runs reproducible.  The generator is advanced past the draws."""
    device = torch.device(device)
    seed, offset = _reserve_perm_stream(size, device)
    out = torch.empty((n, size), device=device, dtype=torch.long)
    check(_lib.lib().dg_super_perms(seed, offset, n, size, ptr(out), stream_ptr(device.index)), "dg_super_perms")
    return out


def sample_nonzero_locations(t: torch.Tensor, target_size, randint_fn=None) -> torch.Tensor:
    """Drop-in for src/modules.py:1191-1204 (the ``use_salience`` coordinate sampler): S*S of each image's non-zero
    salience pixels drawn with replacement — uniform pixels for an all-zero map — as (x, y) coordinates in [-1, 1).
    Index arithmetic on the device with torch ops (like ``super_perm``, this is RNG plumbing, not a kernel); the
    random draws are the reference's own ``torch.randint`` calls in the reference's order, so a shared seed gives the
    same coordinates: per image one CPU-generator draw of n indices (:1199), or one device draw of [n,2] (:1197)."""
    if t.dim() != 3:
        raise ValueError(f"salience must be [B,H,W], got {tuple(t.shape)}")
    if randint_fn is None:
        randint_fn = lambda high, size, device: (torch.randint(high, size=size) if device is None  # noqa: E731
                                                 else torch.randint(high, size=size, device=device))
    B, n = t.shape[0], int(target_size[1]) * int(target_size[2])
    nz = torch.nonzero(t)                                   # [m,3] rows (image, y, x), sorted by image
    counts = torch.bincount(nz[:, 0], minlength=B).tolist()  # one sync: the draw sizes depend on the counts
    rows, off = [], 0
    for i in range(B):
        if counts[i] == 0:
            rows.append(randint_fn(t.shape[1], (n, 2), t.device).to(t.device))
        else:
            pick = randint_fn(counts[i], (n,), None).to(t.device)
            rows.append(nz[off + pick, 1:])
        off += counts[i]
    # (a device-tensor divisor: with a Python scalar torch's CUDA kernel multiplies by the reciprocal, which is one ulp
    #  off the IEEE division the reference's CPU arithmetic — and its goldens — perform)
    h = torch.full((), float(t.shape[1]), device=t.device, dtype=torch.float32)
    coords = torch.stack(rows).reshape(B, int(target_size[1]), int(target_size[2]), 2).to(torch.float32) / h
    return torch.flip(coords * 2 - 1, dims=[-1])


def _strides(t: torch.Tensor):
    return _lib.i64_array(t.stride())


def _gather(t, coords, S, set_coord, set_slot, perm, eps, Prows, ld, out, rnorm, meanvec, fmt=_lib.PANEL_F32,
            out_lo=None, out16_hi=None, out16_lo=None):
    B, Cdim, H, W = t.shape
    with torch.cuda.device(t.device):
        check(_lib.lib().dg_gather_norm(ptr(t), _strides(t), B, Cdim, H, W, ptr(coords), S, len(set_coord),
                                        _lib.i32_array(set_coord), _lib.i32_array(set_slot), ptr(perm), eps, Prows, ld,
                                        fmt, ptr(out), ptr(out_lo), ptr(out16_hi), ptr(out16_lo), ptr(rnorm),
                                        ptr(meanvec), stream_ptr(t.device.index)), "dg_gather_norm")


def corr_kernel_choice(P: int, D: int) -> str:
    """Which correlation kernel a shape gets: the tcgen05 kernel needs S*S <= 1024 and dim <= 128
    (above 256 points it works on column groups of two 128-wide tiles); everything else runs the generic
    CUDA-core kernel.  DEPTHG_B200_CORR=simt forces the generic one."""
    import os
    if os.environ.get("DEPTHG_B200_CORR", "") == "simt":
        return "simt"
    return "umma" if (P <= 1024 and _lib.panel_ld(D) <= 128) else "simt"


def _check_coords(coords, B):
    require_cuda_f32(coords, "coords")
    if coords.dim() != 4 or coords.shape[0] != B or coords.shape[-1] != 2 or coords.shape[1] != coords.shape[2]:
        raise ValueError(f"coords must be [B,S,S,2], got {tuple(coords.shape)}")
    return coords.shape[1]


def _forward_only(name: str, *tensors):
    """The free functions below are forward-only launches: they return tensors without a grad_fn.  The reference
    trainer differentiates through ``norm`` / ``sample`` in its optional rec / aug-alignment / CRF losses
    (src/train_segmentation.py:396, 407, 416); shadowing those names with these would silently stop training the
    head, so an input that still requires grad is refused (same convention as depthg_b200.probes)."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise ValueError(f"depthg_b200.modules.{name} is forward-only and its input requires grad: call it under "
                         f"torch.no_grad() / on detached tensors, or keep the reference's differentiable {name} for "
                         "that call site (only the loss classes carry a backward, see INTEGRATION.md)")


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
    _forward_only("sample", t, coords)
    panel, rn, S = _sample_panel(t, coords, NORM_EPS)
    B, _, Cdim = panel.shape
    # the kernel normalises; undo it with the stored 1/max(||x||,eps) to return raw samples
    raw = panel / rn.unsqueeze(-1).clamp_min(1e-30)
    return raw.permute(0, 2, 1).reshape(B, Cdim, S, S)


def sample_norm(t: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """norm(sample(t, coords)) in one kernel — the form the loss consumes."""
    _forward_only("sample_norm", t, coords)
    panel, _, S = _sample_panel(t, coords, NORM_EPS)
    B, _, Cdim = panel.shape
    return panel.permute(0, 2, 1).reshape(B, Cdim, S, S)


def norm(t: torch.Tensor) -> torch.Tensor:
    """Drop-in for src/modules.py:789-790 (``F.normalize(t, dim=1, eps=1e-10)``, forward only): any tensor with at
    least two dimensions, any strides.  One launch of the strided normalise kernel (dg_norm_dim1)."""
    require_cuda_f32(t, "t")
    _forward_only("norm", t)
    if t.dim() < 2:
        raise ValueError("norm expects at least [N,C]")
    if t.numel() == 0:
        return torch.empty_like(t)
    N, Cdim = t.shape[0], t.shape[1]
    inner = t.numel() // (N * Cdim)
    # view as [N, C, inner] with element strides (sN, sC, sP): NCHW-contiguous or channels-last without a copy
    sp = 1
    if t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last):
        sp = Cdim
    elif not t.is_contiguous():
        t = t.contiguous()
    out = torch.empty_like(t)                    # preserve_format: same strides as the (dense) input
    st = t.stride()
    with torch.cuda.device(t.device):
        check(_lib.lib().dg_norm_dim1(ptr(t), N, Cdim, inner, st[0], st[1], sp, NORM_EPS, ptr(out),
                                      stream_ptr(t.device.index)), "dg_norm_dim1")
    return out


def tensor_correlation(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Drop-in for src/modules.py:797-809 (forward only): einsum('nchw,ncij->nhwij')
    through the correlation kernel (pair loss disabled, dense cd output)."""
    _forward_only("tensor_correlation", a, b)
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
    with torch.cuda.device(dev):
        check(_lib.lib().dg_corr_loss(C.byref(pan), None, None, 2, n, P, Prows, 32, 32, c, ld,
                                      _lib.f32_array([0.0, 0.0]), _lib.i32_array([_lib.GROUP_INTRA, _lib.GROUP_INTER]),
                                      0.0, 0, ptr(out8), ptr(dC1), ptr(dC2), ptr(cd), None, None, None, ptr(ws),
                                      ws_bytes, stream_ptr(dev.index)), "dg_corr_loss")
    return cd[1].reshape(n, h, w, h, w)


def depth_correlation(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Drop-in for src/modules.py:812-814: the same einsum on [N,1,S,S] depth maps (an outer product per image).
    Inside the loss this term is never materialised (the kernel multiplies the two depth signs in registers)."""
    return tensor_correlation(a, b)


# --------------------------------------------------------------------------- the loss
_compiled_mod = None
_plan_cache = {}
_prows_cache = {}


def compiled_binding():
    """The compiled autograd binding of dg_loss_forward / dg_loss_backward (csrc/torch_binding.cpp -> _C.so), or False.
    It does what ``_CorrLossFn`` below does through ctypes (struct fill, arena / output allocation, the autograd node)
    in C++, because at ~0.2 ms of GPU work per step the Python version of that plumbing paced the step.  Same C-ABI
    calls, same kernels; ``DEPTHG_B200_BINDING=ctypes`` (or a missing _C.so) selects the ctypes path."""
    global _compiled_mod
    if _compiled_mod is None:
        _compiled_mod = False
        if os.environ.get("DEPTHG_B200_BINDING", "") != "ctypes":
            try:
                _lib.lib()                       # _C.so links against libdepthg_b200.so: load it first, loudly
                from . import _C as mod          # noqa: N812
                _compiled_mod = mod
            except ImportError:
                pass
    return _compiled_mod


class _CorrLossFn(torch.autograd.Function):
    """forward: FPS / gathers / depth signs / fused correlation loss (values + unit gradients) in ONE
    C-ABI call over one arena allocation; backward: one call that weights the unit gradients by
    the upstream scalars and scatters them through normalise + bilinear gather into the code grads."""

    debug = False           # tests: keep views of the unit gradients / dump raw tcgen05 feature correlations
    last_coords_off = 0
    last_fd = None
    last_unit_grads = None

    @staticmethod
    def forward(ctx, feats, feats_pos, code, code_pos, depth, depth_pos, coords, perms, desc, materialize,
                perms_event=None, aug_feats=None, perm_gen=None):
        lib = _lib.lib()
        dev = feats.device
        plan = _lib.LossPlan()
        check(lib.dg_loss_plan(C.byref(desc), C.byref(plan)), "dg_loss_plan")
        # no zero tensors for outputs nobody differentiated (the arena alone would be a 151 MB fill per step)
        ctx.set_materialize_grads(False)
        B, S, nneg, npairs = desc.B, desc.S, desc.neg_samples, plan.npairs
        P = S * S
        has_depth = bool(desc.flags & _lib.FLAG_DEPTH_TERM)
        arena = torch.empty(plan.total, device=dev, dtype=torch.uint8)
        out8 = torch.empty(8, device=dev, dtype=torch.float32)
        f32 = dict(device=dev, dtype=torch.float32)
        cd_out = torch.empty((npairs, B, P, P), **f32) if materialize else None
        loss_out = torch.empty((npairs, B, P, P), **f32) if materialize else None
        dd_out = torch.empty((B, P, P), **f32) if (materialize and has_depth) else None
        fd_dbg = None
        if _CorrLossFn.debug and plan.kernel >= 1:
            fd_dbg = torch.zeros((npairs, B, plan.Prows, plan.Prows), **f32)
            _CorrLossFn.last_fd = fd_dbg
        io = _lib.LossIO()
        io.feats, io.feats_pos, io.code, io.code_pos = (feats.data_ptr(), feats_pos.data_ptr(), code.data_ptr(),
                                                        code_pos.data_ptr())
        io.feats_strides[:] = feats.stride()
        io.feats_pos_strides[:] = feats_pos.stride()
        io.code_strides[:] = code.stride()
        io.code_pos_strides[:] = code_pos.stride()
        io.depth = depth.data_ptr() if depth is not None else None
        io.depth_pos = depth_pos.data_ptr() if depth_pos is not None else None
        io.coords = coords.data_ptr() if coords is not None else None
        io.perms = perms.data_ptr() if perms is not None else None
        io.arena, io.out8 = arena.data_ptr(), out8.data_ptr()
        io.perms_ready = perms_event.cuda_event if perms_event is not None else None
        if perm_gen is not None:           # `perms` is an output: the forward draws the permutations itself
            io.gen_perms, io.perm_seed, io.perm_offset = 1, perm_gen[0], perm_gen[1]
        if aug_feats is not None:
            io.aug_feats = aug_feats.data_ptr()
            io.aug_feats_strides[:] = aug_feats.stride()
        for name, t in (("cd_out", cd_out), ("loss_out", loss_out), ("dd_out", dd_out), ("fd_dbg", fd_dbg)):
            setattr(io, name, t.data_ptr() if t is not None else None)
        check(lib.dg_loss_forward(C.byref(desc), C.byref(io), stream_ptr(dev.index)), "dg_loss_forward")

        ctx.desc, ctx.plan = desc, plan
        ctx.keep = (arena, coords, perms)          # the arena holds coords / panels / unit gradients for backward
        ctx.code_like = (code, code_pos)
        if _CorrLossFn.debug:   # views of the unit gradients, summed over their per-tile partial buffers
            if plan.kernel == 2:      # persistent tcgen05 kernel: dC2 per 128-row tile, dC1 per 128-column tile
                ni = nj = plan.Prows // 128
                niu = nju = -(-P // 128)
            elif plan.kernel == 1:    # round-1 tcgen05 kernel: dC1 per 256-column group above 256 points
                ni, niu = plan.Prows // 128, -(-P // 128)
                nj, nju = (plan.Prows // 256, -(-P // 256)) if P > 256 else (1, 1)
            else:
                ni = nj = niu = nju = 1
            n1 = (npairs + 1) * B * plan.Prows * plan.ldc
            d1 = arena[plan.dC1:plan.dC1 + n1 * nj * 4].view(torch.float32).view(npairs + 1, nj, B, plan.Prows,
                                                                                 plan.ldc)[:, :nju].sum(1)
            d2 = arena[plan.dC2:plan.dC2 + n1 * ni * 4].view(torch.float32).view(npairs + 1, ni, B, plan.Prows,
                                                                                 plan.ldc)[:, :niu].sum(1)
            _CorrLossFn.last_unit_grads = (d1, d2)
        else:
            _CorrLossFn.last_unit_grads = None
        # the coordinates FPS produced live at plan.coords in the arena; the module slices them lazily (last_coords)
        coords_src = arena if (desc.flags & _lib.FLAG_FPS) else out8.new_empty(0)
        ctx.coords_off = plan.coords
        _CorrLossFn.last_coords_off = plan.coords
        u = out8.unbind(0)               # one op instead of four index kernels-worth of host time
        outs = [u[0], u[2], u[4], u[6], out8.detach(), coords_src]
        dense = [cd_out, loss_out, dd_out]
        ctx.mark_non_differentiable(outs[4], outs[5], *[t for t in dense if t is not None])
        return (*outs, *dense)

    @staticmethod
    def backward(ctx, g_intra, g_inter, g_neg, g_depth, *unused):
        arena, coords, perms = ctx.keep
        code, code_pos = ctx.code_like
        need = ctx.needs_input_grad
        gr = _lib.LossGrads()
        keep = []
        for i, g in enumerate((g_intra, g_inter, g_neg, g_depth)):
            if g is not None:
                g = g.contiguous() if g.dtype == torch.float32 else g.float()
                keep.append(g)
                gr.g[i] = g.data_ptr()
        d_code = d_code_pos = None
        if need[2]:
            d_code = torch.zeros_like(code)
            gr.d_code = d_code.data_ptr()
            gr.d_code_strides[:] = d_code.stride()
        if need[3]:
            d_code_pos = torch.zeros_like(code_pos)
            gr.d_code_pos = d_code_pos.data_ptr()
            gr.d_code_pos_strides[:] = d_code_pos.stride()
        io = _lib.LossIO()
        io.arena = arena.data_ptr()
        io.coords = coords.data_ptr() if coords is not None else None
        io.perms = perms.data_ptr() if perms is not None else None
        with torch.cuda.device(arena.device):   # autograd may run this on a thread whose current device differs
            check(_lib.lib().dg_loss_backward(C.byref(ctx.desc), C.byref(io), C.byref(gr),
                                              stream_ptr(arena.device.index)), "dg_loss_backward")
        return (None, None, d_code, d_code_pos) + (None,) * 9


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

    # per-call bookkeeping attributes: plain Python values, set several times per forward.  nn.Module.__setattr__
    # walks its parameter / buffer / module checks for every assignment (~1.5 us each, ~15 us a step on a 180 us step);
    # these names go straight to the instance dict.
    _PLAIN_ATTRS = frozenset(("wait_after_fps", "last_perms", "_last_coords", "_next_job", "_presampled",
                              "last_used_presampled"))

    def __setattr__(self, name, value):
        if name in ContrastiveCorrelationLoss._PLAIN_ATTRS:
            object.__setattr__(self, name, value)
        else:
            super().__setattr__(name, value)

    def __init__(self, cfg, materialize_cd: bool = False, negative_sampler: str = "fused"):
        super().__init__()
        self.cfg = cfg
        self.materialize_cd = materialize_cd
        if negative_sampler not in ("torch", "fused"):
            raise ValueError("negative_sampler must be 'torch' (the reference's randperm stream) or 'fused'")
        # "fused" (default): one dg_super_perms launch - the same distribution (uniform permutations with the
        #          reference's fixed-point bump), drawn from a Philox stream keyed by torch's CUDA generator, so
        #          torch.manual_seed still makes runs reproducible; ~60 us less host time per step than
        # "torch": neg_samples x torch.randperm in the reference's order (src/modules.py:1341) - bit-identical
        #          permutations to the reference running on the same GPU / torch build under the same seed
        self.negative_sampler = negative_sampler
        # replay the torch.randperm sequence as one CUDA graph (same RNG stream, ~15x less host time)
        self.graph_negative_sampler = True
        # test hooks (CPU and CUDA RNG streams differ): same contract as the oracle's
        self.perm_fn = super_perm
        self.rand_fn = lambda shape, device: torch.rand(shape, device=device)
        self.randint_fn = None      # use_salience draws; None = the reference's torch.randint calls
        # Optional torch.cuda.Event of foreign work the next forward may overlap with its FPS kernel but must not
        # overlap with its gathers (a DDP-style gradient all-reduce of the previous step on another stream): the
        # forward waits for it after launching FPS.  Consumed (reset to None) by the call.
        self.wait_after_fps = None
        self.last_perms = None
        self._last_coords = None
        # sampling done ahead of the forward that uses it (queue_next_sampling / prefetch_sampling)
        self._next_job = None       # (depth, depth_pos) whose sampling rides in the NEXT forward call
        self._presampled = None     # dict(key, coords, dsign, perms, event) waiting for the forward of that batch

    @property
    def last_coords(self):
        """The [2,B,S,S,2] sample coordinates of the last call (test hook, same contract as the oracle's).  With FPS
        sampling they are a view into the call's arena, built on first access."""
        v = self._last_coords
        if isinstance(v, tuple):
            arena, off, B, S = v
            v = arena[off:off + 2 * B * S * S * 2 * 4].view(torch.float32).view(2, B, S, S, 2)
            self._last_coords = v
        return v

    @last_coords.setter
    def last_coords(self, value):
        self._last_coords = value

    @staticmethod
    def _depth4(depth):
        """[B,Hd,Wd] -> [B,1,Hd,Wd] (the Potsdam dataset yields depth without a channel axis, src/data.py:226), contiguous."""
        if depth is not None and depth.dim() == 3:
            depth = depth.unsqueeze(1)
        return depth.contiguous() if depth is not None else None

    @staticmethod
    def _sampling_key(depth, depth_pos, S, H, W, nneg):
        return (depth.data_ptr(), depth._version, tuple(depth.shape), depth_pos.data_ptr(), depth_pos._version,
                tuple(depth_pos.shape), S, H, W, nneg)

    def queue_next_sampling(self, depth, depth_pos):
        """Hand the NEXT batch's depth maps to the loss before calling ``forward`` on the CURRENT batch: the next
        batch's farthest-point sampling (coordinates, depth signs, negative permutations) then rides as extra CTAs of
        this forward's correlation kernel, on SMs its item list leaves idle (dg_loss_io_t::next_*), and the forward that
        later receives exactly these tensors starts at its gathers — FPS (src/modules.py:999-1037, 14 % of a step) is
        off the critical path.  A forward that receives other depth tensors simply samples as usual.  Only
        ``depth_sampling in ('fps', 'fps_depth_feat')`` uses it."""
        if depth is None or depth_pos is None:
            self._next_job = None
            return
        for name, t in (("depth", depth), ("depth_pos", depth_pos)):
            require_cuda_f32(t, name)
        self._next_job = (self._depth4(depth), self._depth4(depth_pos))

    def prefetch_sampling(self, depth, depth_pos, grid_hw, stream=None):
        """The sampling of the batch whose ``forward`` comes later in this training step, launched NOW on a side stream
        (dg_loss_presample) — e.g. at the top of ``training_step``, so that it runs under the backbone's forward pass.
        ``grid_hw`` = (H, W) of the feature maps the loss will see.  The forward that receives these depth tensors waits
        for the side stream and skips its own FPS launch."""
        cfg = self.cfg
        for name, t in (("depth", depth), ("depth_pos", depth_pos)):
            require_cuda_f32(t, name)
        depth, depth_pos = self._depth4(depth), self._depth4(depth_pos)
        if depth.dim() != 4 or depth.shape[1] != 1 or depth_pos.shape != depth.shape:
            raise ValueError("depth / depth_pos must be [B,1,Hd,Wd] (or [B,Hd,Wd]) of one shape")
        dev = depth.device
        B, _, Hd, Wd = depth.shape
        H, W = int(grid_hw[0]), int(grid_hw[1])
        S, nneg = int(cfg.feature_samples), int(cfg.neg_samples)
        fused = nneg > 0 and self.negative_sampler == "fused" and self.perm_fn is super_perm
        desc = _lib.LossDesc(B, 32, 32, H, W, Hd, Wd, S, nneg, 0, 0.0, 0.0, 0.0, 0.0)
        plan = _lib.LossPlan()
        check(_lib.lib().dg_loss_plan(C.byref(desc), C.byref(plan)), "dg_loss_plan")
        side = stream if stream is not None else _side_stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))      # the depth maps are ready on the caller's stream
        seed, offset = _reserve_perm_stream(B, dev) if fused else (0, 0)
        with torch.cuda.device(dev), torch.cuda.stream(side):
            coords = torch.empty((2, B, S, S, 2), device=dev, dtype=torch.float32)
            dsign = torch.empty((B, plan.Prows), device=dev, dtype=torch.float32)
            perms = torch.empty((nneg, B), device=dev, dtype=torch.long) if fused else None
            check(_lib.lib().dg_loss_presample(C.byref(desc), ptr(depth), ptr(depth_pos), ptr(coords), ptr(dsign),
                                               ptr(perms), nneg if fused else 0, seed, offset, side.cuda_stream),
                  "dg_loss_presample")
            ev = torch.cuda.Event()
            ev.record(side)
        self._presampled = dict(key=self._sampling_key(depth, depth_pos, S, H, W, nneg), coords=coords, dsign=dsign,
                                perms=perms, event=ev, keep=(depth, depth_pos))

    def _flags(self):
        cfg = self.cfg
        return ((_lib.FLAG_POINTWISE if cfg.pointwise else 0) | (_lib.FLAG_ZERO_CLAMP if cfg.zero_clamp else 0) |
                (_lib.FLAG_STABALIZE if cfg.stabalize else 0))

    def forward(self, orig_feats, orig_feats_pos, orig_salience, orig_salience_pos, orig_code, orig_code_pos,
                depth=None, depth_pos=None):
        return self._run(orig_feats, orig_feats_pos, orig_code, orig_code_pos, depth, depth_pos, None,
                         salience=(orig_salience, orig_salience_pos))

    def _run(self, orig_feats, orig_feats_pos, orig_code, orig_code_pos, depth, depth_pos, aug_feats,
             depth_sampling=None, depth_term=None, salience=(None, None)):
        """``depth_sampling`` / ``depth_term`` override the cfg keys of the same meaning (the depth-only-intra variant
        fixes them whatever cfg says); cfg itself is only ever read."""
        cfg = self.cfg
        depth_sampling = cfg.depth_sampling if depth_sampling is None else depth_sampling
        depth_term = bool(cfg.depth_feat_correlation_loss) if depth_term is None else bool(depth_term)
        self._last_coords = None   # do not keep the previous call's arena alive across this call's allocation
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
        dev = orig_feats.device
        flags = self._flags()
        coords = None
        perms_event = None
        if aug_feats is not None:
            require_cuda_f32(aug_feats, "depth_aug_feats")
            if aug_feats.shape != orig_feats.shape:
                raise ValueError("depth_aug_feats must have the shape of orig_feats")
            flags |= _lib.FLAG_AUG_INTRA
        if cfg.use_salience:
            # src/modules.py:1291-1298 (takes precedence over depth_sampling, as in the reference's if/elif chain):
            # 90 % of the points from the non-zero salience pixels, 10 % uniform
            sal, sal_pos = salience
            if sal is None or sal_pos is None:
                raise ValueError("use_salience=True needs orig_salience and orig_salience_pos")
            shape = [B, S, S, 2]
            c1n = sample_nonzero_locations(sal, shape, self.randint_fn)
            c2n = sample_nonzero_locations(sal_pos, shape, self.randint_fn)
            c1r = self.rand_fn(shape, dev) * 2 - 1
            c2r = self.rand_fn(shape, dev) * 2 - 1
            mask = (self.rand_fn(shape[:-1], dev) > .1).unsqueeze(-1).to(torch.float32)
            coords = torch.stack([c1n * mask + c1r * (1 - mask), c2n * mask + c2r * (1 - mask)]).float().contiguous()
        elif depth_sampling in ("fps", "fps_depth_feat") and aug_feats is None:
            # "fps_depth_feat" = the same call with include_feats=True (:1313-1317), an argument
            # farthest_point_sampling_depth never reads (:999-1037): identical coordinates
            if depth is None or depth_pos is None:
                raise ValueError(f"depth_sampling={depth_sampling!r} needs depth and depth_pos")
            flags |= _lib.FLAG_FPS
        elif depth_sampling == "simple" and aug_feats is None:
            raise NotImplementedError("depth_sampling='simple' (simple_depth_informed_sampling, src/modules.py:828-883: "
                                      "S instead of S*S points per image) is outside the accelerated path")
        else:
            shape = [B, S, S, 2]
            c1 = self.rand_fn(shape, dev) * 2 - 1
            c2 = self.rand_fn(shape, dev) * 2 - 1
            coords = torch.stack([c1, c2]).float().contiguous()
        depth_term = depth_term and aug_feats is None
        Hd = Wd = 0
        if depth_term or (flags & _lib.FLAG_FPS):
            if depth is None:
                raise ValueError("depth_feat_correlation_loss=True needs depth")
            if depth.dim() == 3:        # [B,Hd,Wd]: the Potsdam dataset yields depth without a channel axis
                depth = depth.unsqueeze(1)   # (src/data.py:226); adaptive_avg_pool2d treats it the same way
            if depth_pos is not None and depth_pos.dim() == 3:
                depth_pos = depth_pos.unsqueeze(1)
            for name, t in (("depth", depth), ("depth_pos", depth_pos)):
                if t is not None:
                    require_cuda_f32(t, name)
                    if t.dim() != 4 or t.shape[1] != 1 or t.shape[0] != B:
                        raise ValueError(f"{name} must be [B,1,Hd,Wd] (or [B,Hd,Wd]), got {tuple(t.shape)}")
            if depth_pos is not None and depth_pos.shape != depth.shape:
                raise ValueError(f"depth_pos {tuple(depth_pos.shape)} must have the shape of depth {tuple(depth.shape)}")
            depth = depth.contiguous()
            depth_pos = depth_pos.contiguous() if depth_pos is not None else None
            Hd, Wd = depth.shape[-2:]
            if depth_term:
                flags |= _lib.FLAG_DEPTH_TERM
        # Was this batch sampled ahead of time (queue_next_sampling of the previous call / prefetch_sampling)?  Then the
        # forward takes coordinates, depth signs and permutations from there and launches no FPS.
        pre, self._presampled = self._presampled, None
        dsign_in = None
        use_pre = False
        self.last_used_presampled = False      # test hook
        if pre is not None and (flags & _lib.FLAG_FPS) and depth_pos is not None and compiled_binding() and \
                not _CorrLossFn.debug and not torch.cuda.is_current_stream_capturing() and \
                pre["key"] == self._sampling_key(depth, depth_pos, S, H, W, nneg):
            use_pre = self.last_used_presampled = True
            flags &= ~_lib.FLAG_FPS
            coords = pre["coords"]
            dsign_in = pre["dsign"] if depth_term else None
            if pre["event"] is not None:     # produced on a side stream (prefetch_sampling)
                cur = torch.cuda.current_stream(dev)
                cur.wait_event(pre["event"])
                for t in (pre["coords"], pre["dsign"], pre["perms"]):
                    if t is not None:        # allocated on the side stream, consumed here: keep the allocator from
                        t.record_stream(cur)  # handing the block out again before this stream is done with it
        also_wait, self.wait_after_fps = self.wait_after_fps, None
        perm_gen = None
        self.last_perms = None
        if not nneg:
            perms = None
        elif use_pre and pre["perms"] is not None and self.perm_fn is super_perm and self.negative_sampler == "fused":
            perms = pre["perms"]        # drawn together with the coordinates
        elif self.perm_fn is not super_perm:
            perms = torch.stack([self.perm_fn(B, dev) for _ in range(nneg)]).to(torch.long).contiguous()
        elif self.negative_sampler == "fused" and not torch.cuda.is_current_stream_capturing():
            # drawn by the forward itself (one extra CTA of its FPS launch): only the buffer and the stream position
            # are fixed here
            perms = torch.empty((nneg, B), device=dev, dtype=torch.long)
            perm_gen = _reserve_perm_stream(B, dev)
        elif self.negative_sampler == "fused":
            # under CUDA-graph capture the generator's offset cannot be read on the host: torch.randperm is graph-safe
            perms = super_perms(nneg, B, dev)
        elif self.graph_negative_sampler:
            perms, perms_event = _GraphedSuperPerms.draw_async(nneg, B, dev, also_wait)
        else:
            perms = super_perms(nneg, B, dev)
        if also_wait is not None and perms_event is None:
            # no sampler stream to fold it into: the library waits for this event itself, right after launching FPS
            # (dg_loss_io_t.perms_ready is exactly that wait)
            perms_event = also_wait

        self.last_perms = perms     # [neg_samples,B] source image of every negative (test hook; filled by the forward
        #                             itself with the default sampler)
        if corr_kernel_choice(S * S, orig_code.shape[1]) == "simt":
            flags |= _lib.FLAG_FORCE_SIMT
        elif Cdim % 128 == 0 and orig_feats.is_contiguous() and orig_feats_pos.is_contiguous() and \
                (aug_feats is None or aug_feats.is_contiguous()):
            flags |= _lib.FLAG_STAGE_NHWC   # NCHW inputs: let the library stage channels-last copies for the gather
        desc = _lib.LossDesc(B, Cdim, orig_code.shape[1], H, W, Hd, Wd, S, nneg, flags, float(cfg.pos_intra_shift),
                             float(cfg.pos_inter_shift), float(cfg.neg_inter_shift),
                             float(cfg.depth_feat_shift) if depth_term else 0.0)
        for name, t in (("orig_feats_pos", orig_feats_pos), ("orig_code", orig_code), ("orig_code_pos", orig_code_pos),
                        ("depth", depth), ("depth_pos", depth_pos), ("depth_aug_feats", aug_feats)):
            if t is not None and t.device != dev:
                raise ValueError(f"{name} is on {t.device}, orig_feats on {dev}")
        binding = None if _CorrLossFn.debug else compiled_binding()
        if binding:
            ints = (B, Cdim, orig_code.shape[1], H, W, Hd, Wd, S, nneg, flags)
            shifts = (desc.pos_intra_shift, desc.pos_inter_shift, desc.neg_inter_shift, desc.depth_feat_shift)
            coords_off = _plan_cache.get(ints)
            if coords_off is None:
                coords_off = _plan_cache[ints] = binding.loss_plan(list(ints))[1]
            if perm_gen is not None:     # (seed as a signed 64-bit value: the binding casts it back)
                seed = perm_gen[0] - (1 << 64) if perm_gen[0] >= (1 << 63) else perm_gen[0]
                ints = ints + (seed, perm_gen[1])
            # the next batch's sampling rides in this forward (queue_next_sampling)
            nxt, self._next_job = self._next_job, None
            n_coords = n_dsign = n_perms = None
            n_rng = []
            fps_mode = use_pre or bool(flags & _lib.FLAG_FPS)
            if nxt is not None and fps_mode and aug_feats is None and nxt[0].shape == depth.shape and \
                    nxt[1].shape == depth.shape and not torch.cuda.is_current_stream_capturing():
                Prows = _prows_cache.get(ints[:10])
                if Prows is None:
                    Prows = _prows_cache[ints[:10]] = binding.loss_plan(list(ints[:10]))[5]
                n_coords = torch.empty((2, B, S, S, 2), device=dev, dtype=torch.float32)
                n_dsign = torch.empty((B, Prows), device=dev, dtype=torch.float32)
                if nneg and self.negative_sampler == "fused" and self.perm_fn is super_perm:
                    n_perms = torch.empty((nneg, B), device=dev, dtype=torch.long)
                    sd, off = _reserve_perm_stream(B, dev)
                    n_rng = [sd - (1 << 64) if sd >= (1 << 63) else sd, off]
            else:
                nxt = None
            res = binding.corr_loss(orig_feats, orig_feats_pos, orig_code, orig_code_pos, depth, depth_pos, coords,
                                    perms, aug_feats, ints, shifts, bool(self.materialize_cd), False,
                                    torch.is_grad_enabled(),   # gradient buffers zeroed inside the forward's launches
                                    perms_event.cuda_event if perms_event is not None else 0,
                                    dsign_in, nxt[0] if nxt else None, nxt[1] if nxt else None, n_coords, n_dsign,
                                    n_perms, n_rng)
            if nxt is not None:     # same stream as the forward that will use it: no event needed
                self._presampled = dict(key=self._sampling_key(nxt[0], nxt[1], S, H, W, nneg), coords=n_coords,
                                        dsign=n_dsign, perms=n_perms, event=None, keep=nxt)
            intra, inter, neg, dloss, out8, coords_src = res[:6]
            cd_out, loss_out = (res[6], res[7]) if self.materialize_cd else (None, None)
            dd_out = res[8] if (self.materialize_cd and depth_term) else None
        else:
            with torch.cuda.device(dev):    # kernels, attribute set-up and the stream all belong to the tensors' device
                res = _CorrLossFn.apply(orig_feats, orig_feats_pos, orig_code, orig_code_pos, depth, depth_pos, coords,
                                        perms, desc, bool(self.materialize_cd), perms_event, aug_feats, perm_gen)
            intra, inter, neg, dloss, out8, coords_src, cd_out, loss_out, dd_out = res
            coords_off = _CorrLossFn.last_coords_off
        # FPS coordinates stay in the arena until someone asks for them (see the last_coords property)
        self._last_coords = (coords_src, coords_off, B, S) if (flags & _lib.FLAG_FPS) else coords.view(2, B, S, S, 2)
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
            v = out8.unbind(0)
            intra_cd, inter_cd, neg_cd, dd = v[1], v[3], v[5], v[7]
            neg_loss = neg
        head = (intra, intra_cd, inter, inter_cd, neg_loss, neg_cd)
        if depth_term:
            return head + (dloss, dd)
        return head


class DepthContrastiveCorrelationLoss(ContrastiveCorrelationLoss):
    """Drop-in for src/modules.py:1370-1463 (used when ``use_depth_only_intra``, src/train_segmentation.py:133-134):
    the same helper, but the intra pair correlates depth-augmented features with themselves, coordinates are always
    random, there is no depth term, and the output is the 6-tuple.  Runs the same kernels (the intra pair reads one
    extra feature panel slot).  ``depth_aug_feats_pos`` is sampled and never used by the reference; it is ignored."""

    def forward(self, orig_feats, orig_feats_pos, orig_salience, orig_salience_pos, orig_code, orig_code_pos,
                depth_aug_feats, depth_aug_feats_pos=None):
        if depth_aug_feats is None:
            raise ValueError("DepthContrastiveCorrelationLoss needs depth_aug_feats")
        # this variant never uses depth-guided sampling or the depth term, whatever cfg says (:1413-1425)
        return self._run(orig_feats, orig_feats_pos, orig_code, orig_code_pos, None, None, depth_aug_feats,
                         depth_sampling="none", depth_term=False, salience=(orig_salience, orig_salience_pos))
