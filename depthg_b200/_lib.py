"""ctypes binding of libdepthg_b200.so (the C ABI declared in include/depthg_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``make -C
depthg_b200/csrc``).  There is no CPU or PyTorch fallback: if the shared
object is missing, or a tensor is not a CUDA fp32 tensor, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdepthg_b200.so")

DG_MAX_PAIRS = 32
DG_MAX_SETS = 32
GROUP_INTRA, GROUP_INTER, GROUP_NEG, GROUP_DEPTH = 0, 1, 2, 3
FLAG_POINTWISE, FLAG_ZERO_CLAMP, FLAG_STABALIZE = 1, 2, 4

_c_i32p = C.POINTER(C.c_int32)
_c_i64p = C.POINTER(C.c_int64)
_c_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/depthg_b200.h declares
SIGNATURES = {
    "dg_version": (C.c_int, []),
    "dg_last_error_string": (C.c_char_p, []),
    "dg_kernel_launches": (C.c_ulonglong, []),
    "dg_profile_enable": (C.c_int, [C.c_int]),
    "dg_debug_set_clock_buffer": (C.c_int, [_vp]),
    "dg_profile_collect": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "dg_panel_ld": (C.c_int, [C.c_int]),
    "dg_panel_rows": (C.c_int, [C.c_int]),
    "dg_fps_coords": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                C.c_int, _vp, _vp, _vp]),
    "dg_super_perms": (C.c_int, [C.c_ulonglong, C.c_ulonglong, C.c_int, C.c_int, _vp, _vp]),
    "dg_depth_sign": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _vp, _vp]),
    "dg_gather_norm": (C.c_int, [_vp, _c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _c_i32p,
                                 _c_i32p, _vp, C.c_float, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dg_norm_dim1": (C.c_int, [_vp, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong,
                               C.c_float, _vp, _vp]),
    "dg_gather_norm_bwd": (C.c_int, [_vp, _c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _c_i32p,
                                     _c_i32p, _vp, C.c_float, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int,
                                     _c_i32p, _c_f32p, C.c_int, _vp, _vp]),
    "dg_corr_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dg_corr_loss": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, _c_f32p, _c_i32p, C.c_float, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                               C.c_size_t, _vp]),
    "dg_loss_plan": (C.c_int, [_vp, _vp]),
    "dg_loss_forward": (C.c_int, [_vp, _vp, _vp]),
    "dg_loss_presample": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_ulonglong, C.c_ulonglong, _vp]),
    "dg_loss_backward": (C.c_int, [_vp, _vp, _vp, _vp]),
    "dg_knn_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dg_knn_topk": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dg_knn_shard_begin": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, C.c_int, _vp]),
    "dg_knn_shard_finish": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_size_t,
                                      C.c_int, C.c_int, _vp]),
    "dg_knn_panel_count": (C.c_int, []),
    "dg_knn_panel_layout": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_size_t)]),
    "dg_memcpy_batch": (C.c_int, [C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_size_t), C.POINTER(_vp)]),
    "dg_pool_normalize": (C.c_int, [_vp, _c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _vp, _vp]),
    "dg_probe_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dg_linear_probe_ce": (C.c_int, [_vp, _c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _c_i64p,
                                     C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dg_cluster_probe": (C.c_int, [_vp, _c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_float,
                                   _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
}

PANEL_F32, PANEL_FEATS_SPLIT, PANEL_CODE_SPLIT = 0, 1, 2
KNN_SPLIT_LOCAL, KNN_PASS_LOCAL, KNN_BEGIN_ALL, KNN_ALL_SMS = 1, 2, 3, 8
KNN_SPLIT_REMOTE, KNN_PASS_REMOTE, KNN_RERANK, KNN_FINISH_ALL = 1, 2, 4, 7


class Panels(C.Structure):
    """dg_panels_t of include/depthg_b200.h."""
    _fields_ = [("format", C.c_int), ("f_hi", _vp), ("f_lo", _vp), ("c_hi", _vp), ("c_lo", _vp), ("cb_hi", _vp),
                ("cb_lo", _vp)]


def make_panels(fmt, f_hi, f_lo, c_hi, c_lo, cb_hi, cb_lo):
    dp = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    return Panels(fmt, dp(f_hi), dp(f_lo), dp(c_hi), dp(c_lo), dp(cb_hi), dp(cb_lo))


FLAG_DEPTH_TERM, FLAG_FPS, FLAG_FORCE_SIMT, FLAG_STAGE_NHWC, FLAG_AUG_INTRA = 8, 16, 32, 64, 128
_i64x4 = C.c_int64 * 4


class LossDesc(C.Structure):
    """dg_loss_desc_t"""
    _fields_ = [(n, C.c_int) for n in ("B", "C", "D", "H", "W", "Hd", "Wd", "S", "neg_samples", "flags")] + \
               [(n, C.c_float) for n in ("pos_intra_shift", "pos_inter_shift", "neg_inter_shift", "depth_feat_shift")]


class LossPlan(C.Structure):
    """dg_loss_plan_t"""
    _fields_ = [(n, C.c_size_t) for n in ("total", "coords", "frn", "fmean", "crn", "f_hi", "f_lo", "c_hi", "c_lo",
                                          "cb_hi", "cb_lo", "dsign", "dC1", "dC2", "ws", "ws_bytes", "stage")] + \
               [(n, C.c_int) for n in ("kernel", "Prows", "ldf", "ldc", "npairs")]


class LossIO(C.Structure):
    """dg_loss_io_t"""
    _fields_ = [("feats", _vp), ("feats_pos", _vp), ("code", _vp), ("code_pos", _vp),
                ("feats_strides", _i64x4), ("feats_pos_strides", _i64x4), ("code_strides", _i64x4),
                ("code_pos_strides", _i64x4), ("depth", _vp), ("depth_pos", _vp), ("coords", _vp), ("perms", _vp),
                ("arena", _vp), ("out8", _vp), ("cd_out", _vp), ("loss_out", _vp), ("dd_out", _vp), ("fd_dbg", _vp), ("aug_feats", _vp), ("aug_feats_strides", _i64x4),
                ("perms_ready", _vp), ("perm_seed", C.c_ulonglong), ("perm_offset", C.c_ulonglong),
                ("gen_perms", C.c_int), ("clear", _vp * 2), ("clear_bytes", C.c_size_t * 2),
                ("dsign", _vp), ("next_depth", _vp), ("next_depth_pos", _vp), ("next_coords", _vp), ("next_dsign", _vp),
                ("next_perms", _vp), ("next_perm_seed", C.c_ulonglong), ("next_perm_offset", C.c_ulonglong),
                ("next_n_perms", C.c_int)]


class LossGrads(C.Structure):
    """dg_loss_grads_t"""
    _fields_ = [("g", _vp * 4), ("d_code", _vp), ("d_code_pos", _vp), ("d_code_strides", _i64x4),
                ("d_code_pos_strides", _i64x4)]


_lib = None


class DepthgB200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise loudly if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise DepthgB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C depthg_b200/csrc`). depthg_b200 has no CPU / PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().dg_last_error_string().decode(errors="replace")
        if rc in (-1, -2):
            raise ValueError(f"{what}: {msg} (code {rc})")
        raise DepthgB200Error(f"{what}: {msg} (code {rc})")


def require_cuda_f32(t: torch.Tensor, name: str):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: depthg_b200 has no CPU path (got device {t.device})")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32 (the reference computes in fp32), got {t.dtype}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device_index=None):
    """Raw handle of torch's current CUDA stream (the fast C accessor: this is on the per-step path)."""
    if device_index is None:
        device_index = torch.cuda.current_device()
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(device_index))


def i32_array(vals):
    return (C.c_int32 * len(vals))(*[int(v) for v in vals])


def i64_array(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def f32_array(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def panel_ld(channels: int) -> int:
    return (channels + 31) // 32 * 32


def panel_rows(P: int) -> int:
    return (P + 63) // 64 * 64
