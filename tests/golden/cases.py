"""Seeded input generators shared by make_golden.py (run once, in the build
container, against the real reference) and by the parity tests (run anywhere).

Everything comes from ``numpy.random.RandomState`` whose streams are frozen by
NumPy's compatibility guarantee, so the inputs regenerate identically on the
GPU box and need not be stored.
"""
from types import SimpleNamespace

import numpy as np
import torch


def loss_cfg(**over):
    """The loss-relevant keys of src/configs/local_config.yml (reference defaults)."""
    cfg = dict(feature_samples=11, use_salience=False, depth_sampling="fps", fps_gpu=False, pointwise=True,
               zero_clamp=True, stabalize=False, neg_samples=5, pos_intra_shift=0.18, pos_inter_shift=0.12,
               neg_inter_shift=0.46, depth_feat_correlation_loss=True, depth_feat_shift=0.03)
    cfg.update(over)
    return SimpleNamespace(**cfg)


# name -> (B, C, D, cfg overrides, seed)
LOSS_CASES = {
    "small_fps": (3, 32, 16, dict(feature_samples=5), 11),
    "small_fps_nopointwise": (3, 32, 16, dict(feature_samples=4, pointwise=False), 12),
    "small_fps_noclamp": (2, 24, 10, dict(feature_samples=6, zero_clamp=False), 13),
    "small_fps_stabalize": (2, 24, 10, dict(feature_samples=6, stabalize=True, pos_intra_shift=0.02), 14),
    "small_fps_nodepthterm": (4, 40, 12, dict(feature_samples=7, depth_feat_correlation_loss=False), 15),
    "small_random": (4, 32, 20, dict(feature_samples=8, depth_sampling="none", pointwise=False), 16),
    "small_random_pointwise": (2, 64, 70, dict(feature_samples=11, depth_sampling="none", neg_samples=3), 17),
    "small_zero_depth": (2, 16, 8, dict(feature_samples=3), 18),
    "s12_fps": (2, 48, 90, dict(feature_samples=12), 19),
    "cfg1_vits": (2, 384, 70, dict(feature_samples=11, pos_intra_shift=0.08, pos_inter_shift=0.02,
                                   neg_inter_shift=0.66), 0),
    # Cityscapes five-crop flavour (BASELINE configs[3]): dim 100, random coordinates, pointwise off
    "cfg4_cityscapes": (4, 128, 100, dict(feature_samples=11, depth_sampling="none", pointwise=False,
                                          pos_intra_shift=0.18, pos_inter_shift=0.12, neg_inter_shift=0.46,
                                          depth_feat_shift=0.01), 21),
    # dense, unsampled correlation (BASELINE configs[4]) on a small grid: every grid point is a sample
    "dense_14x14": (2, 48, 24, dict(feature_samples=14), 22),
    # the full 28x28 stress shape (784 x 784 per pair): FPS then selects every grid point; tcgen05 column-group mode
    "dense_28x28": (2, 32, 16, dict(feature_samples=28, neg_samples=2), 23),
    # 20 x 20 = 400 random points: four row tiles x two column groups, pointwise off
    "dense_400_random": (2, 64, 40, dict(feature_samples=20, depth_sampling="none", pointwise=False, neg_samples=2), 24),
}

# DepthContrastiveCorrelationLoss (src/modules.py:1370-1463) cases: name -> base loss case + seed of the augmented features
AUG_CASES = {
    "aug_small": ("small_random", 31),
    "aug_vits_pointwise": ("small_random_pointwise", 32),
}


def make_aug_inputs(name):
    base, seed = AUG_CASES[name]
    cfg, t = make_loss_inputs(base)
    rs = np.random.RandomState(3000 + seed)
    B, C, H, W = t["feats"].shape
    t["aug"] = torch.from_numpy(correlated(rs, B, C, H, W))
    t["aug_pos"] = torch.from_numpy(correlated(rs, B, C, H, W))
    return cfg, t


# use_salience sampling (src/modules.py:1291-1298, :1191-1204): name -> (base loss case, seed).  The salience maps, the
# uniform numbers behind every torch.randint draw and the 90/10 mask draw are case data, so the real reference (golden),
# the oracle and the CUDA path all see the same draws.
SAL_CASES = {
    "salience_small": ("small_fps", 41),             # depth term on; use_salience overrides depth_sampling='fps'
    "salience_random_pointwise": ("small_random_pointwise", 42),
}


def make_sal_inputs(name):
    base, seed = SAL_CASES[name]
    cfg, t = make_loss_inputs(base)
    cfg.use_salience = True
    rs = np.random.RandomState(4000 + seed)
    B, _, H, W = t["feats"].shape
    S = cfg.feature_samples
    sal = (rs.random_sample((B, H, W)) > 0.7).astype(np.float32) * rs.uniform(0.2, 1.0, (B, H, W)).astype(np.float32)
    sal_pos = (rs.random_sample((B, H, W)) > 0.5).astype(np.float32)
    sal[B - 1] = 0.0                                  # an all-zero map: the uniform-pixel fallback (:1196-1197)
    t["salience"], t["salience_pos"] = torch.from_numpy(sal), torch.from_numpy(sal_pos)
    # one [n,2] block of uniforms per randint call: 2 maps x B images, in the reference's call order
    t["randint_u"] = torch.from_numpy(rs.random_sample((2 * B, S * S, 2)))
    t["rand_mask"] = torch.from_numpy(rs.random_sample((B, S, S)).astype(np.float32))
    return cfg, t


def randint_from_uniforms(blocks):
    """A stand-in for torch.randint(high, size=..., [device=...]) that turns the next block of uniforms into
    floor(u * high) with the requested shape ((n,) uses the first column, (n,2) both)."""
    it = iter(blocks)

    def randint(high, size, device=None):
        u = next(it)
        r = torch.floor(u * high).to(torch.int64).clamp_(max=high - 1)
        r = r[:, 0] if len(size) == 1 else r
        assert tuple(r.shape) == tuple(size)
        return r if device is None else r.to(device)

    return randint


# backprop weights for the scalar L = sum w_i * loss_i  (ViT-B paper run, paper_reproduction.sh:8)
LOSS_WEIGHTS = dict(pos_inter=1.0501, pos_intra=0.2305, neg_inter=0.2485, depth_feat=0.1603)


def correlated(rs, B, C, H, W, rank=6):
    """Feature maps with spatial structure (low-rank field + noise) so correlations are not ~0."""
    basis = rs.standard_normal((B, rank, H, W)).astype(np.float32)
    mix = rs.standard_normal((C, rank)).astype(np.float32)
    x = np.einsum("cr,brhw->bchw", mix, basis) + 0.5 * rs.standard_normal((B, C, H, W)).astype(np.float32)
    return np.ascontiguousarray(x.astype(np.float32))


def make_loss_inputs(name, H=28, W=28, Hd=224, Wd=224):
    B, C, D, over, seed = LOSS_CASES[name]
    if name.startswith("dense_"):
        H = W = over["feature_samples"]
        Hd = Wd = 8 * H
    rs = np.random.RandomState(1000 + seed)
    feats = correlated(rs, B, C, H, W)
    feats_pos = correlated(rs, B, C, H, W)
    code = correlated(rs, B, D, H, W, rank=4)
    code_pos = correlated(rs, B, D, H, W, rank=4)
    depth = rs.randint(0, 256, (B, 1, Hd, Wd)).astype(np.float32)
    depth_pos = smooth_depth(rs, B, Hd, Wd)
    if name == "small_zero_depth":
        depth[:] = 0.0
        depth_pos[0, :, :, : Wd // 2] = 0.0
    cfg = loss_cfg(**over)
    S = cfg.feature_samples
    perms = np.stack([bumped_perm(rs, B) for _ in range(cfg.neg_samples)]).astype(np.int64)
    rand1 = rs.random_sample((B, S, S, 2)).astype(np.float32)
    rand2 = rs.random_sample((B, S, S, 2)).astype(np.float32)
    t = {k: torch.from_numpy(v) for k, v in dict(feats=feats, feats_pos=feats_pos, code=code, code_pos=code_pos,
                                                  depth=depth, depth_pos=depth_pos, perms=perms, rand1=rand1,
                                                  rand2=rand2).items()}
    return cfg, t


def bumped_perm(rs, B):
    """super_perm's arithmetic on a RandomState permutation (src/modules.py:1184-1188)."""
    p = rs.permutation(B)
    p = np.where(p == np.arange(B), p + 1, p)
    return p % B


def smooth_depth(rs, B, Hd, Wd):
    """uint8-valued smooth depth like the ZoeDepth PNGs (generate_depth.py:232-240)."""
    yy, xx = np.mgrid[0:Hd, 0:Wd].astype(np.float32)
    out = np.zeros((B, 1, Hd, Wd), np.float32)
    for b in range(B):
        a, c, e = rs.uniform(0.2, 1.0, 3)
        f = 127.5 + 90 * np.sin(a * yy / 23.0 + e) * np.cos(c * xx / 31.0) + 30 * (yy / Hd)
        out[b, 0] = np.clip(np.round(f), 0, 255)
    return out


# FPS depth patterns: name -> generator(rs, B, Hd, Wd) -> float32 [B,1,Hd,Wd]
def _uint8_random(rs, B, Hd, Wd):
    return rs.randint(0, 256, (B, 1, Hd, Wd)).astype(np.float32)


def _piecewise(rs, B, Hd, Wd):
    blocks = rs.randint(0, 4, (B, 1, Hd // 56, Wd // 56)).astype(np.float32) * 60.0
    return np.kron(blocks, np.ones((1, 1, 56, 56), np.float32))


def _constant(rs, B, Hd, Wd):
    return np.full((B, 1, Hd, Wd), 255.0, np.float32)


def _zero(rs, B, Hd, Wd):
    return np.zeros((B, 1, Hd, Wd), np.float32)


def _float_arbitrary(rs, B, Hd, Wd):
    return (rs.random_sample((B, 1, Hd, Wd)) * 10.0).astype(np.float32)


def _half_zero(rs, B, Hd, Wd):
    d = rs.randint(0, 256, (B, 1, Hd, Wd)).astype(np.float32)
    d[:, :, : Hd // 2] = 0.0
    return d


FPS_PATTERNS = {
    "uint8_random": _uint8_random,
    "uint8_smooth": smooth_depth,
    "piecewise": _piecewise,
    "constant255": _constant,
    "zero": _zero,
    "float_arbitrary": _float_arbitrary,
    "half_zero": _half_zero,
}
FPS_S = (3, 5, 11, 12)
FPS_B = 3


def make_fps_depth(pattern, Hd=224, Wd=224, B=FPS_B):
    rs = np.random.RandomState(abs(hash_name(pattern)) % (2 ** 31))
    return torch.from_numpy(np.ascontiguousarray(FPS_PATTERNS[pattern](rs, B, Hd, Wd)))


def hash_name(s):
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


KNN_CASES = {  # name -> (N, F, k, n_batches, clustered, seed)
    "iid_2048": (2048, 64, 30, 64, False, 5),
    "clustered_3000": (3000, 96, 30, 64, True, 6),
    "ragged_1001": (1001, 48, 8, 64, False, 7),
}


def make_knn_feats(name):
    N, Fdim, k, n_batches, clustered, seed = KNN_CASES[name]
    rs = np.random.RandomState(2000 + seed)
    if clustered:
        cent = rs.standard_normal((50, Fdim)).astype(np.float32)
        x = cent[rs.randint(0, 50, N)] + 0.3 * rs.standard_normal((N, Fdim)).astype(np.float32)
    else:
        x = rs.standard_normal((N, Fdim)).astype(np.float32)
    x = torch.from_numpy(x)
    return torch.nn.functional.normalize(x, dim=1), k, n_batches


# ------------------------------------------------------------------ probe losses
PROBE_CASES = {  # name -> (B, D, h, w, K, Hl, Wl, frac_ignored, seed)
    "probe_small": (2, 16, 7, 7, 5, 56, 56, 0.1, 1),
    "probe_odd": (3, 70, 14, 10, 27, 100, 77, 0.2, 2),      # non-integer scale factors, non-square
    "probe_cfg2": (4, 90, 28, 28, 27, 224, 224, 0.05, 3),   # cocostuff27 ViT-B/8 shapes (B reduced)
    "probe_same": (2, 12, 9, 9, 3, 9, 9, 0.0, 4),           # label at code resolution (scale 1)
}


def make_probe_inputs(name):
    B, D, h, w, K, Hl, Wl, ignored, seed = PROBE_CASES[name]
    rs = np.random.RandomState(3000 + seed)
    code = correlated(rs, B, D, h, w, rank=4)
    weight = (rs.standard_normal((K, D)) * (3.0 / np.sqrt(D))).astype(np.float32)
    bias = (rs.standard_normal((K,)) * 0.3).astype(np.float32)
    clusters = rs.standard_normal((K, D)).astype(np.float32)
    # blocky label map with an ignore class (-1) and an out-of-range id (K), as the datasets produce
    coarse = rs.randint(0, K, (B, (Hl + 7) // 8, (Wl + 7) // 8))
    label = np.repeat(np.repeat(coarse, 8, axis=1), 8, axis=2)[:, :Hl, :Wl].astype(np.int64)
    drop = rs.random_sample(label.shape)
    label[drop < ignored / 2] = -1
    label[(drop >= ignored / 2) & (drop < ignored)] = K
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in
            dict(code=code, weight=weight, bias=bias, clusters=clusters, label=label).items()}
