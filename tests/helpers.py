"""Shared test plumbing: golden loading and running the oracle on a case."""
import os

import numpy as np
import torch

from oracle import depthg_oracle as O
from tests.golden import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def weighted_total(out, depth_term):
    w = cases.LOSS_WEIGHTS
    L = w["pos_intra"] * out[0] + w["pos_inter"] * out[2] + w["neg_inter"] * out[4].mean()
    if depth_term:
        L = L + w["depth_feat"] * out[6]
    return L


def run_oracle_loss(name, dtype=torch.float32):
    """Oracle forward+backward on a golden case with the case's perms / random coords injected."""
    cfg, t = cases.make_loss_inputs(name)
    if dtype != torch.float32:  # higher-precision evaluation of the same algorithm (noise-floor checks)
        t = {k: (v.to(dtype) if v.is_floating_point() and k not in ("rand1", "rand2") else v) for k, v in t.items()}
    code = t["code"].clone().requires_grad_(True)
    code_pos = t["code_pos"].clone().requires_grad_(True)
    fn = O.ContrastiveCorrelationLoss(cfg)
    perm_it, rand_it = iter(t["perms"]), iter([t["rand1"], t["rand2"]])
    fn.perm_fn = lambda B, device: next(perm_it).clone()
    fn.rand_fn = lambda shape, device: next(rand_it).clone()
    out = fn(t["feats"], t["feats_pos"], None, None, code, code_pos, t["depth"], t["depth_pos"])
    depth_term = cfg.depth_feat_correlation_loss
    L = weighted_total(out, depth_term)
    L.backward()
    res = dict(coords1=fn.last_coords[0].numpy(), coords2=fn.last_coords[1].numpy(),
               scalars=np.array([out[0].item(), out[2].item(), out[4].mean().item(),
                                 out[6].item() if depth_term else np.nan]),
               cd_means=np.array([out[1].mean().item(), out[3].mean().item(), out[5].mean().item(),
                                  out[7].mean().item() if depth_term else np.nan]),
               total=L.item(), d_code=code.grad.numpy(), d_code_pos=code_pos.grad.numpy(), out=out)
    return cfg, t, res


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _aug_total(out):
    w = cases.LOSS_WEIGHTS
    return w["pos_intra"] * out[0] + w["pos_inter"] * out[2] + w["neg_inter"] * out[4].mean()


def _aug_result(out, code, code_pos):
    L = _aug_total(out)
    L.backward()
    return dict(scalars=np.array([out[0].item(), out[2].item(), out[4].mean().item()]),
                cd_means=np.array([out[1].mean().item(), out[3].mean().item(), out[5].mean().item()]),
                total=L.item(), d_code=code.grad.detach().cpu().numpy(), d_code_pos=code_pos.grad.detach().cpu().numpy())


def run_oracle_aug(name):
    cfg, t = cases.make_aug_inputs(name)
    code = t["code"].clone().requires_grad_(True)
    code_pos = t["code_pos"].clone().requires_grad_(True)
    fn = O.DepthContrastiveCorrelationLoss(cfg)
    perm_it, rand_it = iter(t["perms"]), iter([t["rand1"], t["rand2"]])
    fn.perm_fn = lambda B, device: next(perm_it).clone()
    fn.rand_fn = lambda shape, device: next(rand_it).clone()
    out = fn(t["feats"], t["feats_pos"], None, None, code, code_pos, t["aug"], t["aug_pos"])
    return _aug_result(out, code, code_pos)


def run_cuda_aug(name, channels_last=False, force_simt=False):
    import os
    from depthg_b200.modules import DepthContrastiveCorrelationLoss
    cfg, t = cases.make_aug_inputs(name)
    dev = torch.device("cuda:0")

    def put(x):
        x = x.to(dev)
        return x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2) if channels_last else x

    code = put(t["code"]).detach().requires_grad_(True)
    code_pos = put(t["code_pos"]).detach().requires_grad_(True)
    fn = DepthContrastiveCorrelationLoss(cfg)
    perm_it, rand_it = iter(t["perms"].to(dev)), iter([t["rand1"].to(dev), t["rand2"].to(dev)])
    fn.perm_fn = lambda B, device: next(perm_it).clone()
    fn.rand_fn = lambda shape, device: next(rand_it).clone()
    old = os.environ.get("DEPTHG_B200_CORR")
    if force_simt:
        os.environ["DEPTHG_B200_CORR"] = "simt"
    try:
        out = fn(put(t["feats"]), put(t["feats_pos"]), None, None, code, code_pos, put(t["aug"]), put(t["aug_pos"]))
    finally:
        if force_simt:
            if old is None:
                del os.environ["DEPTHG_B200_CORR"]
            else:
                os.environ["DEPTHG_B200_CORR"] = old
    assert len(out) == 6
    return _aug_result(out, code, code_pos)


def _sal_result(fn, out, code, code_pos, depth_term):
    L = weighted_total(out, depth_term)
    L.backward()
    c1, c2 = fn.last_coords[0], fn.last_coords[1]
    return dict(coords1=c1.detach().cpu().numpy(), coords2=c2.detach().cpu().numpy(),
                scalars=np.array([out[0].item(), out[2].item(), out[4].mean().item(),
                                  out[6].item() if depth_term else np.nan]),
                cd_means=np.array([out[1].mean().item(), out[3].mean().item(), out[5].mean().item(),
                                   out[7].mean().item() if depth_term else np.nan]),
                total=L.item(), d_code=code.grad.detach().cpu().numpy(), d_code_pos=code_pos.grad.detach().cpu().numpy())


def run_sal(name, impl, device="cpu", channels_last=False):
    """use_salience case through `impl` (the oracle's or the product's ContrastiveCorrelationLoss) with the case's
    draws injected: perms, rand x3 (coords1_reg, coords2_reg, mask) and the uniforms behind every randint."""
    cfg, t = cases.make_sal_inputs(name)
    dev = torch.device(device)

    def put(x):
        x = x.to(dev)
        return x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2) if channels_last else x

    code = put(t["code"]).detach().requires_grad_(True)
    code_pos = put(t["code_pos"]).detach().requires_grad_(True)
    fn = impl(cfg)
    perm_it = iter(t["perms"].to(dev))
    rand_it = iter([t["rand1"].to(dev), t["rand2"].to(dev), t["rand_mask"].to(dev)])
    fn.perm_fn = lambda B, device: next(perm_it).clone()
    fn.rand_fn = lambda shape, device: next(rand_it).clone()
    fn.randint_fn = cases.randint_from_uniforms(list(t["randint_u"]))
    out = fn(put(t["feats"]), put(t["feats_pos"]), t["salience"].to(dev), t["salience_pos"].to(dev), code, code_pos,
             t["depth"].to(dev), t["depth_pos"].to(dev))
    return _sal_result(fn, out, code, code_pos, cfg.depth_feat_correlation_loss)


def clamp_tie_images(out64, perms, B, tau=1e-6):
    """Images whose gradient may legitimately differ between two correct evaluations of the loss: the zero clamp is a
    step function of cd, and a code correlation within `tau` of zero can land on either side depending on the fp32
    summation order (the reference's own fp32 run flips such indicators against its fp64 run).  `out64` is the fp64
    oracle's output tuple.  Returns (set of images of d_code, set of images of d_code_pos) touched by such a pair:
    the intra pair of image b touches code[b]; the inter pair code[b] and code_pos[b]; negative n of image b code[b] and
    code[perm_n[b]]."""
    code, code_pos = set(), set()
    near = lambda cd: (cd.detach().abs().flatten(1) < tau).any(1).numpy()   # noqa: E731
    for b in np.nonzero(near(out64[1]))[0]:
        code.add(int(b))
    for b in np.nonzero(near(out64[3]))[0]:
        code.add(int(b))
        code_pos.add(int(b))
    neg = near(out64[5]).reshape(-1, B)
    for n, b in zip(*np.nonzero(neg)):
        code.add(int(b))
        code.add(int(perms[n][b]))
    return code, code_pos


def check_grads_tie_aware(r, want64, perms, rtol=1e-4, tau=1e-6):
    """Per-image gradient parity against the fp64 evaluation of the reference algorithm: every image WITHOUT a clamp tie
    must meet `rtol`; images with one are allowed the flip of that indicator (bounded loosely) — and there must be few."""
    B = r["d_code"].shape[0]
    ties = clamp_tie_images(want64["out"], perms, B, tau)
    worst_clean = 0.0
    for key, tied in (("d_code", ties[0]), ("d_code_pos", ties[1])):
        for b in range(B):
            e = rel_err(r[key][b], want64[key][b])
            if b in tied:
                assert e < 5e-2, (key, b, e)
            else:
                worst_clean = max(worst_clean, e)
                assert e < rtol, (key, b, e, sorted(tied))
        assert rel_err(r[key], want64[key]) < 5e-3, key
        assert len(tied) <= max(2, B // 2), (key, sorted(tied))
    return worst_clean, ties


def clamp_tie_pixel_masks(out64, coords1, coords2, perms, H, W, tau=1e-6):
    """Pixel-level version of clamp_tie_images for small cases: boolean masks [B,H,W] over code / code_pos of the pixels
    a flipped clamp indicator can move.  cd[b,h,w,i,j] pairs the sample at coords1[b,w,h] (the reference's S-axis swap)
    with the one at coords2[b,j,i]; each sample spreads over its <= 4 bilinear corner pixels."""
    B, S = coords1.shape[0], coords1.shape[1]
    m_code, m_pos = np.zeros((B, H, W), bool), np.zeros((B, H, W), bool)

    def mark(mask, b, xy):
        ix = min(max((xy[0] + 1.0) / 2.0 * (W - 1), 0.0), W - 1.0)
        iy = min(max((xy[1] + 1.0) / 2.0 * (H - 1), 0.0), H - 1.0)
        for yy in (int(np.floor(iy)), min(int(np.floor(iy)) + 1, H - 1)):
            for xx in (int(np.floor(ix)), min(int(np.floor(ix)) + 1, W - 1)):
                mask[b, yy, xx] = True

    def scan(cd, second_coords, second_target):
        """cd: [n*B,S,S,S,S]; the first sample always belongs to code[b] at coords1[b]."""
        idx = np.argwhere(np.abs(cd.detach().numpy()) < tau)
        for nb, h, w, i, j in idx:
            b = int(nb) % B
            mark(m_code, b, coords1[b][w][h])
            sm, sb = second_target(int(nb))
            mark(sm, sb, second_coords[b][j][i])
        return len(idx)

    n = scan(out64[1], coords1, lambda nb: (m_code, nb))                             # intra: code[b] twice
    n += scan(out64[3], coords2, lambda nb: (m_pos, nb))                             # inter: code_pos[b]
    n += scan(out64[5], coords2, lambda nb: (m_code, int(perms[nb // B][nb % B])))   # negatives: code[perm_n[b]]
    return m_code, m_pos, n


def masked_rel_err(a, b, mask):
    keep = ~mask[:, None, :, :]
    a, b = np.asarray(a, np.float64) * keep, np.asarray(b, np.float64) * keep
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
