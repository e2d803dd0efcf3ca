"""Generate tests/golden/*.npz by running the REAL reference (/root/reference,
imported read-only with stub modules — see tests/refimport.py) on the seeded
inputs of tests/golden/cases.py.  Run in the build container only:

    python -m tests.golden.make_golden

The reference has no golden vectors of its own, so these
files are what pins both the oracle and the CUDA path.  Torch 2.11 CPU fp32,
NumPy 2.3.
"""
import contextlib
import os

import numpy as np
import torch

from tests import refimport
from tests.golden import cases

HERE = os.path.dirname(os.path.abspath(__file__))


@contextlib.contextmanager
def injected(M, perms, rands):
    """Feed the reference the permutations / random coordinates of the case
    instead of its own RNG draws (same call order: rand(coords1), rand(coords2),
    then one super_perm per negative; src/modules.py:1320-1321, :1341)."""
    perm_it, rand_it = iter(perms), iter(rands)
    old_perm, old_rand = M.super_perm, torch.rand
    M.super_perm = lambda size, device: next(perm_it).clone()
    torch.rand = lambda shape, device=None: next(rand_it).clone()
    try:
        yield
    finally:
        M.super_perm, torch.rand = old_perm, old_rand


def run_loss_case(M, name):
    cfg, t = cases.make_loss_inputs(name)
    S = cfg.feature_samples
    code = t["code"].clone().requires_grad_(True)
    code_pos = t["code_pos"].clone().requires_grad_(True)
    loss_fn = M.ContrastiveCorrelationLoss(cfg)
    with injected(M, list(t["perms"]), [t["rand1"], t["rand2"]]):
        out = loss_fn(t["feats"], t["feats_pos"], None, None, code, code_pos, t["depth"], t["depth_pos"])
    w = cases.LOSS_WEIGHTS
    L = w["pos_intra"] * out[0] + w["pos_inter"] * out[2] + w["neg_inter"] * out[4].mean()
    if cfg.depth_feat_correlation_loss:
        L = L + w["depth_feat"] * out[6]
    L.backward()
    if cfg.depth_sampling == "fps":
        c1 = M.farthest_point_sampling_depth(t["feats"], t["depth"], S) * 2 - 1
        c2 = M.farthest_point_sampling_depth(t["feats_pos"], t["depth_pos"], S) * 2 - 1
    else:
        c1, c2 = t["rand1"] * 2 - 1, t["rand2"] * 2 - 1
    g = dict(coords1=c1.numpy(), coords2=c2.numpy(),
             scalars=np.array([out[0].item(), out[2].item(), out[4].mean().item(),
                               out[6].item() if len(out) == 8 else np.nan], np.float64),
             cd_means=np.array([out[1].mean().item(), out[3].mean().item(), out[5].mean().item(),
                                out[7].mean().item() if len(out) == 8 else np.nan], np.float64),
             total=np.float64(L.item()), d_code=code.grad.numpy(), d_code_pos=code_pos.grad.numpy())
    if S <= 8:
        g.update(intra_cd=out[1].detach().numpy(), inter_cd=out[3].detach().numpy(),
                 neg_loss=out[4].detach().numpy(), neg_cd=out[5].detach().numpy())
        if len(out) == 8:
            g["depth_dd"] = out[7].detach().numpy()
    np.savez_compressed(os.path.join(HERE, f"loss_{name}.npz"), **g)
    print(name, g["scalars"], g["cd_means"], float(np.linalg.norm(g["d_code"])), float(np.linalg.norm(g["d_code_pos"])))


def run_aug_case(M, name):
    cfg, t = cases.make_aug_inputs(name)
    code = t["code"].clone().requires_grad_(True)
    code_pos = t["code_pos"].clone().requires_grad_(True)
    loss_fn = M.DepthContrastiveCorrelationLoss(cfg)
    with injected(M, list(t["perms"]), [t["rand1"], t["rand2"]]):
        out = loss_fn(t["feats"], t["feats_pos"], None, None, code, code_pos, t["aug"], t["aug_pos"])
    w = cases.LOSS_WEIGHTS
    L = w["pos_intra"] * out[0] + w["pos_inter"] * out[2] + w["neg_inter"] * out[4].mean()
    L.backward()
    g = dict(scalars=np.array([out[0].item(), out[2].item(), out[4].mean().item()], np.float64),
             cd_means=np.array([out[1].mean().item(), out[3].mean().item(), out[5].mean().item()], np.float64),
             total=np.float64(L.item()), d_code=code.grad.numpy(), d_code_pos=code_pos.grad.numpy())
    np.savez_compressed(os.path.join(HERE, f"aug_{name}.npz"), **g)
    print(name, g["scalars"], float(np.linalg.norm(g["d_code"])))


def run_sal_case(M, name):
    """use_salience sampling through the real reference (both loss classes share the code, :1291-1298 / :1413-1420)."""
    cfg, t = cases.make_sal_inputs(name)
    code = t["code"].clone().requires_grad_(True)
    code_pos = t["code_pos"].clone().requires_grad_(True)
    loss_fn = M.ContrastiveCorrelationLoss(cfg)
    old_randint = torch.randint
    ri = cases.randint_from_uniforms(list(t["randint_u"]))
    torch.randint = lambda high, size, device=None: ri(high, size, device)
    try:
        with injected(M, list(t["perms"]), [t["rand1"], t["rand2"], t["rand_mask"]]):
            out = loss_fn(t["feats"], t["feats_pos"], t["salience"], t["salience_pos"], code, code_pos, t["depth"],
                          t["depth_pos"])
    finally:
        torch.randint = old_randint
    w = cases.LOSS_WEIGHTS
    L = w["pos_intra"] * out[0] + w["pos_inter"] * out[2] + w["neg_inter"] * out[4].mean()
    if cfg.depth_feat_correlation_loss:
        L = L + w["depth_feat"] * out[6]
    L.backward()
    # the coordinates the reference built, recomputed with its own function under the same draws
    ri = cases.randint_from_uniforms(list(t["randint_u"]))
    torch.randint = lambda high, size, device=None: ri(high, size, device)
    try:
        S = cfg.feature_samples
        shape = [t["feats"].shape[0], S, S, 2]
        n1, n2 = M.sample_nonzero_locations(t["salience"], shape), M.sample_nonzero_locations(t["salience_pos"], shape)
    finally:
        torch.randint = old_randint
    mask = (t["rand_mask"] > .1).unsqueeze(-1).to(torch.float32)
    c1 = n1 * mask + (t["rand1"] * 2 - 1) * (1 - mask)
    c2 = n2 * mask + (t["rand2"] * 2 - 1) * (1 - mask)
    g = dict(coords1=c1.numpy(), coords2=c2.numpy(), nonzero1=n1.numpy(), nonzero2=n2.numpy(),
             scalars=np.array([out[0].item(), out[2].item(), out[4].mean().item(),
                               out[6].item() if len(out) == 8 else np.nan], np.float64),
             cd_means=np.array([out[1].mean().item(), out[3].mean().item(), out[5].mean().item(),
                                out[7].mean().item() if len(out) == 8 else np.nan], np.float64),
             total=np.float64(L.item()), d_code=code.grad.numpy(), d_code_pos=code_pos.grad.numpy())
    np.savez_compressed(os.path.join(HERE, f"sal_{name}.npz"), **g)
    print(name, g["scalars"], float(np.linalg.norm(g["d_code"])), float(np.linalg.norm(g["d_code_pos"])))


def run_fps(M):
    out = {}
    for pat in cases.FPS_PATTERNS:
        depth = cases.make_fps_depth(pat)
        t = torch.zeros(depth.shape[0], 1, 28, 28)
        for S in cases.FPS_S:
            coords = M.farthest_point_sampling_depth(t, depth, S)          # [B,S,S,2] in [0,1)
            rc = torch.round(coords.reshape(depth.shape[0], S * S, 2) * 28).long()
            out[f"{pat}_S{S}"] = (rc[..., 0] * 28 + rc[..., 1]).numpy().astype(np.int32)
            out[f"{pat}_S{S}_coords"] = coords.numpy()
    np.savez_compressed(os.path.join(HERE, "fps_index_sets.npz"), **out)
    print("fps:", len(out), "arrays")


def run_misc(M):
    """Small known-answer vectors for the free functions."""
    rs = np.random.RandomState(77)
    t = torch.from_numpy(rs.standard_normal((2, 5, 28, 28)).astype(np.float32))
    coords = torch.from_numpy((rs.random_sample((2, 4, 4, 2)) * 2.4 - 1.2).astype(np.float32))  # some out of range
    depth = torch.from_numpy(rs.randint(0, 256, (2, 1, 224, 224)).astype(np.float32))
    depth[0, 0, :100] = 0
    d7 = torch.nn.functional.interpolate(depth, size=(7, 7), mode="bilinear", align_corners=True)
    pts = M.depth2points(torch.nn.functional.adaptive_avg_pool2d(depth, (28, 28))[1, 0], fov=90)
    np.savez_compressed(os.path.join(HERE, "misc.npz"),
                        sample_out=M.sample(t, coords).numpy(), norm_out=M.norm(t).numpy(),
                        corr_out=M.tensor_correlation(t[:, :, :3, :3], t[:, :, 5:9, 5:9]).numpy(),
                        depth_sign7=M.norm(d7).numpy(), points=pts.numpy())


def run_knn():
    """precompute_knns.py cannot be imported (hydra/lightning absent); its loop
    (src/precompute_knns.py:99-113) is stock torch and is restated verbatim here."""
    out = {}
    for name in cases.KNN_CASES:
        normed_feats, k, n_batches = cases.make_knn_feats(name)
        all_nns = []
        step = normed_feats.shape[0] // n_batches
        for i in range(0, normed_feats.shape[0], step):
            batch_feats = normed_feats[i:i + step, :]
            pairwise_sims = torch.einsum("nf,mf->nm", batch_feats, normed_feats)
            all_nns.append(torch.topk(pairwise_sims, k)[1])
        nn_idx = torch.cat(all_nns, dim=0)
        out[name] = nn_idx.numpy()
        out[name + "_vals"] = torch.gather(normed_feats @ normed_feats.T, 1, nn_idx).numpy()
    np.savez_compressed(os.path.join(HERE, "knn.npz"), **out)
    print("knn:", {k: v.shape for k, v in out.items()})


def run_probes(M):
    """ClusterLookup is the reference's own class (src/modules.py:646-675).  The linear-probe
    branch lives inside LitUnsupervisedSegmenter.training_step (src/train_segmentation.py:419-437),
    which cannot be instantiated here (lightning/hydra/datasets absent); its stock-torch call
    sequence is replayed line by line on an nn.Conv2d / CrossEntropyLoss built as in :113,:127."""
    out = {}
    for name in cases.PROBE_CASES:
        t = cases.make_probe_inputs(name)
        K, D = t["weight"].shape
        probe = M.ClusterLookup(D, K)
        with torch.no_grad():
            probe.clusters.copy_(t["clusters"])
        loss, probs = probe(t["code"], None)
        loss.backward()
        out[name + "_cluster_loss"] = loss.detach().numpy()
        assert bool(((probs == 0) | (probs == 1)).all()) and bool((probs.sum(1) == 1).all())
        out[name + "_cluster_argmax"] = probs.argmax(1).numpy().astype(np.uint8)   # the one-hot map, compactly
        out[name + "_cluster_grad"] = probe.clusters.grad.numpy()
        with torch.no_grad():
            sl, sp = probe(t["code"], 2)
            out[name + "_soft_loss"] = sl.numpy()
            if t["code"].numel() < 50000:   # keep the fixtures small: full maps only for the small cases
                out[name + "_soft_probs"] = sp.numpy()
                out[name + "_log_probs"] = probe(t["code"], 2, log_probs=True).numpy()

        linear_probe = torch.nn.Conv2d(D, K, (1, 1))
        with torch.no_grad():
            linear_probe.weight.copy_(t["weight"].reshape(K, D, 1, 1))
            linear_probe.bias.copy_(t["bias"])
        linear_probe_loss_fn = torch.nn.CrossEntropyLoss()
        label, code, n_classes = t["label"], t["code"], K
        flat_label = label.reshape(-1)
        mask = (flat_label >= 0) & (flat_label < n_classes)
        detached_code = torch.clone(code.detach())
        linear_logits = linear_probe(detached_code)
        linear_logits = torch.nn.functional.interpolate(linear_logits, label.shape[-2:], mode='bilinear',
                                                        align_corners=False)
        linear_logits = linear_logits.permute(0, 2, 3, 1).reshape(-1, n_classes)
        linear_loss = linear_probe_loss_fn(linear_logits[mask], flat_label[mask]).mean()
        linear_loss.backward()
        out[name + "_linear_loss"] = linear_loss.detach().numpy()
        out[name + "_linear_dw"] = linear_probe.weight.grad.reshape(K, D).numpy()
        out[name + "_linear_db"] = linear_probe.bias.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "probes.npz"), **out)
    print("probes:", {k: (v.shape if v.ndim else float(v)) for k, v in out.items() if "loss" in k})


def main():
    assert refimport.have_reference(), "run in the build container (needs /root/reference)"
    torch.set_num_threads(8)
    M = refimport.load_reference_modules()
    only = os.environ.get("GOLDEN_ONLY")
    if only == "probes":
        return run_probes(M)
    for name in cases.LOSS_CASES:
        if only is None or name in only.split(","):
            run_loss_case(M, name)
    for name in cases.AUG_CASES:
        if only is None or name in only.split(","):
            run_aug_case(M, name)
    for name in cases.SAL_CASES:
        if only is None or name in only.split(","):
            run_sal_case(M, name)
    if only is not None:
        return
    run_fps(M)
    run_misc(M)
    run_knn()
    run_probes(M)


if __name__ == "__main__":
    main()
