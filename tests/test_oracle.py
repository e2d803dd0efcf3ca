"""CPU tests: the oracle reproduces the REAL reference — through the committed
golden vectors everywhere, and by direct import where /root/reference exists."""
import numpy as np
import pytest
import torch

from oracle import depthg_oracle as O
from tests import refimport
from tests.golden import cases
from tests.helpers import golden, run_oracle_loss

FAST_LOSS_CASES = [n for n in cases.LOSS_CASES]


@pytest.mark.parametrize("name", FAST_LOSS_CASES)
def test_oracle_loss_matches_reference_golden(name):
    g = golden("loss_" + name)
    cfg, t, r = run_oracle_loss(name)
    assert np.array_equal(r["coords1"], g["coords1"]) and np.array_equal(r["coords2"], g["coords2"])
    np.testing.assert_allclose(r["scalars"], g["scalars"], rtol=1e-6, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(r["cd_means"], g["cd_means"], rtol=1e-6, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(r["d_code"], g["d_code"], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(r["d_code_pos"], g["d_code_pos"], rtol=1e-5, atol=1e-10)
    if "intra_cd" in g:
        out = r["out"]
        np.testing.assert_allclose(out[1].detach().numpy(), g["intra_cd"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(out[4].detach().numpy(), g["neg_loss"], rtol=1e-6, atol=1e-7)


def test_fov_factor_bits():
    f = 2.0 * torch.tan(torch.tensor([90.0]) / 2.0)
    assert f.numpy().view(np.uint32)[0] == O.FOV_FACTOR_BITS


@pytest.mark.parametrize("pattern", list(cases.FPS_PATTERNS))
def test_fps_oracle_and_masked_spec_match_reference_golden(pattern):
    g = golden("fps_index_sets")
    depth = cases.make_fps_depth(pattern)
    for S in (3, 5, 11):
        want = g[f"{pattern}_S{S}"]
        got = O.fps_index_sets((1, 1, 28, 28), depth, S)
        assert np.array_equal(got, want), (pattern, S)
        # the GPU-shaped statement (sequential pool, one rounding per op, masked argmax)
        for b in range(depth.shape[0]):
            X, Y, Z = O.pooled_points_fp32(depth[b, 0].numpy())
            assert np.array_equal(O.fps_masked(X, Y, Z, S * S), want[b]), (pattern, S, b)
        coords = O.farthest_point_sampling_depth(torch.zeros(1, 1, 28, 28), depth, S)
        assert np.array_equal(coords.numpy(), g[f"{pattern}_S{S}_coords"])


def test_zero_depth_fps_picks_first_indices():
    idx = O.fps_index_sets((1, 1, 28, 28), torch.zeros(1, 1, 224, 224), 3)
    assert idx.tolist() == [list(range(9))]


def test_free_functions_match_reference_golden():
    g = golden("misc")
    rs = np.random.RandomState(77)
    t = torch.from_numpy(rs.standard_normal((2, 5, 28, 28)).astype(np.float32))
    coords = torch.from_numpy((rs.random_sample((2, 4, 4, 2)) * 2.4 - 1.2).astype(np.float32))
    depth = torch.from_numpy(rs.randint(0, 256, (2, 1, 224, 224)).astype(np.float32))
    depth[0, 0, :100] = 0
    assert np.array_equal(O.sample(t, coords).numpy(), g["sample_out"])
    np.testing.assert_allclose(O.sample_explicit(t, coords).numpy(), g["sample_out"], rtol=0, atol=2e-5)
    assert np.array_equal(O.norm(t).numpy(), g["norm_out"])
    assert np.array_equal(O.tensor_correlation(t[:, :, :3, :3], t[:, :, 5:9, 5:9]).numpy(), g["corr_out"])
    np.testing.assert_allclose(O.depth_sign_explicit(depth, 7).numpy(), g["depth_sign7"][:, 0], atol=1e-6)
    pooled = torch.nn.functional.adaptive_avg_pool2d(depth, (28, 28))[1, 0]
    assert np.array_equal(O.depth2points(pooled, fov=90).numpy(), g["points"])
    X, Y, Z = O.pooled_points_fp32(depth[1, 0].numpy())
    assert np.array_equal(np.stack([X, Y, Z]).reshape(3, 28, 28), g["points"])


@pytest.mark.parametrize("name", list(cases.KNN_CASES))
def test_knn_oracle_matches_golden(name):
    g = golden("knn")
    feats, k, n_batches = cases.make_knn_feats(name)
    idx = O.knn_topk_chunked(feats, k, n_batches)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == g[name].shape
    assert np.array_equal(idx.numpy(), g[name])
    v, i = O.knn_rows(feats[100:164], feats, k)
    np.testing.assert_allclose(v.numpy(), g[name + "_vals"][100:164], atol=1e-6)


def test_super_perm_has_no_fixed_points_and_matches_reference_stream():
    torch.manual_seed(3)
    p = O.super_perm(9)
    assert not (p == torch.arange(9)).any()
    if refimport.have_reference():
        M = refimport.load_reference_modules()
        torch.manual_seed(3)
        assert torch.equal(p, M.super_perm(9, torch.device("cpu")))


@pytest.mark.skipif(not refimport.have_reference(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("seed,mode", [(0, "fps"), (1, "none")])
def test_oracle_equals_real_reference_with_shared_rng(seed, mode):
    """Same torch seed -> the oracle and the real reference draw the same RNG
    stream (rand, rand, randperm x N) and must agree bit-for-bit."""
    M = refimport.load_reference_modules()
    cfg = cases.loss_cfg(feature_samples=6, depth_sampling=mode)
    g = torch.Generator().manual_seed(100 + seed)
    f, fp = torch.randn(3, 24, 28, 28, generator=g), torch.randn(3, 24, 28, 28, generator=g)
    c, cp = torch.randn(3, 12, 28, 28, generator=g), torch.randn(3, 12, 28, 28, generator=g)
    d = torch.randint(0, 256, (3, 1, 224, 224), generator=g).float()
    dp = torch.randint(0, 256, (3, 1, 224, 224), generator=g).float()
    outs = []
    for impl in (M.ContrastiveCorrelationLoss, O.ContrastiveCorrelationLoss):
        torch.manual_seed(seed)
        ci, cpi = c.clone().requires_grad_(True), cp.clone().requires_grad_(True)
        out = impl(cfg)(f, fp, None, None, ci, cpi, d, dp)
        (out[0] + out[2] + out[4].mean() + out[6]).backward()
        outs.append([o.detach() for o in out] + [ci.grad, cpi.grad])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", list(cases.AUG_CASES))
def test_oracle_depth_contrastive_variant_matches_reference_golden(name):
    """DepthContrastiveCorrelationLoss (src/modules.py:1370-1463) restatement vs the real reference's output."""
    from tests.helpers import run_oracle_aug
    g = golden("aug_" + name)
    r = run_oracle_aug(name)
    np.testing.assert_allclose(r["scalars"], g["scalars"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(r["cd_means"], g["cd_means"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(r["d_code"], g["d_code"], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(r["d_code_pos"], g["d_code_pos"], rtol=1e-5, atol=1e-10)


@pytest.mark.parametrize("name", list(cases.SAL_CASES))
def test_oracle_salience_sampling_matches_reference_golden(name):
    """use_salience coordinates (src/modules.py:1191-1204, :1291-1298) and the loss on them vs the real reference."""
    from tests.helpers import run_sal
    g = golden("sal_" + name)
    r = run_sal(name, O.ContrastiveCorrelationLoss)
    assert np.array_equal(r["coords1"], g["coords1"]) and np.array_equal(r["coords2"], g["coords2"])
    np.testing.assert_allclose(r["scalars"], g["scalars"], rtol=1e-6, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(r["d_code"], g["d_code"], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(r["d_code_pos"], g["d_code_pos"], rtol=1e-5, atol=1e-10)
    # the all-zero salience map took the uniform-pixel fallback: integer pixel coordinates / H * 2 - 1
    cfg, t = cases.make_sal_inputs(name)
    n1 = O.sample_nonzero_locations(t["salience"], list(g["nonzero1"].shape), cases.randint_from_uniforms(list(t["randint_u"])))
    assert np.array_equal(n1.numpy(), g["nonzero1"])


@pytest.mark.skipif(not refimport.have_reference(), reason="/root/reference only exists in the build container")
def test_oracle_salience_and_fps_depth_feat_equal_real_reference_with_shared_rng():
    """Shared torch seed: the salience path draws randint (CPU generator) per image, then rand x3; the
    'fps_depth_feat' mode is the 'fps' call with an ignored argument.  Bit-for-bit against the live reference."""
    M = refimport.load_reference_modules()
    g = torch.Generator().manual_seed(7)
    f, fp = torch.randn(3, 24, 28, 28, generator=g), torch.randn(3, 24, 28, 28, generator=g)
    c, cp = torch.randn(3, 12, 28, 28, generator=g), torch.randn(3, 12, 28, 28, generator=g)
    d = torch.randint(0, 256, (3, 1, 224, 224), generator=g).float()
    dp = torch.randint(0, 256, (3, 1, 224, 224), generator=g).float()
    sal = (torch.rand(3, 28, 28, generator=g) > 0.6).float()
    sal[1] = 0
    salp = (torch.rand(3, 28, 28, generator=g) > 0.3).float()
    for cfg in (cases.loss_cfg(feature_samples=5, use_salience=True),
                cases.loss_cfg(feature_samples=4, depth_sampling="fps_depth_feat")):
        outs = []
        for impl in (M.ContrastiveCorrelationLoss, O.ContrastiveCorrelationLoss):
            torch.manual_seed(3)
            ci, cpi = c.clone().requires_grad_(True), cp.clone().requires_grad_(True)
            out = impl(cfg)(f, fp, sal, salp, ci, cpi, d, dp)
            (out[0] + out[2] + out[4].mean() + out[6]).backward()
            outs.append([o.detach() for o in out] + [ci.grad, cpi.grad])
        for a, b in zip(*outs):
            assert torch.equal(a, b)


# ------------------------------------------------------------------ probe losses
def _oracle_probes(name):
    t = cases.make_probe_inputs(name)
    K, D = t["weight"].shape
    probe = O.ClusterLookup(D, K)
    with torch.no_grad():
        probe.clusters.copy_(t["clusters"])
    loss, probs = probe(t["code"], None)
    loss.backward()
    weight, bias = t["weight"].clone().requires_grad_(True), t["bias"].clone().requires_grad_(True)
    lin = O.linear_probe_loss(t["code"], weight, bias, t["label"])
    lin.backward()
    return t, probe, loss, probs, lin, weight, bias


@pytest.mark.parametrize("name", list(cases.PROBE_CASES))
def test_oracle_probes_match_reference_golden(name):
    g = golden("probes")
    t, probe, loss, probs, lin, weight, bias = _oracle_probes(name)
    np.testing.assert_allclose(loss.item(), g[name + "_cluster_loss"], rtol=1e-6)
    assert np.array_equal(probs.argmax(1).numpy(), g[name + "_cluster_argmax"])
    np.testing.assert_allclose(probe.clusters.grad.numpy(), g[name + "_cluster_grad"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(lin.item(), g[name + "_linear_loss"], rtol=1e-6)
    np.testing.assert_allclose(weight.grad.numpy(), g[name + "_linear_dw"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(bias.grad.numpy(), g[name + "_linear_db"], rtol=1e-5, atol=1e-9)
    with torch.no_grad():
        sl, sp = probe(t["code"], 2)
        np.testing.assert_allclose(sl.item(), g[name + "_soft_loss"], rtol=1e-6)
        if name + "_soft_probs" in g:
            np.testing.assert_allclose(sp.numpy(), g[name + "_soft_probs"], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(probe(t["code"], 2, log_probs=True).numpy(), g[name + "_log_probs"], rtol=1e-6,
                                       atol=1e-7)


@pytest.mark.skipif(not refimport.have_reference(), reason="needs /root/reference")
def test_oracle_cluster_lookup_matches_reference_class_directly():
    M = refimport.load_reference_modules()
    t = cases.make_probe_inputs("probe_odd")
    K, D = t["weight"].shape
    ref, mine = M.ClusterLookup(D, K), O.ClusterLookup(D, K)
    with torch.no_grad():
        ref.clusters.copy_(t["clusters"])
        mine.clusters.copy_(t["clusters"])
    for alpha in (None, 0.5, 3):
        (l0, p0), (l1, p1) = ref(t["code"], alpha), mine(t["code"], alpha)
        assert torch.equal(l0, l1) and torch.equal(p0, p1)
