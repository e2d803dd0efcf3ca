"""GPU parity tests: the CUDA path (through the C ABI) against the committed
golden vectors of the REAL reference and against the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): FPS index sets and KNN indices bit-exact
(KNN ties allowed only where fp32 similarities differ by < 1e-6); loss values and
gradients within 1e-4 relative.  The scalar losses are means of +-O(0.1) terms that
cancel to O(1e-3), so a 2e-7 absolute floor is allowed next to the relative bound.
"""
import numpy as np
import pytest
import torch

from oracle import depthg_oracle as O
from tests.golden import cases
from tests.helpers import golden, rel_err, run_oracle_loss

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 2e-7


def dev():
    return torch.device("cuda:0")


# ------------------------------------------------------------------ FPS (a5-a7)
@pytest.mark.parametrize("pattern", list(cases.FPS_PATTERNS))
@pytest.mark.parametrize("S", cases.FPS_S)
def test_fps_index_sets_bit_exact_vs_reference_golden(pattern, S):
    from depthg_b200.modules import farthest_point_sampling_depth, fps_index_sets
    g = golden("fps_index_sets")
    depth = cases.make_fps_depth(pattern).to(dev())
    idx = fps_index_sets(depth, 28, 28, S).cpu().numpy()
    assert np.array_equal(idx, g[f"{pattern}_S{S}"])
    coords = farthest_point_sampling_depth(torch.zeros(depth.shape[0], 1, 28, 28, device=dev()), depth, S)
    assert tuple(coords.shape) == (depth.shape[0], S, S, 2)
    assert np.array_equal(coords.cpu().numpy(), g[f"{pattern}_S{S}_coords"])


def test_fps_full_batch_matches_oracle_and_is_raster_sorted():
    from depthg_b200.modules import fps_index_sets
    rs = np.random.RandomState(5)
    depth = torch.from_numpy(rs.randint(0, 256, (32, 1, 224, 224)).astype(np.float32))
    idx = fps_index_sets(depth.to(dev()), 28, 28, 11).cpu().numpy()
    assert idx.shape == (32, 121)
    assert (np.diff(idx, axis=1) > 0).all() and idx.min() >= 0 and idx.max() < 784
    want = O.fps_index_sets((1, 1, 28, 28), depth[:6], 11)
    assert np.array_equal(idx[:6], want)


def test_fps_non_divisible_pooling_and_other_grid():
    """adaptive_avg_pool2d window rule when Hd/H is not an integer, and a 14x14 grid."""
    from depthg_b200.modules import fps_index_sets
    rs = np.random.RandomState(9)
    depth = torch.from_numpy(rs.randint(0, 256, (2, 1, 100, 120)).astype(np.float32))
    for (H, W, S) in ((28, 28, 6), (14, 14, 5), (40, 40, 7)):
        if H > 100:
            continue
        got = fps_index_sets(depth.to(dev()), H, W, S).cpu().numpy()
        want = O.fps_index_sets((1, 1, H, W), depth, S)
        assert np.array_equal(got, want), (H, W, S)


# ------------------------------------------------------------------ gather / norm / correlation (a1-a4)
def _misc_inputs():
    rs = np.random.RandomState(77)
    t = torch.from_numpy(rs.standard_normal((2, 5, 28, 28)).astype(np.float32))
    coords = torch.from_numpy((rs.random_sample((2, 4, 4, 2)) * 2.4 - 1.2).astype(np.float32))
    return t, coords


@pytest.mark.parametrize("channels_last", [False, True])
def test_sample_and_norm_match_reference_golden(channels_last):
    from depthg_b200.modules import norm, sample, sample_norm
    g = golden("misc")
    t, coords = _misc_inputs()
    tc = t.to(dev())
    if channels_last:
        tc = tc.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    got = sample(tc, coords.to(dev())).cpu().numpy()
    np.testing.assert_allclose(got, g["sample_out"], rtol=1e-4, atol=2e-5)
    want_n = O.norm(torch.from_numpy(g["sample_out"])).numpy()
    np.testing.assert_allclose(sample_norm(tc, coords.to(dev())).cpu().numpy(), want_n, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(norm(tc).cpu().numpy(), g["norm_out"], rtol=1e-4, atol=1e-5)


def test_norm_accepts_what_the_reference_accepts_and_helpers_are_forward_only():
    """src/modules.py:789-790 is F.normalize(t, dim=1, eps=1e-10) on ANY tensor: non-square maps, 2-D / 3-D / 5-D inputs,
    NCHW and channels-last; zero vectors stay zero.  The free functions carry no backward, so they refuse inputs that
    require grad instead of silently detaching them."""
    from depthg_b200.modules import norm, sample, sample_norm, tensor_correlation
    rs = np.random.RandomState(5)
    for shape in ((3, 17, 9, 14), (4, 90), (2, 33, 50), (2, 6, 3, 4, 5), (1, 768, 28, 28)):
        t = torch.from_numpy(rs.standard_normal(shape).astype(np.float32))
        t[0, :, ...] = 0.0 if t.dim() == 2 else t[0, :, ...] * (torch.arange(t.shape[2]) > 0).view(-1, *([1] * (t.dim() - 3)))
        want = O.norm(t).numpy()
        np.testing.assert_allclose(norm(t.to(dev())).cpu().numpy(), want, rtol=1e-5, atol=1e-7)
        if t.dim() == 4:
            cl = t.to(dev()).contiguous(memory_format=torch.channels_last)
            got = norm(cl)
            assert got.stride() == cl.stride()
            np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-7)
    z = torch.zeros(2, 5, 3, 3, device=dev())
    assert torch.equal(norm(z), z)
    t = torch.randn(2, 8, 28, 28, device=dev(), requires_grad=True)
    coords = torch.rand(2, 4, 4, 2, device=dev()) * 2 - 1
    for fn, args in ((norm, (t,)), (sample, (t, coords)), (sample_norm, (t, coords)),
                     (tensor_correlation, (t[:, :, :4, :4], t[:, :, :4, :4]))):
        with pytest.raises(ValueError, match="forward-only"):
            fn(*args)
        with torch.no_grad():
            fn(*args)                      # fine without autograd


def test_sample_wide_channels_vector_path():
    """C multiple of 4, channels-last: exercises the 128-bit path; C=90: scalar path."""
    from depthg_b200.modules import sample_norm
    rs = np.random.RandomState(3)
    for C in (768, 90):
        t = torch.from_numpy(rs.standard_normal((3, C, 28, 28)).astype(np.float32))
        coords = torch.from_numpy((rs.random_sample((3, 11, 11, 2)) * 2 - 1).astype(np.float32))
        want = O.norm(O.sample(t, coords)).numpy()
        for cl in (False, True):
            tc = t.to(dev())
            if cl:
                tc = tc.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
            got = sample_norm(tc, coords.to(dev())).cpu().numpy()
            assert rel_err(got, want) < 1e-5, (C, cl)


def test_tensor_correlation_matches_reference_golden():
    from depthg_b200.modules import tensor_correlation
    rs = np.random.RandomState(21)
    a = torch.from_numpy(rs.standard_normal((3, 70, 6, 6)).astype(np.float32))
    b = torch.from_numpy(rs.standard_normal((3, 70, 6, 6)).astype(np.float32))
    got = tensor_correlation(a.to(dev()), b.to(dev())).cpu()
    want = O.tensor_correlation(a, b)
    assert got.shape == want.shape
    assert rel_err(got.numpy(), want.numpy()) < 1e-5


def test_depth_correlation_is_the_single_channel_einsum():
    """src/modules.py:812-814 (a3): the c = 1 case of tensor_correlation, on the {0, 1} depth signs the loss uses."""
    from depthg_b200.modules import depth_correlation
    rs = np.random.RandomState(22)
    a = torch.from_numpy((rs.random_sample((2, 1, 7, 7)) > 0.3).astype(np.float32))
    b = torch.from_numpy((rs.random_sample((2, 1, 7, 7)) > 0.5).astype(np.float32))
    got = depth_correlation(a.to(dev()), b.to(dev())).cpu()
    want = O.depth_correlation(a, b)
    assert got.shape == want.shape == (2, 7, 7, 7, 7)
    assert torch.equal(got, want)


def test_depth_sign_matches_reference_golden():
    import ctypes
    from depthg_b200 import _lib
    g = golden("misc")
    rs = np.random.RandomState(77)
    rs.standard_normal((2, 5, 28, 28)); rs.random_sample((2, 4, 4, 2))
    depth = torch.from_numpy(rs.randint(0, 256, (2, 1, 224, 224)).astype(np.float32))
    depth[0, 0, :100] = 0
    d = depth.to(dev())
    out = torch.empty((2, 64), device=dev())
    _lib.check(_lib.lib().dg_depth_sign(_lib.ptr(d), 2, 224, 224, 7, 1e-10, 64, _lib.ptr(out), _lib.stream_ptr()), "x")
    np.testing.assert_allclose(out[:, :49].cpu().numpy().reshape(2, 7, 7), g["depth_sign7"][:, 0], atol=1e-6)
    assert (out[:, 49:] == 0).all()


# ------------------------------------------------------------------ the loss (a9-a11)
def _check_loss(r, g, cfg, name=None):
    assert np.array_equal(r["coords1"], g["coords1"]) and np.array_equal(r["coords2"], g["coords2"])
    np.testing.assert_allclose(r["scalars"], g["scalars"], rtol=RTOL, atol=ATOL, equal_nan=True)
    np.testing.assert_allclose(r["cd_means"], g["cd_means"], rtol=RTOL, atol=ATOL, equal_nan=True)
    np.testing.assert_allclose(r["total"], g["total"], rtol=RTOL, atol=ATOL)
    if rel_err(r["d_code"], g["d_code"]) < RTOL and rel_err(r["d_code_pos"], g["d_code_pos"]) < RTOL:
        return
    # The zero clamp is a step function: a code correlation within ~1e-7 of zero lands on either side depending on the
    # fp32 summation order, in the reference's own fp32 run as much as in ours (dense cases hold millions of
    # correlations).  Such an indicator moves the gradient only at the bilinear corner pixels of its two samples: find
    # them with the fp64 evaluation of the reference algorithm and hold 1e-4 everywhere else.
    assert name is not None and cfg.zero_clamp, (rel_err(r["d_code"], g["d_code"]), rel_err(r["d_code_pos"], g["d_code_pos"]))
    from tests.helpers import clamp_tie_pixel_masks, masked_rel_err
    _, t, want64 = run_oracle_loss(name, dtype=torch.float64)
    H, W = r["d_code"].shape[-2:]
    m_code, m_pos, nties = clamp_tie_pixel_masks(want64["out"], g["coords1"], g["coords2"], t["perms"].numpy(), H, W)
    assert 0 < nties and m_code.mean() < 0.05 and m_pos.mean() < 0.05, (nties, m_code.mean(), m_pos.mean())
    assert masked_rel_err(r["d_code"], g["d_code"], m_code) < RTOL, (nties, masked_rel_err(r["d_code"], g["d_code"], m_code))
    assert masked_rel_err(r["d_code_pos"], g["d_code_pos"], m_pos) < RTOL
    assert rel_err(r["d_code"], g["d_code"]) < 5e-3 and rel_err(r["d_code_pos"], g["d_code_pos"]) < 5e-3


@pytest.mark.parametrize("name", list(cases.LOSS_CASES))
@pytest.mark.parametrize("channels_last", [False, True])
def test_loss_and_grads_match_reference_golden(name, channels_last):
    from tests.gpu_helpers import run_cuda_loss
    g = golden("loss_" + name)
    cfg, t, r = run_cuda_loss(name, channels_last=channels_last)
    _check_loss(r, g, cfg, name)
    assert r["grad_strides"][0] == r["grad_strides"][1]  # gradient comes back in the input's layout
    for o in r["out"]:
        assert o.dim() == 0  # fast mode: 0-dim means, nothing 5-D in HBM


@pytest.mark.parametrize("name", ["small_fps", "small_random", "small_fps_nodepthterm", "small_fps_stabalize"])
def test_materialized_5d_outputs_match_reference_golden(name):
    from tests.gpu_helpers import run_cuda_loss
    g = golden("loss_" + name)
    cfg, t, r = run_cuda_loss(name, materialize=True)
    _check_loss(r, g, cfg, name)
    out = r["out"]
    B, S = t["feats"].shape[0], cfg.feature_samples
    assert tuple(out[1].shape) == (B, S, S, S, S) and tuple(out[4].shape) == (cfg.neg_samples * B, S, S, S, S)
    np.testing.assert_allclose(out[1].detach().cpu().numpy(), g["intra_cd"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(out[3].detach().cpu().numpy(), g["inter_cd"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(out[4].detach().cpu().numpy(), g["neg_loss"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(out[5].detach().cpu().numpy(), g["neg_cd"], rtol=1e-4, atol=2e-6)
    if cfg.depth_feat_correlation_loss:
        np.testing.assert_allclose(out[7].detach().cpu().numpy(), g["depth_dd"], rtol=1e-4, atol=2e-6)


def test_cfg2_full_size_matches_oracle():
    """cocostuff27 ViT-B/8 training shape (B=32, C=768, D=90, S=11) against the oracle run on the host.

    At this size the reference's own fp32 gradient is only reproducible to ~1e-3: the zero-clamp is a
    step function, and one correlation among the 3.3 M whose sign differs between two fp32 evaluation
    orders moves the gradient by ~3e-4 relative (measured: fp32 oracle vs fp64 oracle = 7.6e-4).  So the
    1e-4 bound is checked where it is meaningful — against the fp64 evaluation of the same algorithm,
    relative to the reference's own fp32 error — and a 1e-3 bound is kept against the fp32 oracle."""
    from tests.gpu_helpers import run_cuda_loss
    cases.LOSS_CASES["_cfg2"] = (32, 768, 90, dict(feature_samples=11, pos_intra_shift=0.2103, pos_inter_shift=0.1233,
                                                    neg_inter_shift=0.9748, depth_feat_shift=0.0359), 99)
    try:
        inputs = cases.make_loss_inputs("_cfg2")
        cfg, t, r = run_cuda_loss("_cfg2", channels_last=True, inputs=inputs)
        _, _, want = run_oracle_loss("_cfg2")
        _, _, want64 = run_oracle_loss("_cfg2", dtype=torch.float64)
    finally:
        del cases.LOSS_CASES["_cfg2"]
    assert np.array_equal(r["coords1"], want["coords1"]) and np.array_equal(r["coords2"], want["coords2"])
    np.testing.assert_allclose(r["scalars"], want["scalars"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(r["cd_means"], want["cd_means"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(r["scalars"], want64["scalars"], rtol=RTOL, atol=ATOL)
    for key in ("d_code", "d_code_pos"):
        ref_noise = rel_err(want[key], want64[key])          # the reference's own fp32 error
        ours = rel_err(r[key], want64[key])
        assert ours < max(RTOL, 1.5 * ref_noise), (key, ours, ref_noise)
        assert rel_err(r[key], want[key]) < 1e-3, key
    # the same statement image by image: every image without a code correlation within 1e-6 of the zero clamp meets
    # 1e-4 against the fp64 evaluation (measured: ~2e-6); only images with such a tie carry the flip
    from tests.helpers import check_grads_tie_aware
    worst_clean, ties = check_grads_tie_aware(r, want64, inputs[1]["perms"].numpy(), rtol=RTOL)
    assert worst_clean < 2e-5, worst_clean


def test_loss_rng_stream_follows_reference_call_order():
    """Without hooks the module draws rand, rand, randperm x N on the device, like the reference."""
    from depthg_b200.modules import ContrastiveCorrelationLoss, super_perm
    cfg = cases.loss_cfg(feature_samples=5, depth_sampling="none", neg_samples=3)
    B = 4
    g = torch.Generator().manual_seed(0)
    f, fp = torch.randn(B, 32, 28, 28, generator=g).to(dev()), torch.randn(B, 32, 28, 28, generator=g).to(dev())
    c, cp = torch.randn(B, 16, 28, 28, generator=g).to(dev()), torch.randn(B, 16, 28, 28, generator=g).to(dev())
    d = torch.randint(0, 256, (B, 1, 224, 224), generator=g).float().to(dev())
    fn = ContrastiveCorrelationLoss(cfg, negative_sampler="torch")   # the reference's own randperm stream
    torch.manual_seed(123)
    out = fn(f, fp, None, None, c, cp, d, d)
    torch.manual_seed(123)
    c1 = torch.rand([B, 5, 5, 2], device=dev()) * 2 - 1
    c2 = torch.rand([B, 5, 5, 2], device=dev()) * 2 - 1
    perms = [super_perm(B, dev()) for _ in range(3)]
    assert torch.equal(fn.last_coords[0], c1) and torch.equal(fn.last_coords[1], c2)
    # replay on the oracle with the same draws
    ofn = O.ContrastiveCorrelationLoss(cfg)
    pit, rit = iter([p.cpu() for p in perms]), iter([((c1 + 1) / 2).cpu(), ((c2 + 1) / 2).cpu()])
    ofn.perm_fn = lambda n, device: next(pit)
    ofn.rand_fn = lambda shape, device: next(rit)
    want = ofn(f.cpu(), fp.cpu(), None, None, c.cpu(), cp.cpu(), d.cpu(), d.cpu())
    for i in (0, 2, 6):
        np.testing.assert_allclose(out[i].item(), want[i].item(), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(out[4].item(), want[4].mean().item(), rtol=1e-3, atol=1e-6)


def test_module_has_no_state_and_rereads_cfg():
    from depthg_b200.modules import ContrastiveCorrelationLoss
    cfg = cases.loss_cfg(feature_samples=4)
    fn = ContrastiveCorrelationLoss(cfg)
    assert len(fn.state_dict()) == 0 and len(list(fn.parameters())) == 0
    _, t = cases.make_loss_inputs("small_fps")
    a = [t[k].to(dev()) for k in ("feats", "feats_pos", "code", "code_pos", "depth", "depth_pos")]
    fn(a[0], a[1], None, None, a[2], a[3], depth=a[4], depth_pos=a[5])
    assert fn.last_coords.shape[2] == 4
    cfg.feature_samples = 6          # the trainer mutates cfg between steps (train_segmentation.py:356-375)
    cfg.depth_sampling = "none"
    fn(a[0], a[1], None, None, a[2], a[3], depth=a[4], depth_pos=a[5])
    assert fn.last_coords.shape[2] == 6


def test_depth_shapes_3d_accepted_and_mismatch_refused():
    """[B,Hd,Wd] depth (Potsdam, src/data.py:226) is what adaptive_avg_pool2d also accepts; a depth_pos of another
    size must raise instead of reading out of bounds."""
    from depthg_b200.modules import ContrastiveCorrelationLoss
    cfg, t = cases.make_loss_inputs("small_fps")
    a = {k: t[k].to(dev()) for k in ("feats", "feats_pos", "code", "code_pos", "depth", "depth_pos")}
    fn = ContrastiveCorrelationLoss(cfg)
    pit = iter(t["perms"].to(dev()))
    fn.perm_fn = lambda B, device: next(pit).clone()
    ref = fn(a["feats"], a["feats_pos"], None, None, a["code"], a["code_pos"], a["depth"], a["depth_pos"])
    pit = iter(t["perms"].to(dev()))
    got = fn(a["feats"], a["feats_pos"], None, None, a["code"], a["code_pos"], a["depth"][:, 0], a["depth_pos"][:, 0])
    for x, y in zip(ref, got):
        assert torch.equal(x, y)
    pit = iter(t["perms"].to(dev()))
    with pytest.raises(ValueError, match="shape of depth"):
        fn(a["feats"], a["feats_pos"], None, None, a["code"], a["code_pos"], a["depth"], a["depth_pos"][:, :, :100, :100])


def test_depth_only_intra_variant_leaves_cfg_untouched():
    """DepthContrastiveCorrelationLoss must not write to the shared cfg (read-only OmegaConf nodes, other threads)."""
    from depthg_b200.modules import DepthContrastiveCorrelationLoss

    class Frozen:
        def __init__(self, ns):
            object.__setattr__(self, "_ns", ns)

        def __getattr__(self, k):
            return getattr(object.__getattribute__(self, "_ns"), k)

        def __setattr__(self, k, v):
            raise AttributeError("cfg is read-only")

    cfg, t = cases.make_aug_inputs("aug_small")
    cfg.depth_sampling, cfg.depth_feat_correlation_loss = "fps", True      # ignored by this variant (:1413-1425)
    fn = DepthContrastiveCorrelationLoss(Frozen(cfg))
    a = {k: v.to(dev()) for k, v in t.items() if k not in ("perms", "rand1", "rand2")}
    out = fn(a["feats"], a["feats_pos"], None, None, a["code"], a["code_pos"], a["aug"], a["aug_pos"])
    assert len(out) == 6 and all(torch.isfinite(o).all() for o in out)


# ------------------------------------------------------------------ KNN (a12-a14)
def _check_knn(idx, feats, want_idx, k):
    sims = feats @ feats.T
    got_v = torch.gather(sims, 1, torch.from_numpy(idx)).numpy()
    want_v = torch.gather(sims, 1, torch.from_numpy(want_idx)).numpy()
    mism = idx != want_idx
    # ties: positions may differ only where the fp32 similarities differ by < 1e-6
    assert np.all(np.abs(got_v - want_v)[mism] < 1e-6), f"{mism.sum()} slots differ beyond the tie rule"
    assert (np.diff(got_v, axis=1) <= 1e-6).all()
    for r in np.nonzero(mism.any(1))[0][:50]:
        assert len(set(idx[r].tolist())) == k
    return int(mism.sum())


def _check_knn_rows(idx, queries, db, want_idx):
    sims = queries @ db.T
    gv = torch.gather(sims, 1, torch.from_numpy(idx)).numpy()
    wv = torch.gather(sims, 1, torch.from_numpy(want_idx)).numpy()
    mism = idx != want_idx
    assert np.all(np.abs(gv - wv)[mism] < 1e-6), f"{mism.sum()} slots differ beyond the tie rule"


@pytest.mark.parametrize("k", [30, 7])
@pytest.mark.parametrize("path", ["umma", "umma_split3", "umma_mq2", "simt"])
def test_knn_tensor_core_and_exact_kernels_agree(path, k, monkeypatch):
    """All KNN paths (tcgen05 fast fp16 pass = default, tcgen05 3-term bf16 pass, exact CUDA-core kernel) against the
    oracle on a problem big enough for the tcgen05 path, incl. ragged sizes."""
    from depthg_b200.precompute_knns import knn_topk
    if path == "simt":
        monkeypatch.setenv("DEPTHG_B200_KNN", "simt")
    if path == "umma_split3":
        monkeypatch.setenv("DEPTHG_B200_KNN_PASS", "split3")
    if path == "umma_mq2":      # the 256 x 256 tile variant of the fast pass (two query blocks per CTA)
        monkeypatch.setenv("DEPTHG_B200_KNN_MQ", "2")
    rs = np.random.RandomState(41)
    cent = rs.standard_normal((40, 200)).astype(np.float32)
    x = cent[rs.randint(0, 40, 4100)] + 0.25 * rs.standard_normal((4100, 200)).astype(np.float32)
    x = torch.nn.functional.normalize(torch.from_numpy(x), dim=1)
    idx, sims = knn_topk(x[:1500].to(dev()), x.to(dev()), k, return_sims=True)
    _, want = O.knn_rows(x[:1500], x, k)
    _check_knn_rows(idx.cpu().numpy(), x[:1500], x, want.numpy())
    assert (np.diff(sims.cpu().numpy(), axis=1) <= 0).all()


def test_knn_certificate_failure_falls_back_to_exact_kernel():
    """Massive exact ties (duplicated rows) defeat the candidate certificate; the flagged rows must be
    recomputed by the exact kernel and still satisfy the tie rule."""
    from depthg_b200.precompute_knns import knn_topk
    rs = np.random.RandomState(43)
    base = torch.nn.functional.normalize(torch.from_numpy(rs.standard_normal((64, 96)).astype(np.float32)), dim=1)
    x = base.repeat(40, 1)                      # 2560 rows, every vector appears 40 times
    idx = knn_topk(x.to(dev()), x.to(dev()), 30).cpu().numpy()
    sims = (x @ x.T)
    got = torch.gather(sims, 1, torch.from_numpy(idx)).numpy()
    assert np.all(got > 1 - 1e-5)               # all 30 neighbours are copies of the query
    assert all(len(set(r.tolist())) == 30 for r in idx[:200])


@pytest.mark.parametrize("name", list(cases.KNN_CASES))
def test_knn_indices_match_reference_golden(name):
    from depthg_b200.precompute_knns import build_knn_index, knn_topk
    g = golden("knn")
    feats, k, n_batches = cases.make_knn_feats(name)
    idx = build_knn_index(feats.to(dev()), k, n_batches)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == g[name].shape
    nm = _check_knn(idx.cpu().numpy(), feats, g[name], k)
    assert nm <= 0.001 * idx.numel()
    # query-row shard against the full database == the same rows of the full build
    lo, hi = 130, 517
    part, sims = knn_topk(feats[lo:hi].to(dev()), feats.to(dev()), k, return_sims=True)
    _check_knn_rows(part.cpu().numpy(), feats[lo:hi], feats, g[name][lo:hi])
    np.testing.assert_allclose(sims.cpu().numpy(), g[name + "_vals"][lo:hi], atol=2e-6)


def test_knn_larger_unaligned_problem_vs_oracle():
    from depthg_b200.precompute_knns import knn_topk
    rs = np.random.RandomState(31)
    x = torch.nn.functional.normalize(torch.from_numpy(rs.standard_normal((5003, 768)).astype(np.float32)), dim=1)
    idx = knn_topk(x[:777].to(dev()), x.to(dev()), 30).cpu().numpy()
    _, want = O.knn_rows(x[:777], x, 30)
    sims = x[:777] @ x.T
    gv = torch.gather(sims, 1, torch.from_numpy(idx)).numpy()
    wv = torch.gather(sims, 1, want).numpy()
    mism = idx != want.numpy()
    assert np.all(np.abs(gv - wv)[mism] < 1e-6)
    assert (idx[:, 0] == np.arange(777)).all()  # every image is its own nearest neighbour


def test_pool_normalize_matches_get_feats():
    from depthg_b200.precompute_knns import pool_normalize
    rs = np.random.RandomState(8)
    fm = torch.from_numpy(rs.standard_normal((5, 96, 7, 7)).astype(np.float32))
    want = O.pooled_normed_feats(fm).numpy()
    for cl in (False, True):
        x = fm.to(dev())
        if cl:
            x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        assert rel_err(pool_normalize(x).cpu().numpy(), want) < 1e-6


def test_precompute_and_save_writes_the_file_the_dataset_loads(tmp_path):
    """get_feats pooling -> KNN build -> nns_*.npz, read back the way src/data.py:1056-1064 does."""
    from depthg_b200.precompute_knns import load_nns, precompute_and_save
    rs = np.random.RandomState(12)
    fm = torch.from_numpy(rs.standard_normal((700, 64, 5, 5)).astype(np.float32))
    path = precompute_and_save(fm.to(dev()), str(tmp_path), "vit_base", "cocostuff27", "val", "five")
    assert path.endswith("nns/nns_vit_base_cocostuff27_val_five_392.npz")
    nns = load_nns(path)
    assert nns.dtype == np.int64 and nns.shape == (700, 30)
    feats = O.pooled_normed_feats(fm)
    want = O.knn_topk_chunked(feats, 30, 64).numpy()
    _check_knn(nns, feats, want, 30)


# ------------------------------------------------------------------ tcgen05 kernel vs the generic CUDA-core kernel
def _run_both_kernels(name, monkeypatch):
    """Same inputs through the tcgen05 kernel and (DEPTHG_B200_CORR=simt) the generic kernel."""
    from depthg_b200 import modules as M
    from tests.gpu_helpers import run_cuda_loss
    res = {}
    for kind in ("umma", "simt"):
        if kind == "simt":
            monkeypatch.setenv("DEPTHG_B200_CORR", "simt")
        else:
            monkeypatch.delenv("DEPTHG_B200_CORR", raising=False)
        M._CorrLossFn.debug = True
        try:
            cfg, t, r = run_cuda_loss(name, materialize=True)
        finally:
            M._CorrLossFn.debug = False
        r["dC1"], r["dC2"] = [x.clone() for x in M._CorrLossFn.last_unit_grads]
        r["fd"] = M._CorrLossFn.last_fd.clone() if kind == "umma" else None
        res[kind] = r
    monkeypatch.delenv("DEPTHG_B200_CORR", raising=False)
    return cfg, t, res


@pytest.mark.parametrize("name", ["small_fps", "small_random_pointwise", "cfg1_vits", "small_fps_noclamp", "s12_fps",
                                  "dense_14x14", "cfg4_cityscapes", "dense_400_random"])
def test_tcgen05_kernel_matches_generic_kernel_stage_by_stage(name, monkeypatch):
    from depthg_b200.modules import corr_kernel_choice
    cfg, t, res = _run_both_kernels(name, monkeypatch)
    S, D = cfg.feature_samples, t["code"].shape[1]
    assert corr_kernel_choice(S * S, D) == "umma"
    u, s = res["umma"], res["simt"]
    P = S * S
    # (1) feature correlations straight out of TMEM vs an fp64 product of the oracle's normalised samples
    c1, c2 = torch.from_numpy(u["coords1"]), torch.from_numpy(u["coords2"])
    f1 = O.norm(O.sample(t["feats"], c1)).double().flatten(2)            # [B,C,P]
    f2 = O.norm(O.sample(t["feats_pos"], c2)).double().flatten(2)
    fd_inter = torch.einsum("bcp,bcq->bpq", f1, f2).numpy()
    fd_intra = torch.einsum("bcp,bcq->bpq", f1, f1).numpy()
    got = u["fd"].cpu().numpy()
    np.testing.assert_allclose(got[0][:, :P, :P], fd_intra, atol=2e-5)
    np.testing.assert_allclose(got[1][:, :P, :P], fd_inter, atol=2e-5)
    if P % 128:
        assert np.abs(got[:, :, P:, :]).max() == 0 and np.abs(got[:, :, :, P:]).max() == 0   # zero padding rows
    # (2) code correlations (tf32 x3 split must be fp32-grade) and the dense loss tensors
    for i, atol in ((1, 1e-6), (3, 1e-6), (5, 1e-6), (4, 1e-5)):   # cd tensors fp32-grade; loss carries the bf16x3 fd error
        np.testing.assert_allclose(u["out"][i].detach().cpu().numpy(), s["out"][i].detach().cpu().numpy(),
                                   rtol=1e-4, atol=atol)
    # (3) unit gradients of every pair (and the depth term) from the tensor-core GEMMs
    for key in ("dC1", "dC2"):
        a, b_ = u[key].cpu().numpy()[:, :, :P], s[key].cpu().numpy()[:, :, :P]
        if P % 128:
            assert np.abs(u[key].cpu().numpy()[:-1, :, P:]).max() == 0
        for k in range(a.shape[0]):
            if np.linalg.norm(b_[k]) > 0:
                assert rel_err(a[k], b_[k]) < 5e-5, (key, k, rel_err(a[k], b_[k]))
    # (4) end results
    np.testing.assert_allclose(u["scalars"], s["scalars"], rtol=RTOL, atol=ATOL, equal_nan=True)
    assert rel_err(u["d_code"], s["d_code"]) < RTOL and rel_err(u["d_code_pos"], s["d_code_pos"]) < RTOL


def test_generic_kernel_still_matches_reference_golden(monkeypatch):
    from tests.gpu_helpers import run_cuda_loss
    monkeypatch.setenv("DEPTHG_B200_CORR", "simt")
    for name in ("small_fps", "cfg1_vits", "small_random"):
        cfg, t, r = run_cuda_loss(name)
        _check_loss(r, golden("loss_" + name), cfg, name)


# ------------------------------------------------------------------ negative sampler / backprop forms
def test_fused_negative_sampler_properties():
    from depthg_b200.modules import fused_super_perms
    torch.manual_seed(11)
    a = fused_super_perms(5, 32, dev())
    b = fused_super_perms(5, 32, dev())
    torch.manual_seed(11)
    a2 = fused_super_perms(5, 32, dev())
    assert a.dtype == torch.int64 and tuple(a.shape) == (5, 32)
    assert torch.equal(a, a2) and not torch.equal(a, b)          # reproducible under manual_seed, advances the generator
    ar = torch.arange(32, device=dev())
    assert not (a == ar).any() and a.min() >= 0 and a.max() < 32  # super_perm's no-fixed-point property
    # before the bump every row was a permutation: at most the bumped entries collide
    for row in a.cpu():
        assert len(set(row.tolist())) >= 32 - 3
    big = fused_super_perms(64, 7, dev()).cpu()
    assert not (big == torch.arange(7)).any()
    counts = torch.stack([(big == v).sum(0) for v in range(7)]).float()   # roughly uniform over positions
    assert counts.max() < 30


def test_fused_sampler_loss_runs_and_matches_oracle_on_its_own_perms():
    """The default sampler: the forward draws the permutations itself (one extra CTA of its FPS launch) from the Philox
    stream ``fused_super_perms`` uses - same permutations under the same seed - and the loss on them equals the oracle's."""
    from depthg_b200.modules import ContrastiveCorrelationLoss, fused_super_perms
    cfg, t = cases.make_loss_inputs("small_fps")
    fn = ContrastiveCorrelationLoss(cfg, negative_sampler="fused")
    a = {k: t[k].to(dev()) for k in ("feats", "feats_pos", "code", "code_pos", "depth", "depth_pos")}
    torch.manual_seed(77)
    out = fn(a["feats"], a["feats_pos"], None, None, a["code"], a["code_pos"], a["depth"], a["depth_pos"])
    perms = fn.last_perms.cpu()
    torch.manual_seed(77)
    assert torch.equal(perms, fused_super_perms(int(cfg.neg_samples), a["feats"].shape[0], dev()).cpu())
    B = a["feats"].shape[0]
    assert int(perms.min()) >= 0 and int(perms.max()) < B
    assert not bool((perms == torch.arange(B)).any())          # no image is its own negative
    ofn = O.ContrastiveCorrelationLoss(cfg)
    pit = iter(perms)
    ofn.perm_fn = lambda B, device: next(pit)
    want = ofn(t["feats"], t["feats_pos"], None, None, t["code"], t["code_pos"], t["depth"], t["depth_pos"])
    np.testing.assert_allclose(out[4].item(), want[4].mean().item(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out[0].item(), want[0].item(), rtol=RTOL, atol=ATOL)
    # without FPS there is no launch to ride on: the forward falls back to the sampler's own launch, same stream
    cfg2, t2 = cases.make_loss_inputs("small_fps")
    cfg2.depth_sampling = "none"
    fn2 = ContrastiveCorrelationLoss(cfg2, negative_sampler="fused")
    fn2.rand_fn = lambda shape, device: torch.full(shape, 0.25, device=device)
    torch.manual_seed(77)
    fn2(a["feats"], a["feats_pos"], None, None, a["code"], a["code_pos"], a["depth"], a["depth_pos"])
    assert torch.equal(fn2.last_perms.cpu(), perms)


def test_grad_tensors_backprop_equals_weighted_sum_backward():
    """bench.backprop() (autograd.backward with the loss weights as grad_tensors) == weighted(out).backward()."""
    import bench
    from depthg_b200.modules import ContrastiveCorrelationLoss
    cfg, t = cases.make_loss_inputs("small_fps")
    grads = []
    for mode in ("sum", "grad_tensors"):
        fn = ContrastiveCorrelationLoss(cfg)
        pit = iter(t["perms"].to(dev()))
        fn.perm_fn = lambda B, device: next(pit).clone()
        a = {k: t[k].to(dev()) for k in ("feats", "feats_pos", "depth", "depth_pos")}
        code = t["code"].to(dev()).requires_grad_(True)
        code_pos = t["code_pos"].to(dev()).requires_grad_(True)
        out = fn(a["feats"], a["feats_pos"], None, None, code, code_pos, a["depth"], a["depth_pos"])
        if mode == "sum":
            bench.weighted(out).backward()
        else:
            bench.backprop(out)
        grads.append((code.grad.clone(), code_pos.grad.clone()))
    assert rel_err(grads[0][0].cpu().numpy(), grads[1][0].cpu().numpy()) < 1e-6
    assert rel_err(grads[0][1].cpu().numpy(), grads[1][1].cpu().numpy()) < 1e-6


def test_partial_and_repeated_backward_use_fresh_zeroed_gradient_buffers():
    """The forward pre-zeroes the code-gradient buffers of the FIRST backward inside its own launches and leaves
    undefined upstream gradients undefined (no zero tensors materialised for unused outputs): (a) backprop through one
    loss only == the same loss with explicit zero weights for the others; (b) a second backward over a retained graph
    accumulates exactly one more copy of the gradient; (c) a forward under no_grad still returns the same values."""
    from depthg_b200.modules import ContrastiveCorrelationLoss
    cfg, t = cases.make_loss_inputs("small_fps")
    a = {k: t[k].to(dev()) for k in ("feats", "feats_pos", "depth", "depth_pos")}

    def run(how):
        fn = ContrastiveCorrelationLoss(cfg)
        pit = iter(t["perms"].to(dev()))
        fn.perm_fn = lambda B, device: next(pit).clone()
        code = t["code"].to(dev()).requires_grad_(True)
        code_pos = t["code_pos"].to(dev()).requires_grad_(True)
        if how == "nograd":
            with torch.no_grad():
                out = fn(a["feats"], a["feats_pos"], None, None, code, code_pos, a["depth"], a["depth_pos"])
            return [float(out[i].mean()) for i in (0, 2, 4, 6)], None, None
        out = fn(a["feats"], a["feats_pos"], None, None, code, code_pos, a["depth"], a["depth_pos"])
        one, zero = torch.ones((), device=dev()), torch.zeros((), device=dev())
        if how == "inter_only":
            out[2].backward()
        elif how == "inter_zero_weights":
            torch.autograd.backward([out[0], out[2], out[4].mean(), out[6]], grad_tensors=[zero, one, zero, zero])
        elif how == "twice":
            L = out[0] + out[2] + out[4].mean() + out[6]
            L.backward(retain_graph=True)
            L.backward()
        else:
            (out[0] + out[2] + out[4].mean() + out[6]).backward()
        torch.cuda.synchronize()
        return [float(out[i].mean()) for i in (0, 2, 4, 6)], code.grad.cpu().numpy(), code_pos.grad.cpu().numpy()

    v_once, g_once, gp_once = run("once")
    v_ng, _, _ = run("nograd")
    assert v_ng == v_once
    _, g1, gp1 = run("inter_only")
    _, g2, gp2 = run("inter_zero_weights")
    assert np.abs(g1).max() > 0 and rel_err(g1, g2) < 1e-6 and rel_err(gp1, gp2) < 1e-6
    _, g_twice, gp_twice = run("twice")
    assert rel_err(g_twice, 2 * g_once) < 1e-6 and rel_err(gp_twice, 2 * gp_once) < 1e-6


def _sampling_ahead_inputs(seed, B=4, C=128, D=32):
    g = torch.Generator().manual_seed(seed)
    feat = lambda ch: torch.randn((B, 28, 28, ch), generator=g).permute(0, 3, 1, 2).to(dev())  # noqa: E731
    depth = lambda: torch.randint(0, 256, (B, 1, 224, 224), generator=g).float().to(dev())     # noqa: E731
    return dict(feats=feat(C), feats_pos=feat(C), code=feat(D), code_pos=feat(D), depth=depth(), depth_pos=depth())


@pytest.mark.parametrize("how", ["ride_along", "side_stream", "ride_along_no_ride_kernel"])
def test_sampling_done_ahead_of_the_forward_changes_nothing(how, monkeypatch):
    """queue_next_sampling (the next batch's FPS rides as extra CTAs of this forward's correlation kernel) and
    prefetch_sampling (dg_loss_presample on a side stream): the forward that consumes the pre-computed coordinates,
    depth signs and permutations must give what a plain forward gives on the same permutations - bit-identical
    coordinates (src/modules.py:999-1037), losses and gradients equal up to the atomics' summation order.  A forward
    that receives OTHER depth tensors than the ones sampled ahead must ignore the stale sampling."""
    from depthg_b200.modules import ContrastiveCorrelationLoss
    from types import SimpleNamespace
    if how == "ride_along_no_ride_kernel":
        monkeypatch.setenv("DEPTHG_B200_NO_RIDE", "1")      # the next_* job as a launch of its own behind the forward
    cfg = SimpleNamespace(feature_samples=11, neg_samples=3, depth_sampling="fps", pointwise=True, zero_clamp=True,
                          stabalize=False, use_salience=False, depth_feat_correlation_loss=True, pos_intra_shift=0.18,
                          pos_inter_shift=0.12, neg_inter_shift=0.46, depth_feat_shift=0.0)
    s1, s2, s3 = (_sampling_ahead_inputs(k) for k in (11, 12, 13))

    def step(fn, s):
        code = s["code"].detach().clone().requires_grad_(True)
        code_pos = s["code_pos"].detach().clone().requires_grad_(True)
        out = fn(s["feats"], s["feats_pos"], None, None, code, code_pos, s["depth"], s["depth_pos"])
        (out[0] + out[2] + out[4].mean() + out[6]).backward()
        torch.cuda.synchronize()
        return ([float(out[i].mean()) for i in (0, 2, 4, 6)], code.grad.cpu().numpy(), code_pos.grad.cpu().numpy(),
                fn.last_coords.cpu().numpy().copy(), fn.last_perms.cpu().clone())

    torch.manual_seed(3)
    fn = ContrastiveCorrelationLoss(cfg)
    if how == "side_stream":
        step(fn, s1)
        fn.prefetch_sampling(s2["depth"], s2["depth_pos"], (28, 28))
    else:
        fn.queue_next_sampling(s2["depth"], s2["depth_pos"])
        step(fn, s1)
    assert fn._presampled is not None
    got = step(fn, s2)
    assert fn.last_used_presampled and fn._presampled is None

    def plain(s, perms):
        ref = ContrastiveCorrelationLoss(cfg)
        it = iter(perms.to(dev()))
        ref.perm_fn = lambda B, device: next(it).clone()
        r = step(ref, s)
        assert not ref.last_used_presampled
        return r

    want = plain(s2, got[4])
    assert np.array_equal(got[3], want[3])                      # coordinates: bit-identical
    assert np.allclose(got[0], want[0], rtol=1e-6, atol=1e-7)
    assert rel_err(got[1], want[1]) < 1e-6 and rel_err(got[2], want[2]) < 1e-6
    # stale sampling: queued for s3, but the next forward sees s2 again -> ignored, plain result
    fn.queue_next_sampling(s3["depth"], s3["depth_pos"])
    step(fn, s1)
    again = step(fn, s2)
    assert not fn.last_used_presampled
    want2 = plain(s2, again[4])
    assert np.array_equal(again[3], want2[3]) and rel_err(again[1], want2[1]) < 1e-6


def test_graphed_super_perms_reproduce_the_eager_torch_stream():
    """The CUDA-graph replay of neg_samples x torch.randperm must give the eager calls' permutations
    (same seed -> same stream), call after call, and advance the generator identically."""
    from depthg_b200.modules import _GraphedSuperPerms, super_perms
    _GraphedSuperPerms._cache.clear()
    torch.manual_seed(1234)
    eager = [super_perms(5, 32, dev()).clone() for _ in range(4)]
    tail_eager = torch.rand(3, device=dev())
    torch.manual_seed(1234)
    graphed = [_GraphedSuperPerms.draw(5, 32, dev()) for _ in range(4)]
    tail_graphed = torch.rand(3, device=dev())
    assert _GraphedSuperPerms._cache[(5, 32, dev().index)] is not False, "graph capture failed"
    for a, b in zip(eager, graphed):
        assert torch.equal(a, b)
    assert torch.equal(tail_eager, tail_graphed)


# ------------------------------------------------------------------ DepthContrastiveCorrelationLoss (next row)
@pytest.mark.parametrize("name", list(cases.AUG_CASES))
@pytest.mark.parametrize("variant", ["nchw", "channels_last", "simt"])
def test_depth_contrastive_variant_matches_reference_golden(name, variant):
    from tests.helpers import run_cuda_aug
    g = golden("aug_" + name)
    r = run_cuda_aug(name, channels_last=(variant == "channels_last"), force_simt=(variant == "simt"))
    np.testing.assert_allclose(r["scalars"], g["scalars"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(r["cd_means"], g["cd_means"], rtol=RTOL, atol=ATOL)
    assert rel_err(r["d_code"], g["d_code"]) < RTOL and rel_err(r["d_code_pos"], g["d_code_pos"]) < RTOL


# ------------------------------------------------------------------ non-default sampling modes
@pytest.mark.parametrize("name", list(cases.SAL_CASES))
@pytest.mark.parametrize("channels_last", [False, True])
def test_salience_sampling_matches_reference_golden(name, channels_last):
    """cfg.use_salience (src/modules.py:1291-1298 + sample_nonzero_locations :1191-1204): coordinates bit-equal to the
    real reference's under the same draws, loss and gradients within 1e-4."""
    from depthg_b200.modules import ContrastiveCorrelationLoss
    from tests.helpers import run_sal
    g = golden("sal_" + name)
    r = run_sal(name, ContrastiveCorrelationLoss, device="cuda:0", channels_last=channels_last)
    assert np.array_equal(r["coords1"], g["coords1"]) and np.array_equal(r["coords2"], g["coords2"])
    np.testing.assert_allclose(r["scalars"], g["scalars"], rtol=RTOL, atol=ATOL, equal_nan=True)
    np.testing.assert_allclose(r["cd_means"], g["cd_means"], rtol=RTOL, atol=ATOL, equal_nan=True)
    assert rel_err(r["d_code"], g["d_code"]) < RTOL and rel_err(r["d_code_pos"], g["d_code_pos"]) < RTOL


def test_salience_default_draws_follow_the_reference_call_order():
    """Without hooks: randint on the CPU generator per image (device generator for an all-zero map), like :1197-1199."""
    from depthg_b200.modules import sample_nonzero_locations
    sal = (torch.rand(3, 28, 28) > 0.5).float()
    sal[2] = 0
    torch.manual_seed(5)
    got = sample_nonzero_locations(sal.to(dev()), [3, 4, 4, 2])
    torch.manual_seed(5)
    want = O.sample_nonzero_locations(sal, [3, 4, 4, 2])
    assert torch.equal(got[:2].cpu(), want[:2])          # CPU-generator draws: same stream on both sides
    assert got[2].min() >= -1 and got[2].max() < 1       # all-zero map: device draw, only the range is comparable


def test_fps_depth_feat_mode_equals_fps_mode():
    """depth_sampling='fps_depth_feat' passes include_feats=True, which the reference function ignores (:999-1037)."""
    from tests.gpu_helpers import run_cuda_loss
    cfg, t = cases.make_loss_inputs("small_fps")
    _, _, a = run_cuda_loss("small_fps", inputs=(cfg, t))
    cfg.depth_sampling = "fps_depth_feat"
    _, _, b = run_cuda_loss("small_fps", inputs=(cfg, t))
    assert np.array_equal(a["coords1"], b["coords1"]) and np.array_equal(a["scalars"], b["scalars"])
    assert rel_err(a["d_code"], b["d_code"]) < 1e-6      # (the scatter backward adds with atomics: not bit-reproducible)
    _check_loss(b, golden("loss_small_fps"), cfg)


# ------------------------------------------------------------------ edge shapes against the oracle (computed on the fly)
@pytest.mark.parametrize("shape", [
    dict(B=1, C=32, D=16, S=4, neg=0, sampling="fps"),          # single image, no negatives
    dict(B=2, C=100, D=128, S=5, neg=2, sampling="none"),       # C not a multiple of 32/128, code dim at the 128 limit
    dict(B=3, C=256, D=33, S=16, neg=1, sampling="fps"),        # S*S = 256: the largest tensor-core tiling
    dict(B=2, C=20, D=7, S=3, neg=3, sampling="fps", H=20, W=12, Hd=80, Wd=60),   # non-square grid, odd pooling windows
    dict(B=2, C=64, D=24, S=20, neg=2, sampling="none"),        # 400 points: tcgen05 column groups (4 row tiles x 2 groups)
    dict(B=2, C=32, D=16, S=28, neg=1, sampling="fps"),         # the dense 28x28 stress shape: 784 points, 7 x 4 blocks
    dict(B=2, C=128, D=40, S=17, neg=1, sampling="none"),       # 289 points: just above the single-group limit
])
def test_edge_shapes_match_oracle(shape):
    from depthg_b200.modules import ContrastiveCorrelationLoss
    H, W = shape.get("H", 28), shape.get("W", 28)
    Hd, Wd = shape.get("Hd", 8 * H), shape.get("Wd", 8 * W)
    B, C, D, S, neg = shape["B"], shape["C"], shape["D"], shape["S"], shape["neg"]
    rs = np.random.RandomState(cases.hash_name(str(sorted(shape.items()))) % (2 ** 31))   # stable across processes
    t = dict(feats=cases.correlated(rs, B, C, H, W), feats_pos=cases.correlated(rs, B, C, H, W),
             code=cases.correlated(rs, B, D, H, W, rank=3), code_pos=cases.correlated(rs, B, D, H, W, rank=3),
             depth=rs.randint(0, 256, (B, 1, Hd, Wd)).astype(np.float32),
             depth_pos=rs.randint(0, 256, (B, 1, Hd, Wd)).astype(np.float32))
    t = {k: torch.from_numpy(v) for k, v in t.items()}
    perms = [torch.from_numpy(cases.bumped_perm(rs, B)) for _ in range(neg)] if B > 1 else []
    rands = [torch.from_numpy(rs.random_sample((B, S, S, 2)).astype(np.float32)) for _ in range(2)]
    cfg = cases.loss_cfg(feature_samples=S, neg_samples=neg, depth_sampling=shape["sampling"])
    res = []
    for impl, devc in ((O.ContrastiveCorrelationLoss, "cpu"), (ContrastiveCorrelationLoss, "cuda:0")):
        a = {k: v.to(devc) for k, v in t.items()}
        code, code_pos = a["code"].clone().requires_grad_(True), a["code_pos"].clone().requires_grad_(True)
        fn = impl(cfg)
        pit, rit = iter([p.to(devc) for p in perms]), iter([r.to(devc) for r in rands])
        fn.perm_fn = lambda n, device: next(pit).clone()
        fn.rand_fn = lambda shp, device: next(rit).clone()
        if neg == 0:   # the reference cannot cat an empty list of negatives; compare the positive terms only
            if impl is O.ContrastiveCorrelationLoss:
                cfg1 = cases.loss_cfg(feature_samples=S, neg_samples=1, depth_sampling=shape["sampling"])
                fn = impl(cfg1)
                fn.perm_fn = lambda n, device: torch.zeros(n, dtype=torch.long)
                fn.rand_fn = lambda shp, device: next(rit).clone()
            out = fn(a["feats"], a["feats_pos"], None, None, code, code_pos, a["depth"], a["depth_pos"])
            L = out[0] * 0.3 + out[2] * 0.7 + out[6] * 0.2
        else:
            out = fn(a["feats"], a["feats_pos"], None, None, code, code_pos, a["depth"], a["depth_pos"])
            L = out[0] * 0.3 + out[2] * 0.7 + out[4].mean() * 0.5 + out[6] * 0.2
        L.backward()
        res.append((float(L), [float(out[i].mean()) for i in (0, 2, 6)], code.grad.cpu().numpy(), code_pos.grad.cpu().numpy()))
    (L0, s0, g0, gp0), (L1, s1, g1, gp1) = res
    np.testing.assert_allclose(s1, s0, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(L1, L0, rtol=RTOL, atol=ATOL)
    # dense shapes: hundreds of thousands of cd entries per pair sit within fp32 rounding of the zero clamp, and one
    # indicator that differs between two fp32 evaluation orders moves the gradient by ~1e-4 (same effect, and same
    # treatment, as test_cfg2_full_size_matches_oracle); the golden-vector tests of these shapes hold 1e-4
    tol = RTOL if S * S <= 256 else 3e-4
    assert rel_err(g1, g0) < tol and rel_err(gp1, gp0) < tol


# ------------------------------------------------------------------ probe losses
@pytest.mark.parametrize("layout", ["nchw", "channels_last"])
@pytest.mark.parametrize("name", list(cases.PROBE_CASES))
def test_probe_losses_match_reference_golden(name, layout):
    from depthg_b200.probes import ClusterLookup, linear_probe_loss
    g = golden("probes")
    t = {k: v.to(dev()) for k, v in cases.make_probe_inputs(name).items()}
    code = t["code"].contiguous(memory_format=torch.channels_last) if layout == "channels_last" else t["code"]
    K, D = t["weight"].shape
    probe = ClusterLookup(D, K).to(dev())
    with torch.no_grad():
        probe.clusters.copy_(t["clusters"])
    loss, probs = probe(code, None)
    (loss * 3.0).backward()
    np.testing.assert_allclose(loss.item(), g[name + "_cluster_loss"], rtol=RTOL)
    p = probs.cpu().numpy()
    assert p.shape == (code.shape[0], K) + tuple(code.shape[2:])
    assert ((p == 0) | (p == 1)).all() and (p.sum(1) == 1).all()
    # assignments may only differ where the two best similarities tie to fp32 rounding
    diff = p.argmax(1) != g[name + "_cluster_argmax"]
    assert diff.mean() < 1e-3
    assert rel_err(probe.clusters.grad.cpu().numpy() / 3.0, g[name + "_cluster_grad"]) < (1e-3 if diff.any() else RTOL)

    weight = t["weight"].reshape(K, D, 1, 1).clone().requires_grad_(True)
    bias = t["bias"].clone().requires_grad_(True)
    lin = linear_probe_loss(code, weight, bias, t["label"])
    (lin * 0.5).backward()
    np.testing.assert_allclose(lin.item(), g[name + "_linear_loss"], rtol=RTOL)
    assert weight.grad.shape == weight.shape
    assert rel_err(weight.grad.reshape(K, D).cpu().numpy() * 2.0, g[name + "_linear_dw"]) < RTOL
    assert rel_err(bias.grad.cpu().numpy() * 2.0, g[name + "_linear_db"]) < RTOL

    with torch.no_grad():
        sl, sp = probe(code, 2)
        np.testing.assert_allclose(sl.item(), g[name + "_soft_loss"], rtol=RTOL, atol=ATOL)
        if name + "_soft_probs" in g:
            np.testing.assert_allclose(sp.cpu().numpy(), g[name + "_soft_probs"], rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(probe(code, 2, log_probs=True).cpu().numpy(), g[name + "_log_probs"], rtol=RTOL,
                                       atol=1e-5)


def test_probes_refuse_attached_code_and_all_masked_labels_give_nan():
    from depthg_b200.probes import ClusterLookup, linear_probe_loss
    t = {k: v.to(dev()) for k, v in cases.make_probe_inputs("probe_small").items()}
    K, D = t["weight"].shape
    with pytest.raises(ValueError, match="requires grad"):
        linear_probe_loss(t["code"].clone().requires_grad_(True), t["weight"], t["bias"], t["label"])
    with pytest.raises(ValueError, match="requires grad"):
        ClusterLookup(D, K).to(dev())(t["code"].clone().requires_grad_(True), None)
    with pytest.raises(NotImplementedError):
        ClusterLookup(D, K).to(dev())(t["code"], 2.0)
    nothing = torch.full_like(t["label"], -1)
    assert torch.isnan(linear_probe_loss(t["code"], t["weight"], t["bias"], nothing))   # mean over no pixels, as torch
    # forward-only call (no parameter requires grad) and a bias-free probe
    a = linear_probe_loss(t["code"], t["weight"], None, t["label"])
    b = O.linear_probe_loss(t["code"].cpu(), t["weight"].cpu(), None, t["label"].cpu())
    np.testing.assert_allclose(a.item(), b.item(), rtol=RTOL)


def test_linear_probe_full_size_matches_oracle():
    """cfg2 shapes at the bench batch (B=32, 224x224 labels): loss and gradients against the CPU oracle."""
    from depthg_b200.probes import linear_probe_loss
    rs = np.random.RandomState(99)
    B, D, K = 32, 90, 27
    code = torch.from_numpy(cases.correlated(rs, B, D, 28, 28, rank=4))
    weight = torch.from_numpy((rs.standard_normal((K, D)) / np.sqrt(D)).astype(np.float32))
    bias = torch.from_numpy((rs.standard_normal((K,)) * 0.1).astype(np.float32))
    label = torch.from_numpy(rs.randint(-1, K, (B, 224, 224)).astype(np.int64))
    w0, b0 = weight.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    O.linear_probe_loss(code, w0, b0, label).backward()
    w1, b1 = weight.to(dev()).requires_grad_(True), bias.to(dev()).requires_grad_(True)
    lin = linear_probe_loss(code.to(dev()), w1, b1, label.to(dev()))
    lin.backward()
    assert rel_err(w1.grad.cpu().numpy(), w0.grad.numpy()) < RTOL
    assert rel_err(b1.grad.cpu().numpy(), b0.grad.numpy()) < RTOL


def test_dense_shapes_run_on_the_tensor_core_kernel():
    """S*S > 128 must go through a tcgen05 kernel (above 256 points: column groups + row_means_kernel), not the
    CUDA-core fallback; either tensor-core kernel can be forced with DEPTHG_B200_CORR=pipe|umma1."""
    import ctypes
    from depthg_b200 import _lib
    from depthg_b200.modules import ContrastiveCorrelationLoss, corr_kernel_choice
    assert corr_kernel_choice(784, 90) == "umma" and corr_kernel_choice(1025, 90) == "simt"
    g = torch.Generator(device=dev()).manual_seed(3)
    B, C, D, S = 2, 64, 32, 18
    f = lambda ch: torch.randn((B, ch, 28, 28), generator=g, device=dev())   # noqa: E731
    code, code_pos = f(D).requires_grad_(True), f(D).requires_grad_(True)
    fn = ContrastiveCorrelationLoss(cases.loss_cfg(feature_samples=S, neg_samples=2, depth_sampling="none"))
    lib = _lib.lib()
    lib.dg_profile_enable(1)
    depth = torch.rand((B, 1, 224, 224), generator=g, device=dev()) * 255
    out = fn(f(C), f(C), None, None, code, code_pos, depth, depth)
    (out[0] + out[2] + out[4].mean() + out[6]).backward()
    torch.cuda.synchronize()
    n = lib.dg_profile_collect(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    lib.dg_profile_collect(buf, n + 16)
    lib.dg_profile_enable(0)
    names = buf.value.decode()
    assert ("corr_umma_kernel" in names or "corr_pipe_kernel" in names) and "row_means_kernel" in names
    assert "corr_tile_kernel" not in names
    assert torch.isfinite(code.grad).all() and torch.isfinite(code_pos.grad).all()
