"""CPU tests: the C-ABI library loads and exports every symbol the header
declares, argument validation works without a GPU, the Python host layer refuses
CPU tensors (no fallback), and the sharding helpers are right (gloo, world 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "depthg_b200.h")).read()
    return sorted(set(re.findall(r"DG_API\s+[\w\s\*]+?\b(dg_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from depthg_b200 import _lib
    names = header_symbols()
    assert len(names) >= 13
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), n
    assert lib.dg_version() >= 100
    assert lib.dg_panel_ld(90) == 96 and lib.dg_panel_ld(768) == 768 and lib.dg_panel_rows(121) == 128
    assert lib.dg_corr_loss_workspace_bytes(7, 32, 121) > 0


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "depthg_b200.h"\nint main(void){return DG_OK + (int)sizeof(dg_stream_t)*0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_argument_validation_needs_no_gpu():
    from depthg_b200 import _lib
    lib = _lib.lib()
    assert lib.dg_fps_coords(None, None, 1, 8, 8, 4, 4, 2, 1.0, 5.0, 0, None, None, None) == -1
    assert b"null pointer" in lib.dg_last_error_string()
    assert lib.dg_knn_topk(None, None, 4, 4, 4, 2, None, None, None, 0, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(_lib.DepthgB200Error):
        _lib.check(-4, "x")


def test_host_layer_refuses_cpu_tensors_and_bad_modes():
    """No CPU fallback by design (SURVEY.md 8b error conventions)."""
    from depthg_b200 import modules as M
    from depthg_b200.precompute_knns import knn_topk
    from tests.golden import cases
    cfg = cases.loss_cfg(feature_samples=3)
    fn = M.ContrastiveCorrelationLoss(cfg)
    assert len(fn.state_dict()) == 0
    f = torch.zeros(2, 8, 28, 28)
    d = torch.zeros(2, 1, 224, 224)
    with pytest.raises(ValueError, match="CUDA"):
        fn(f, f, None, None, f, f, d, d)
    with pytest.raises(ValueError, match="CUDA"):
        M.farthest_point_sampling_depth(f, d, 3)
    with pytest.raises(ValueError, match="CUDA"):
        knn_topk(torch.zeros(4, 8), torch.zeros(4, 8), 2)
    with pytest.raises(TypeError):
        M._lib.require_cuda_f32(np.zeros(3), "x")


def test_probe_host_layer_validates_without_a_gpu():
    """probes.py: no CPU fallback, the reference's detach contract is enforced, plan/workspace sizes are sane."""
    from depthg_b200 import _lib
    from depthg_b200.probes import ClusterLookup, linear_probe_loss
    code = torch.zeros(2, 16, 7, 7)
    with pytest.raises(ValueError, match="CUDA"):
        linear_probe_loss(code, torch.zeros(5, 16, 1, 1), torch.zeros(5), torch.zeros(2, 56, 56, dtype=torch.int64))
    probe = ClusterLookup(16, 5)
    assert tuple(probe.clusters.shape) == (5, 16) and list(probe.state_dict()) == ["clusters"]
    with pytest.raises(ValueError, match="CUDA"):
        probe(code, None)
    with pytest.raises(ValueError, match="not supported"):
        ClusterLookup(16, 33)
    lib = _lib.lib()
    assert lib.dg_probe_workspace_bytes(0, 7, 7, 16, 5) == 0
    small, big = lib.dg_probe_workspace_bytes(2, 7, 7, 16, 5), lib.dg_probe_workspace_bytes(32, 28, 28, 90, 27)
    assert 0 < small < big and big >= 2 * 32 * 28 * 28 * 32 * 4
    assert lib.dg_linear_probe_ce(None, None, 2, 16, 7, 7, None, None, 5, None, None, 56, 56, None, None, None, None, 0,
                                  None) == -1
    assert b"null pointer" in lib.dg_last_error_string()
    assert lib.dg_cluster_probe(None, None, 2, 16, 7, 7, None, 5, 0, 0.0, None, None, None, None, 0, None) == -1


def test_loss_plan_sizes_for_sampled_and_dense_shapes():
    """dg_loss_plan is host-only: panel row padding and partial-buffer counts of the tcgen05 path."""
    import ctypes as C
    from depthg_b200 import _lib
    lib = _lib.lib()

    def plan(S, flags=1 | 2, B=4, Cdim=64, D=24):
        desc = _lib.LossDesc(B, Cdim, D, 28, 28, 224, 224, S, 2, flags, 0.1, 0.2, 0.3, 0.0)
        p = _lib.LossPlan()
        _lib.check(lib.dg_loss_plan(C.byref(desc), C.byref(p)), "dg_loss_plan")
        return p

    import os
    p11, p12, q17, q28 = plan(11), plan(12), plan(17), plan(28)
    # default dispatch: kernel 2 (the persistent tcgen05 kernel) up to 128 points, kernel 1 (2 x 2 tile blocks up to 256
    # points, column groups of 256 above: 289 -> 512, 784 -> 1024) beyond
    assert (p11.kernel, p11.Prows) == (2, 128) and (p12.kernel, p12.Prows) == (1, 256)
    assert (q17.kernel, q17.Prows) == (1, 512) and (q28.kernel, q28.Prows) == (1, 1024)
    os.environ["DEPTHG_B200_CORR"] = "pipe"           # the persistent kernel forced: whole 128-row tiles (289 -> 384, 784 -> 896)
    try:
        p17, p28 = plan(17), plan(28)
    finally:
        del os.environ["DEPTHG_B200_CORR"]
    assert (p17.kernel, p17.Prows) == (2, 384) and (p28.kernel, p28.Prows) == (2, 896)
    assert p11.total < p12.total < p17.total < p28.total
    forced = plan(11, flags=1 | 2 | _lib.FLAG_FORCE_SIMT)
    assert (forced.kernel, forced.Prows) == (0, 128)
    desc = _lib.LossDesc(4, 64, 24, 28, 28, 224, 224, 30, 2, 3, 0.1, 0.2, 0.3, 0.0)   # 900 points > 784 grid points
    assert lib.dg_loss_plan(C.byref(desc), C.byref(_lib.LossPlan())) == -1


def test_missing_library_fails_loudly(monkeypatch):
    from depthg_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdepthg_b200.so")
    with pytest.raises(_lib.DepthgB200Error, match="no CPU"):
        _lib.lib()


def test_shard_bounds_cover_everything():
    from depthg_b200.distributed import shard_bounds
    for n in (1, 7, 64, 49629):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from depthg_b200.distributed import shard_bounds, shard_batch, allreduce_mean_, sharded_knn, allgather_rows, knn_shard_bounds
from oracle import depthg_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rs = np.random.RandomState(0)
feats = torch.nn.functional.normalize(torch.from_numpy(rs.standard_normal((301, 32)).astype(np.float32)), dim=1)
lo, hi = shard_bounds(301, world, rank)
cpu_topk = lambda q, db, k: O.knn_rows(q, db, k)[1]
full = sharded_knn(feats[lo:hi], 301, 5, cpu_topk, gather_result=True)
assert torch.equal(full, O.knn_rows(feats, feats, 5)[1]), "sharded KNN != single-process KNN"
# the KNN build's ceil partition: the equal-size all-gather of (end-padded) shards IS the database
klo, khi = knn_shard_bounds(301, world, rank)
per = -(-301 // world)
src = torch.zeros((per, 32)); src[:khi - klo] = feats[klo:khi]
gathered = torch.empty((world * per, 32))
dist.all_gather_into_tensor(gathered, src)
assert torch.equal(gathered[:301], feats), "ceil-partition all-gather is not the database"
even = allgather_rows(feats[rank * 100:(rank + 1) * 100], 200)
assert torch.equal(even, feats[:200])
g = [torch.full((3, 2), float(rank + 1)), None, torch.arange(4.0) * (rank + 1)]
allreduce_mean_(g)
assert torch.allclose(g[0], torch.full((3, 2), 1.5)) and torch.allclose(g[2], torch.arange(4.0) * 1.5)
x = torch.arange(10).view(5, 2)
(part,) = shard_batch([x], world, rank)
assert part.shape[0] == (3 if rank == 0 else 2)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_knn_shard_bounds_cover_rows_with_short_last_shards():
    from depthg_b200.distributed import knn_shard_bounds
    for n, world in ((49629, 8), (301, 2), (5, 8), (64, 8), (1, 3)):
        per = -(-n // world)
        spans = [knn_shard_bounds(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo == min(r * per, n) and hi - lo <= per for r, (lo, hi) in enumerate(spans))


def test_distributed_sharding_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out


def test_nns_file_contract_round_trips_to_the_dataset_loader(tmp_path):
    """what save_nns writes is what ContrastiveSegDataset reads
    (/root/reference/src/precompute_knns.py:71-72,115 -> /root/reference/src/data.py:1056-1064,1079)."""
    from depthg_b200.precompute_knns import PRECOMPUTE_RES, load_nns, nns_filename, save_nns
    from oracle import depthg_oracle as O
    from tests.golden import cases
    feats, k, nb = cases.make_knn_feats("ragged_1001")
    nn = O.knn_topk_chunked(feats, 30, nb)                      # [1001, 30] int64, as the script builds it
    # producer file name: the script's format string with its hard-coded res
    name = nns_filename("vit_base", "cocostuff27", "train", "five", PRECOMPUTE_RES)
    assert name == "nns_{}_{}_{}_{}_{}.npz".format("vit_base", "cocostuff27", "train", "five", 392)
    assert name == "nns_vit_base_cocostuff27_train_five_392.npz"
    # crop_type None formats as the string "None" on both sides (the script's `crop_types = [None]` variant)
    assert nns_filename("vit_small", "directory_name", "val", None, 224) == "nns_vit_small_directory_name_val_None_224.npz"
    path = str(tmp_path / name)
    save_nns(path, nn)
    # consumer (src/data.py:1062-1064): np.load(file)["nns"], one row per dataset item
    loaded = np.load(path)
    assert loaded.files == ["nns"]
    nns = loaded["nns"]
    assert nns.dtype == np.int64 and nns.shape == (1001, 30)
    assert np.array_equal(nns, nn.numpy()) and np.array_equal(load_nns(path), nns)
    # src/data.py:1079: ind_pos = nns[ind][randint(1, num_neighbors + 1)] - columns 1..7 at cfg.num_neighbors = 7 are
    # valid dataset indices, and column 0 is the image itself
    num_neighbors = 7
    assert (nns[:, 0] == np.arange(1001)).all()
    picks = nns[:, 1:num_neighbors + 1]
    assert picks.min() >= 0 and picks.max() < 1001 and (picks != np.arange(1001)[:, None]).all()
    with pytest.raises(ValueError):
        save_nns(path, nn[0])
    # int32 / CUDA-produced tensors are widened to the int64 the reference stores
    save_nns(path, nn.to(torch.int32))
    assert np.load(path)["nns"].dtype == np.int64


def test_fps_round_loop_has_no_fused_multiply_add():
    """The FPS distances must round like NumPy's (a-b)**2 summed left to right: one rounding per operation.  ptxas
    contracts packed mul.rn.f32x2 + add.rn.f32x2 into FFMA2 whatever -fmad says, so the kernel keeps its additions
    scalar; this checks the SASS of every fps_kernel instance: between the packed subtractions of a round and the
    redux that ends it there is no fp32 fused multiply-add (FFMA / FFMA2)."""
    import shutil
    obj = os.path.join(ROOT, "depthg_b200", "csrc", "build", "fps.o")
    if shutil.which("cuobjdump") is None or not os.path.isfile(obj):
        pytest.skip("needs cuobjdump and the object file of the in-tree build")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    checked = 0
    for f in funcs:
        name = f.split("\n", 1)[0]
        if "fps_kernel" not in name:
            continue
        ops = [m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", f)]
        # a kernel holds one copy of the round loop per staging variant: every stretch from a packed subtraction to
        # the redux that follows it is distance arithmetic
        starts = [i for i, o in enumerate(ops) if o.startswith("FADD2")]
        assert starts, name
        for i in starts:
            j = next(k for k in range(i, len(ops)) if ops[k].startswith("CREDUX"))
            assert not [o for o in ops[i:j] if o.startswith("FFMA")], name
        assert any(o.startswith("FMUL2") for o in ops), name
        checked += 1
    assert checked >= 4


def test_ctypes_mirrors_have_the_c_struct_sizes(tmp_path):
    """The ctypes Structures of depthg_b200/_lib.py must lay out like the C structs of include/depthg_b200.h (a field
    added on one side only would shift every later field of the debug binding silently)."""
    from depthg_b200 import _lib
    import ctypes as C
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "depthg_b200.h"\nint main(void) { printf("%zu %zu %zu %zu\\n", '
                   'sizeof(dg_loss_desc_t), sizeof(dg_loss_plan_t), sizeof(dg_loss_io_t), sizeof(dg_loss_grads_t)); '
                   'return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.LossDesc), C.sizeof(_lib.LossPlan), C.sizeof(_lib.LossIO), C.sizeof(_lib.LossGrads)]
    assert got == want, (got, want)
