"""Device-resident fwd+bwd timings of the other BASELINE configs (not the headline bench line)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from types import SimpleNamespace
from depthg_b200.modules import ContrastiveCorrelationLoss

dev = torch.device("cuda:0")
def run(name, B, C, D, S, H=28, sampling="fps", pointwise=True, nchw=False, steps=50, sampler="torch"):
    cfg = bench.make_cfg(S); cfg.depth_sampling = sampling; cfg.pointwise = pointwise
    g = torch.Generator(device=dev).manual_seed(1)
    def feat(ch):
        x = torch.randn((B, H, H, ch), generator=g, device=dev).permute(0, 3, 1, 2)
        return x.contiguous() if nchw else x
    f, fp, c, cp = feat(C), feat(C), feat(D).requires_grad_(True), feat(D).requires_grad_(True)
    d = torch.randint(0, 256, (B, 1, 8 * H, 8 * H), generator=g, device=dev).float()
    dp = torch.randint(0, 256, (B, 1, 8 * H, 8 * H), generator=g, device=dev).float()
    fn = ContrastiveCorrelationLoss(cfg, negative_sampler=sampler)
    def step():
        c.grad = None; cp.grad = None
        bench.backprop(fn(f, fp, None, None, c, cp, d, dp))
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": name, "B": B, "C": C, "D": D, "S": S, "grid": H, "sampling": sampling, "layout": "nchw" if nchw else "channels_last",
                      "ms_per_step": round(ms, 4), "samples_per_s": round(B / ms * 1e3)}), flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "dense":
    _skip = True
else:
    _skip = False
_run = run
def run(*a, **k):
    if _skip and "dense" not in a[0]:
        return
    _run(*a, **k)
run("cfg1 ViT-S", 2, 384, 70, 11)
run("cfg2 ViT-B S=11", 32, 768, 90, 11)
run("cfg2 ViT-B S=11 nchw", 32, 768, 90, 11, nchw=True)
run("cfg2' ViT-B S=12", 32, 768, 90, 12)
run("cfg4 Cityscapes", 64, 768, 100, 11, sampling="none", pointwise=False)
import ctypes
from depthg_b200 import _lib
only_dense = len(sys.argv) > 1 and sys.argv[1] == "dense"
def profile_once(fn):
    lib = _lib.lib(); lib.dg_profile_enable(1); fn(); torch.cuda.synchronize()
    n = lib.dg_profile_collect(None, 0); buf = ctypes.create_string_buffer(n + 16); lib.dg_profile_collect(buf, n + 16)
    lib.dg_profile_enable(0); print(buf.value.decode())
run("cfg5 dense 28x28 B=4 (random coords)", 4, 768, 90, 28, sampling="none", steps=10)
run("cfg5 dense 28x28 B=64 (random coords)", 64, 768, 90, 28, sampling="none", steps=5)
os.environ["DEPTHG_B200_CORR"] = "simt"
run("cfg5 dense 28x28 B=4 (generic CUDA-core kernel)", 4, 768, 90, 28, sampling="none", steps=3)
del os.environ["DEPTHG_B200_CORR"]
