"""Per-kernel breakdown of the dense 28x28 stress config (BASELINE configs[4]: B=64, C=768, dim 90, 784x784 per pair)."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from depthg_b200 import _lib
from depthg_b200.modules import ContrastiveCorrelationLoss

dev = torch.device("cuda:0")
B, C, D, S, H = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 768, 90, 28, 28
cfg = bench.make_cfg(S); cfg.depth_sampling = "none"
g = torch.Generator(device=dev).manual_seed(1)
feat = lambda ch: torch.randn((B, H, H, ch), generator=g, device=dev).permute(0, 3, 1, 2)
f, fp, c, cp = feat(C), feat(C), feat(D).requires_grad_(True), feat(D).requires_grad_(True)
d = torch.randint(0, 256, (B, 1, 224, 224), generator=g, device=dev).float()
fn = ContrastiveCorrelationLoss(cfg)
def step():
    c.grad = None; cp.grad = None
    bench.backprop(fn(f, fp, None, None, c, cp, d, d))
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
lib = _lib.lib(); lib.dg_profile_enable(1); step(); torch.cuda.synchronize()
n = lib.dg_profile_collect(None, 0); buf = ctypes.create_string_buffer(n + 16); lib.dg_profile_collect(buf, n + 16)
lib.dg_profile_enable(0)
br = {ln.split("\t")[0]: float(ln.split("\t")[2]) for ln in buf.value.decode().strip().split("\n") if ln}
P = S * S
flops = 7 * B * P * P * 2 * (C + 3 * 96)          # fd + cd + the two gradient GEMMs, counted once
print(json.dumps({"config": "dense 28x28", "B": B, "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3),
                  "breakdown_us": br, "useful_tflops_corr_kernel": round(flops / (br.get("corr_umma_kernel", 1) * 1e-6) / 1e12, 1),
                  "mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}))
