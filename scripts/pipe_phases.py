import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, numpy as np, ctypes
from depthg_b200 import modules as M, _lib
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
BB = int(sys.argv[1]) if len(sys.argv) > 1 else 32
s = bench.synth_inputs(BB, gen, dev)
s["code"].requires_grad_(True); s["code_pos"].requires_grad_(True)
fn = M.ContrastiveCorrelationLoss(bench.make_cfg())
def fwd():
    return fn(s["feats"], s["feats_pos"], None, None, s["code"], s["code_pos"], s["depth"], s["depth_pos"])
for _ in range(5): out = fwd()
print("losses", [float(o) for o in out])
clk = torch.zeros((224, 16), dtype=torch.int64, device=dev)
_lib.lib().dg_debug_set_clock_buffer(clk.data_ptr())
fwd(); torch.cuda.synchronize()
_lib.lib().dg_debug_set_clock_buffer(None)
c = clk.cpu().numpy().astype(np.float64)
live = c[:, 2] > 0
t0 = c[live][:, 1].min()
names = {1: "epi:begin", 2: "acc_full", 3: "rowmean", 4: "U stored", 5: "grad_full", 6: "drained", 7: "acc_full#2", 8: "drained#2", 9: "end"}
for cta in (0, 40, 75, 76, 100, 147):
    row = c[cta]
    print(f"CTA {cta:3d}: " + "  ".join(f"{names[i]}@{(row[i]-t0)/1e3:6.2f}" for i in range(1, 10) if row[i] > 0))
print("kernel span us:", (c[live][:, 9].max() - t0) / 1e3)
lib = _lib.lib(); lib.dg_profile_enable(1)
for _ in range(10): fwd()
n = lib.dg_profile_collect(None, 0); buf = ctypes.create_string_buffer(n + 16); lib.dg_profile_collect(buf, n + 16)
lib.dg_profile_enable(0)
for ln in buf.value.decode().strip().split("\n"):
    nm, k, us = ln.split("\t"); print(f"{nm:28s} {float(us)/int(k):8.2f} us")
