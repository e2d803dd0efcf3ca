import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, numpy as np
from depthg_b200 import modules as M, _lib
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
s = bench.synth_inputs(32, gen, dev)
s["code"].requires_grad_(True); s["code_pos"].requires_grad_(True)
fn = M.ContrastiveCorrelationLoss(bench.make_cfg())
def fwd():
    return fn(s["feats"], s["feats_pos"], None, None, s["code"], s["code_pos"], s["depth"], s["depth_pos"])
for _ in range(5): fwd()
clk = torch.zeros((224, 16), dtype=torch.int64, device=dev)
_lib.lib().dg_debug_set_clock_buffer(clk.data_ptr())
fwd(); torch.cuda.synchronize()
_lib.lib().dg_debug_set_clock_buffer(None)
c = clk.cpu().numpy().astype(np.float64)
t0 = c[:, 0].min()
names = ["start", "epi:begin", "acc_full", "rowmean done", "U stored", "grad_full r0", "drained r0", "grad_full r1", "drained r1", "end", "dealloc"]
for cta in (0, 40, 100, 147, 148, 200, 223):
    row = c[cta]
    print(f"CTA {cta:3d} k={cta//32} start@{(row[0]-t0)/1e3:7.2f}us  " + "  ".join(f"{names[i]}+{(row[i]-row[0])/1e3:6.2f}" for i in (1,2,3,4,5,6,9,10) if row[i] > 0))
print("kernel span us:", (c[:, 10].max() - t0) / 1e3)
starts = np.sort(c[:, 0] - t0) / 1e3
print("CTA start times us (sorted): first", starts[:3], "148th", starts[147:150], "last", starts[-3:])
print("mean per-CTA duration us", ((c[:, 10] - c[:, 0]) / 1e3).mean())
