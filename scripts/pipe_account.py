"""Where the roles of corr_pipe_kernel wait (clock64 sums per CTA), for cfg2 / S=12 / dense shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, numpy as np
from depthg_b200 import modules as M, _lib
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
gen = torch.Generator(device=dev).manual_seed(0)
s = bench.synth_inputs(B, gen, dev)
fn = M.ContrastiveCorrelationLoss(bench.make_cfg(S))
def fwd():
    with torch.no_grad():
        return fn(s["feats"], s["feats_pos"], None, None, s["code"], s["code_pos"], s["depth"], s["depth_pos"])
for _ in range(3): fwd()
clk = torch.zeros((148, 16), dtype=torch.int64, device=dev)
_lib.lib().dg_debug_set_clock_buffer(clk.data_ptr())
fwd(); torch.cuda.synchronize()
_lib.lib().dg_debug_set_clock_buffer(None)
c = clk.cpu().numpy().astype(np.float64)
live = c[:, 15] > 0
tot = c[live, 15]
f = lambda i: (c[live, i] / tot).mean()
print(f"B={B} S={S}: items/CTA {c[live,0].mean():.1f}  kernel cycles/CTA {tot.mean():.0f} ({tot.mean()/1.9e3:.1f} us @1.9GHz)  per item {tot.mean()/max(c[live,0].mean(),1)/1.9e3:.2f} us")
print(f"  MMA thread waits: operand stage {f(10):.2f}  TMEM set free {f(11):.2f}  end-of-item gradient steps {f(12):.2f}")
print(f"  MMA thread busy:  issuing chunk MMAs {f(3):.2f}  commits {f(4):.2f}  polling {f(5):.2f}  gradient steps in-stream {f(6):.2f}")
print(f"  epilogue waits:   accumulators {f(13):.2f}  gradient GEMMs {f(14):.2f}")
