import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from depthg_b200 import modules as M
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
s = bench.synth_inputs(32, gen, dev)
s["code"].requires_grad_(True); s["code_pos"].requires_grad_(True)
fn = M.ContrastiveCorrelationLoss(bench.make_cfg())
def full():
    s["code"].grad = None; s["code_pos"].grad = None
    out = fn(s["feats"], s["feats_pos"], None, None, s["code"], s["code_pos"], s["depth"], s["depth_pos"])
    bench.backprop(out)
for _ in range(30): full()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): full()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host {1e6*(t1-t0)/300:.1f} us/iter, incl drain {1e6*(t2-t0)/300:.1f}")
pr = cProfile.Profile(); pr.enable()
for _ in range(300): full()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
