import sys, os, cProfile, pstats
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
    bench.weighted(out).backward()
for _ in range(30): full()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(300): full()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
