"""Where does the host time of one loss step go?  (run on the GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from depthg_b200 import modules as M

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
s = bench.synth_inputs(32, gen, dev)
s["code"].requires_grad_(True); s["code_pos"].requires_grad_(True)
cfg = bench.make_cfg()
fn = M.ContrastiveCorrelationLoss(cfg)

def timeit(name, f, n=300):
    for _ in range(20): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:44s} host {1e6*(t1-t0)/n:8.1f} us/iter   incl. drain {1e6*(t2-t0)/n:8.1f} us/iter")

timeit("super_perms eager(5,32)", lambda: M.super_perms(5, 32, dev))
timeit("super_perms graphed(5,32)", lambda: M._GraphedSuperPerms.draw(5, 32, dev))
timeit("fused_super_perms(5,32)", lambda: M.fused_super_perms(5, 32, dev))
def fwd():
    return fn(s["feats"], s["feats_pos"], None, None, s["code"], s["code_pos"], s["depth"], s["depth_pos"])
with torch.no_grad():
    timeit("forward (no grad)", fwd)
timeit("forward (grad mode)", fwd)
def full():
    s["code"].grad = None; s["code_pos"].grad = None
    bench.backprop(fwd())
timeit("forward + backprop", full)
fn.negative_sampler = "fused"
timeit("forward + backprop, fused sampler", full)
