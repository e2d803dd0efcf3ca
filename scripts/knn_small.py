import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depthg_b200.precompute_knns import knn_topk
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(N, F, device="cuda"), dim=1)
idx, stats = knn_topk(x, x, 30, return_stats=True)
torch.cuda.synchronize()
want = torch.topk(x @ x.T, 30)[1]
print("stats", stats, "mismatching slots", int((idx != want).sum()))
