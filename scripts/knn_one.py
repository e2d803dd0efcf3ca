import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depthg_b200.precompute_knns import knn_topk
N = int(sys.argv[1]) if len(sys.argv) > 1 else 49629
x = torch.nn.functional.normalize(torch.randn(N, 768, device="cuda"), dim=1)
knn_topk(x, x, 30); torch.cuda.synchronize()
