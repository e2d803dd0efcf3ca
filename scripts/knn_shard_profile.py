"""Two-phase shard build on ONE GPU (the all-gathered database is already there): per-phase time and kernel breakdown."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depthg_b200 import _lib
from depthg_b200.distributed import knn_shard_bounds
from depthg_b200.precompute_knns import KnnShard, knn_topk
N, world, rank = 49629, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = torch.nn.functional.normalize(torch.randn(N, 768, device="cuda"), dim=1)
lo, hi = knn_shard_bounds(N, world, rank)
local = x[lo:hi].contiguous()
lib = _lib.lib()
def ev():
    return torch.cuda.Event(enable_timing=True)
for it in range(3):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record(); sh = KnnShard(local, lo, N, 30).begin(); e1.record(); idx = sh.finish(x); e2.record()
    torch.cuda.synchronize()
    print("begin %.3f ms  finish %.3f ms  total %.3f ms" % (e0.elapsed_time(e1), e1.elapsed_time(e2), e0.elapsed_time(e2)))
e0, e1 = ev(), ev()
e0.record(); one = knn_topk(local, x, 30); e1.record(); torch.cuda.synchronize()
print("one-call shard %.3f ms   equal rows: %.5f" % (e0.elapsed_time(e1), float((one == idx).float().mean())))
lib.dg_profile_enable(1)
sh = KnnShard(local, lo, N, 30).begin(); idx = sh.finish(x); torch.cuda.synchronize()
n = lib.dg_profile_collect(None, 0); buf = ctypes.create_string_buffer(n + 16); lib.dg_profile_collect(buf, n + 16)
lib.dg_profile_enable(0)
print(buf.value.decode())
