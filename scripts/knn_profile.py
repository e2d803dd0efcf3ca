import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depthg_b200 import _lib
from depthg_b200.precompute_knns import knn_topk
N = int(sys.argv[1]) if len(sys.argv) > 1 else 49629
x = torch.nn.functional.normalize(torch.randn(N, 768, device="cuda"), dim=1)
knn_topk(x, x, 30); torch.cuda.synchronize()
lib = _lib.lib()
lib.dg_profile_enable(1)
knn_topk(x, x, 30)
n = lib.dg_profile_collect(None, 0); buf = ctypes.create_string_buffer(n + 16); lib.dg_profile_collect(buf, n + 16)
lib.dg_profile_enable(0)
print(buf.value.decode())
# a multi-GPU-sized shard: N/8 query rows against the full database
import time
q = x[: N // 8].contiguous()
knn_topk(q, x, 30); torch.cuda.synchronize()
t = time.time(); knn_topk(q, x, 30); torch.cuda.synchronize(); print("shard N/8 rows: %.2f ms" % ((time.time() - t) * 1e3))
t = time.time(); knn_topk(x, x, 30); torch.cuda.synchronize(); print("full: %.2f ms" % ((time.time() - t) * 1e3))
# per-kernel breakdown of the shard call
lib.dg_profile_enable(1)
knn_topk(q, x, 30); torch.cuda.synchronize()
n = lib.dg_profile_collect(None, 0); buf = ctypes.create_string_buffer(n + 16); lib.dg_profile_collect(buf, n + 16)
lib.dg_profile_enable(0)
print("shard breakdown:\n" + buf.value.decode())
for rows in (N // 4, N // 2):
    qq = x[:rows].contiguous()
    knn_topk(qq, x, 30); torch.cuda.synchronize()
    t = time.time(); knn_topk(qq, x, 30); torch.cuda.synchronize(); print("rows %d: %.2f ms" % (rows, (time.time() - t) * 1e3))
