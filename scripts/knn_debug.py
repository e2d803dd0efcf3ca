import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depthg_b200 import _lib
from depthg_b200._lib import ptr, stream_ptr, check
N, F, k = 8192, 768, 30
torch.manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(N, F, device="cuda"), dim=1)
lib = _lib.lib()
wsb = lib.dg_knn_workspace_bytes(N, N, F, k)
ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
idx = torch.empty((N, k), dtype=torch.int64, device="cuda")
check(lib.dg_knn_topk(ptr(x), ptr(x), N, N, F, k, ptr(idx), None, ptr(ws), wsb, stream_ptr()), "knn")
torch.cuda.synchronize()
al = lambda b: (b + 255) // 256 * 256
off = 256 + 2 * al(N * F * 2) + 2 * al(N * F * 2)
ci = ws[off:off + N * 32 * 4].view(torch.int32).view(N, 32); off += al(N * 32 * 4)
cv = ws[off:off + N * 32 * 4].view(torch.float32).view(N, 32)
hdr = ws[:8].view(torch.int32)
print("err, fail_count:", hdr.tolist())
sims = x[:256] @ x.T
ev, ei = torch.topk(sims, 32)
approx = cv[:256]
exact_of_cand = torch.gather(sims, 1, ci[:256].long())
print("max |approx - exact| over candidates:", (approx - exact_of_cand).abs().max().item())
print("rows where candidate set == true top32 set:", sum(set(a.tolist()) == set(b.tolist()) for a, b in zip(ci[:256].cpu(), ei.cpu())), "/ 256")
gap = ev[:, 29] - ev[:, 31]
print("exact gap 30th-32nd: min %.2e median %.2e" % (gap.min().item(), gap.median().item()))
print("approx sorted desc ok:", bool((approx[:, :-1] >= approx[:, 1:]).all()))
print(approx[0, :6].tolist(), ev[0, :6].tolist())
