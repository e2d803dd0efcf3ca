"""torchrun script: the sharded KNN build with the peer-copy exchange vs the NCCL all-gather vs a one-GPU build."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch, torch.distributed as dist
from depthg_b200.distributed import knn_shard_bounds, sharded_knn_build
from depthg_b200.precompute_knns import knn_topk
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
N, F, k = int(os.environ.get("KNN_N", 49629)), 768, 30
g = torch.Generator(device=dev).manual_seed(7)
x = torch.nn.functional.normalize(torch.randn((N, F), generator=g, device=dev), dim=1)
lo, hi = knn_shard_bounds(N, world, rank)
local = x[lo:hi].contiguous()
res = {}
for mode in ("peer", "nccl"):
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); idx, stats = sharded_knn_build(local, N, k, return_stats=True, exchange=mode); e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0: print(mode, "iter", it, "max-over-ranks ms %.3f" % t.item(), stats, flush=True)
        if mode == "peer" and it == 3 and rank in (0, world - 1) and os.environ.get("DEPTHG_KNN_TRACE"):
            from depthg_b200 import distributed as D
            st = list(D._symm_states.values())[0]
            n0, h0, ev0 = st.trace[0]
            print("rank", rank, " ".join("%s: gpu %.3f host %.3f |" % (n, ev0.elapsed_time(ev), (h - h0) * 1e3) for n, h, ev in st.trace), flush=True)
    res[mode] = idx
one = knn_topk(local, x, k)
for mode, idx in res.items():
    diff = idx != one
    worst = 0.0
    if bool(diff.any()):
        sims = local @ x.T
        worst = float((torch.gather(sims, 1, idx) - torch.gather(sims, 1, one)).abs()[diff].max())
    t = torch.tensor([worst], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(mode, "vs one-call build: worst fp32 gap over all ranks", t.item(), "differing slots (rank 0)", int(diff.sum()))
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
