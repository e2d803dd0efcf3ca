#!/bin/bash
# 8-GPU bench line as the driver launches it (loss step with the all-reduce under the next FPS; KNN with the peer-copy exchange)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "rc=$?"; python - <<'P'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_n8.json") if l.startswith("{")][-1])
    print("ms_per_step", round(d["ms_per_step"],4), "value", round(d["value"]), "torch-sampler", round(d["torch_negative_sampler"]["ms_per_step"],4), d["config"]["allreduce_issue"], "e2e", round(d["e2e"]["value"]), d["e2e"]["mode"])
    k=d.get("knn")
    if k: print("knn ms", round(k["ms"],3), k["sharding"][:60], k["parity_checked"])
except Exception as e:
    print("ERR", e); print(open('gpurun_out/r02_bench_n8.err').read()[-2500:])
P
