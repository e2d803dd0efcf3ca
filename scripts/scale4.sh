#!/bin/bash
# 4-GPU experiment: where does the weak-scaling loss come from (host contention vs the NCCL kernel)?
mkdir -p gpurun_out; nproc
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 4 --steps 100 --warmup 10 --no-knn --no-cpu-baseline > gpurun_out/s4_$name.json 2> gpurun_out/s4_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/s4_$name.json") if l.startswith("{")][-1])
    print("$name", round(d["value"]), round(d["ms_per_step"],4), "fused", round(d["fused_negative_sampler"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["e2e"].get("unpipelined"))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/s4_$name.err").read()[-800:])
PY
}
run none DEPTHG_BENCH_ALLREDUCE=none
run graph DEPTHG_BENCH_ALLREDUCE=graph
run graph_cta2 DEPTHG_BENCH_ALLREDUCE=graph NCCL_MAX_CTAS=2
