#!/bin/bash
# 8-GPU experiment: how to issue the per-step head-gradient all-reduce.
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 8 --steps 100 --warmup 10 --no-knn --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/s8_$name.json") if l.startswith("{")][-1])
    print("$name", round(d["value"]), round(d["ms_per_step"],4), "fused", round(d["fused_negative_sampler"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/s8_$name.err").read()[-800:])
PY
}
run graph DEPTHG_BENCH_ALLREDUCE=graph
run graph_hp DEPTHG_BENCH_ALLREDUCE=graph_hp
run inline DEPTHG_BENCH_ALLREDUCE=inline
