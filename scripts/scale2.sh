#!/bin/bash
# 2-GPU experiment: how the per-step head-gradient all-reduce is issued vs weak-scaling efficiency.
mkdir -p gpurun_out
for mode in none async graph; do
  DEPTHG_BENCH_ALLREDUCE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 100 --warmup 10 --no-knn --no-cpu-baseline > gpurun_out/scale2_$mode.json 2> gpurun_out/scale2_$mode.err
  echo "$mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale2_$mode.json") if l.startswith("{")][-1])
    print("$mode", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms", "fused:", d.get("fused_negative_sampler",{}).get("ms_per_step"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/scale2_$mode.err").read()[-1500:])
PY
done
timeout 300 python bench.py --steps 100 --warmup 10 --no-knn --no-cpu-baseline > gpurun_out/scale1.json 2>gpurun_out/scale1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale1.json') if l.startswith('{')][-1]); print('N=1', round(d['value']), d['ms_per_step'], 'fused', d['fused_negative_sampler']['ms_per_step'])"
