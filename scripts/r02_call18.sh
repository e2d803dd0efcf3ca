#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-knn > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<P
import json
d=json.load(open('gpurun_out/bench.json'))
print("ms_per_step", round(d["ms_per_step"],4), d["breakdown_us"], "graph", d["cuda_graph"].get("ms_per_step"), "torch", d["torch_negative_sampler"]["ms_per_step"])
for k,v in d["extra_configs"].items(): print(k, v.get("ms_per_step"), v.get("breakdown_us"), v.get("error"))
P
