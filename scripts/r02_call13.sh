#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu-baseline --no-knn > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "graph", d["cuda_graph"]["ms_per_step"], "torch-sampler", d["torch_negative_sampler"]["ms_per_step"])
print("breakdown", d["breakdown_us"])
P
tail -3 gpurun_out/bench.err
