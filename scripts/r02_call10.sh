#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=8 -x > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 300 python scripts/knn_shard_profile.py 3 2>&1 | grep -v Warn
timeout 300 python scripts/knn_shard_profile.py 0 2>&1 | grep -v Warn | head -5
timeout 900 python bench.py --steps 50 --warmup 5 --no-extra > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "graph", d["cuda_graph"]["ms_per_step"], "torch-sampler", d["torch_negative_sampler"]["ms_per_step"])
print("breakdown", d["breakdown_us"])
print("e2e", d["e2e"]["value"], d["e2e"]["mode"])
print("knn", d["knn"]["ms"], d["knn"]["parity_checked"])
P
tail -3 gpurun_out/bench.err
