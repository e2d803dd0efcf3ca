#!/bin/bash
mkdir -p gpurun_out
for h in 1 0; do
DEPTHG_B200_L2HINTS=$h timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-knn --no-extra > gpurun_out/bench_h$h.json 2> gpurun_out/bench.err
python - <<P
import json
d=json.load(open('gpurun_out/bench_h$h.json'))
print("l2hints=$h ms_per_step", round(d["ms_per_step"],4), d["breakdown_us"])
P
done
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -x -k "not knn" > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest.log
