#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -x -k "not knn" > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-knn --no-extra > gpurun_out/bench_x.json 2> gpurun_out/bench.err
python - <<P
import json
d=json.load(open('gpurun_out/bench_x.json'))
print("ms_per_step", round(d["ms_per_step"],4), d["breakdown_us"], "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline"]["traffic"])
P
