#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x -k "not knn" > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest.log
for dbg in 0 2; do
DEPTHG_B200_GATHER_DBG=$dbg timeout 300 python bench.py --steps 30 --warmup 5 --no-extra --no-cpu-baseline --no-knn > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err
python - <<P
import json
d=json.load(open('gpurun_out/bench_dbg.json'))
print("dbg=$dbg ms_per_step", round(d["ms_per_step"],4), d["breakdown_us"])
P
done
