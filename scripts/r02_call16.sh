#!/bin/bash
mkdir -p gpurun_out
for dbg in 3 4 7 5; do
DEPTHG_B200_GATHER_DBG=$dbg timeout 300 python bench.py --steps 30 --warmup 5 --no-extra --no-cpu-baseline --no-knn > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err
python - <<P
import json
d=json.load(open('gpurun_out/bench_dbg.json'))
print("dbg=$dbg ms_per_step", round(d["ms_per_step"],4), d["breakdown_us"].get("gather_bulk_kernel"))
P
done
