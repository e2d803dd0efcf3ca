#!/bin/bash
mkdir -p gpurun_out
DEPTHG_B200_CORR=umma1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-knn > gpurun_out/bench_umma1.json 2> gpurun_out/bench_umma1.err
python - <<P
import json
d=json.load(open('gpurun_out/bench_umma1.json'))
print("umma1 ms_per_step", round(d["ms_per_step"],4), d["breakdown_us"])
for k,v in d["extra_configs"].items(): print(k, v.get("ms_per_step"), v.get("breakdown_us"), v.get("error"))
P
