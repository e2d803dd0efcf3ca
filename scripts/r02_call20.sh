#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest.log
for pdl in 1 0; do
DEPTHG_B200_PDL=$pdl timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-knn --no-extra > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench.err
python - <<P
import json
d=json.load(open('gpurun_out/bench_pdl$pdl.json'))
print("pdl=$pdl ms_per_step", round(d["ms_per_step"],4), "graph", d["cuda_graph"].get("ms_per_step"), "torch", round(d["torch_negative_sampler"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
P
done
