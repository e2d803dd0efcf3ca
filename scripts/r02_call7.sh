#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "dense_28x28" > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "graph", d["cuda_graph"], "fused", d["fused_negative_sampler"]["ms_per_step"])
print("breakdown", d["breakdown_us"])
print("roofline", {k:d["roofline"][k] for k in ("kernel","frac","us_per_step")}, "step frac", d["roofline_step"]["frac"])
print("e2e", d["e2e"]["value"], d["e2e"]["mode"])
print("knn", d["knn"]["ms"], d["knn"]["parity_checked"])
for k,v in d["extra_configs"].items(): print(k, v.get("ms_per_step"), v.get("breakdown_us"), v.get("corr_kernel",{}).get("frac_of_bf16_sustained_div3"), v.get("error"))
P
tail -3 gpurun_out/bench.err
