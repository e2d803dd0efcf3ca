#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 8 --steps 100 --warmup 10 --no-extra --no-cpu-baseline --no-knn > gpurun_out/n8_$name.json 2> gpurun_out/n8_$name.err
  echo "== $name rc=$?"; python - "$name" <<'P'
import json,sys
try:
    d=json.loads([l for l in open("gpurun_out/n8_%s.json"%sys.argv[1]) if l.startswith("{")][-1])
    print("ms_per_step", round(d["ms_per_step"],4), "value", round(d["value"]), "torch-sampler", round(d["torch_negative_sampler"]["ms_per_step"],4), d["config"]["allreduce_issue"], "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("ERR", e); print(open('gpurun_out/n8_%s.err'%sys.argv[1]).read()[-1500:])
P
}
run symm DEPTHG_BENCH_ALLREDUCE=symm
run symm_inline DEPTHG_BENCH_ALLREDUCE=symm_inline
