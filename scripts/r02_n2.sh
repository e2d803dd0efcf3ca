#!/bin/bash
# 2-GPU check of the multi-rank paths: loss step with the all-reduce under the next FPS vs inline; KNN two-phase vs serial
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 50 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/n2_$name.json 2> gpurun_out/n2_$name.err
  echo "== $name rc=$?"; python - "$name" <<'P'
import json,sys
try:
    d=json.loads([l for l in open("gpurun_out/n2_%s.json"%sys.argv[1]) if l.startswith("{")][-1])
    print("ms_per_step", round(d["ms_per_step"],4), "value", round(d["value"]), "torch-sampler", round(d["torch_negative_sampler"]["ms_per_step"],4), d["config"]["allreduce_issue"])
    k=d["knn"]; print("knn ms", round(k["ms"],3), k["parity_checked"])
except Exception as e:
    print("ERR", e); print(open('gpurun_out/n2_%s.err'%sys.argv[1]).read()[-1500:])
P
}
run fps_overlap X=1
run inline_serial DEPTHG_BENCH_ALLREDUCE=inline DEPTHG_BENCH_KNN=serial
