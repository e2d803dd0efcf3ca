#!/bin/bash
mkdir -p gpurun_out
NP=${NP:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $NP --steps 50 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_n$NP.json 2> gpurun_out/r02_bench_n$NP.err
echo "rc=$?"; NP=$NP python - <<'P'
import json,os
np_=os.environ["NP"]
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_n%s.json"%np_) if l.startswith("{")][-1])
    print("N=",np_,"ms_per_step", round(d["ms_per_step"],4), "value", round(d["value"]), d["config"]["allreduce_issue"], "e2e", round(d["e2e"]["value"]), d["e2e"]["mode"])
    k=d.get("knn")
    if k: print("knn ms", round(k["ms"],3), k["parity_checked"])
except Exception as e:
    print("ERR", e); print(open('gpurun_out/r02_bench_n%s.err'%np_).read()[-2500:])
P
