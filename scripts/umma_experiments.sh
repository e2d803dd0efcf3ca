#!/bin/bash
# timing experiments on the tcgen05 correlation kernel (results are garbage in dbg modes; only times matter)
for d in 0 1 2 3; do
  echo "== DEPTHG_B200_UMMA_DBG=$d"
  DEPTHG_B200_UMMA_DBG=$d python bench.py --steps 30 --warmup 5 --no-knn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:round(v,1) for k,v in d['breakdown_us'].items()}, 'step', round(d['ms_per_step']*1e3,1))"
done
