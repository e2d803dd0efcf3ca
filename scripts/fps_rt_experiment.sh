#!/bin/bash
# FPS round-thread count experiment: bit-exactness tests + per-kernel time for RT = 32 / 64 / 128.
for rt in 128 256; do
  export DEPTHG_B200_FPS_RT=$rt
  timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "fps" 2>&1 | tail -1
  timeout 200 python bench.py --steps 50 --warmup 5 --no-knn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('RT=$rt', d['ms_per_step'], d['breakdown_us'])"
done
