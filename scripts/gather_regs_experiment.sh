#!/bin/bash
# try different occupancy targets for the feature gather (rebuilds gather.cu on the GPU box)
for m in 2 3 4; do
  touch depthg_b200/csrc/gather.cu
  make -C depthg_b200/csrc EXTRA=-DGF_MINBLOCKS=$m > /dev/null 2>&1
  grep -A2 "gather_feats_kernelILi6ELi1" depthg_b200/csrc/build/gather.ptxas.log | grep -E "registers|spill" | tr '\n' ' '
  echo
  python bench.py --steps 100 --warmup 10 --no-knn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('GF_MINBLOCKS=$m', {k:round(v,1) for k,v in d['breakdown_us'].items() if 'gather' in k}, 'fused step', round(d['fused_negative_sampler']['ms_per_step']*1e3,1))"
done
