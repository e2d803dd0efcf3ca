#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-knn --no-cpu-baseline --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_bulk_kernel" -s 4 -c 1 -f -o gpurun_out/r02_gather $CMD > gpurun_out/ncu_gather.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_gather.log | cut -c1-200
