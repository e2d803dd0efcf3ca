#!/bin/bash
# ncu evidence for the bench command (run under gpurun): launch list + full captures of the top kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-knn --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-corr_tile_kernel|fps_kernel|gather_norm_kernel|gather_norm_bwd_kernel|pair_means_kernel}" -s ${NCU_SKIP:-33} -c ${NCU_COUNT:-11} -f -o gpurun_out/prof_step $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
