#!/bin/bash
# Round-2 ncu evidence: launch list of the bench command + one --set full capture of every kernel of a steady-state step
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-knn --no-cpu-baseline --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"corr_pipe_kernel|fps_kernel|gather_feats_kernel|gather_code_kernel|gather_norm_bwd_kernel|super_perms_kernel" -s 24 -c 6 -f -o gpurun_out/r02_step $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/r02_step.ncu-rep
