#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:corr_pipe_kernel -s 3 -c 1 -f -o gpurun_out/pipe_prof python scripts/pipe_phases.py > gpurun_out/ncu_pipe.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_pipe.log; ls -la gpurun_out/pipe_prof.ncu-rep
