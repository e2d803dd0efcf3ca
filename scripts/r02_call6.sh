#!/bin/bash
mkdir -p gpurun_out
for b in 32 8; do
for m in 0 2; do
export DEPTHG_B200_PIPE_MMA=$m
timeout 300 python scripts/pipe_phases.py $b > gpurun_out/pp.log 2>&1; echo "== B=$b mma_dbg=$m rc=$?"; grep -v Warn gpurun_out/pp.log | grep "CTA   0\|CTA  40\|CTA 147\|span\|corr_pipe"
done; done
unset DEPTHG_B200_PIPE_MMA
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest.log
