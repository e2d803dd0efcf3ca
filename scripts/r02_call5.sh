#!/bin/bash
mkdir -p gpurun_out
for b in 32 16 8; do
timeout 300 python scripts/pipe_phases.py $b > gpurun_out/pipe_phases_$b.log 2>&1; echo "== B=$b rc=$?"; grep -v Warn gpurun_out/pipe_phases_$b.log | grep "CTA   0\|CTA  40\|CTA 147\|span\|_kernel"
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest.log
