#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/pipe_phases.py > gpurun_out/pipe_phases.log 2>&1; echo "phases rc=$?"; grep -v Warn gpurun_out/pipe_phases.log | tail -20
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 30 gpurun_out/pytest.log
