#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=8 > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
