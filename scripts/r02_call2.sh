#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=8 > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 40 gpurun_out/pytest.log
