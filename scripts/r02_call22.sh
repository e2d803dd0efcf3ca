#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest.log
