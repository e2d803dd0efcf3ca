#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/s12_diag.py 12 122 > gpurun_out/s12_diag.log 2>&1; cat gpurun_out/s12_diag.log | grep -v Warning
timeout 600 python scripts/s12_diag.py 11 99 > gpurun_out/s11_diag.log 2>&1; cat gpurun_out/s11_diag.log | grep -v Warning
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -k "salience or depth_shapes or fps_depth_feat or norm_accepts" > gpurun_out/pytest2.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest2.log
