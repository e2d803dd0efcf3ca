#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "knn" -x > gpurun_out/pytest_knn.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_knn.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 300 python scripts/knn_shard_profile.py 3 2>&1 | grep -v Warn
timeout 300 python scripts/knn_profile.py 2>&1 | grep -v Warn
