#!/bin/bash
# Round-2 GPU call 1: all GPU tests (incl. the at-size parity tests), smoke, the bench line, per-CTA phase stamps of
# the correlation kernel, the KNN per-kernel breakdown and the ncu launch list of the bench command.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x --durations=15 > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 30 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 300 python scripts/umma_phases.py > gpurun_out/phases.log 2>&1
echo "phases rc=$?"; cat gpurun_out/phases.log
timeout 300 python scripts/knn_profile.py > gpurun_out/knn_profile.log 2>&1
echo "knn rc=$?"; cat gpurun_out/knn_profile.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-knn --no-cpu-baseline --no-extra > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
