#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, a short bench; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 ${PYTEST_ARGS:-} > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"
tail -n 40 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-50} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 4000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
