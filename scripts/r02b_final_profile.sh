#!/bin/bash
# Final round-2 evidence on ONE B200: full GPU test run, smoke, the default bench line, the reference arm, the ncu launch
# list of the bench command, one --set full capture of every kernel of a steady-state step, of the KNN kernels (full
# build + two-phase N/8 shard) and of the dense configuration.  Outputs: gpurun_out/r02b_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02b_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=5 > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r02b_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02b_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 1 gpurun_out/r02b_smoke.log
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02b_bench.json; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err
echo "reference arm rc=$?"; tail -c 200 gpurun_out/r02b_bench_reference.json; echo
CMD="python bench.py --steps 3 --warmup 3 --no-knn --no-cpu-baseline --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"corr_pipe_kernel|fps_kernel|gather_bulk_kernel|gather_code_kernel|gather_norm_bwd_kernel" -s 16 -c 6 -f -o gpurun_out/r02b_step $CMD > gpurun_out/ncu_full.log 2>&1
echo "step capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"knn_umma_kernel|knn_rerank_kernel|split_rows_kernel" -s 3 -c 9 -f -o gpurun_out/r02b_knn python scripts/knn_shard_profile.py 3 > gpurun_out/ncu_knn.log 2>&1
echo "knn capture rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"knn_umma_kernel|knn_rerank_kernel" -s 2 -c 2 -f -o gpurun_out/r02b_knn_full python scripts/knn_small.py 49629 768 > gpurun_out/ncu_knn_full.log 2>&1
echo "knn full capture rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"corr_umma_kernel" -s 2 -c 2 -f -o gpurun_out/r02b_dense python scripts/dense_profile.py > gpurun_out/ncu_dense.log 2>&1
echo "dense capture rc=$?"
timeout 200 python scripts/ride_timeline.py > gpurun_out/r02b_ride_timeline.txt 2>&1
timeout 200 python scripts/knn_profile.py > gpurun_out/r02b_knn_profile.txt 2>&1
ls -la gpurun_out/*.ncu-rep
