#!/bin/bash
mkdir -p gpurun_out
DEPTHG_KNN_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 scripts/knn_peer_test.py 2>&1 | grep -v "Warn\|OMP_NUM\|\*\*\*\*" | tail -30
