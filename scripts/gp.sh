#!/bin/bash
# build, check, then run a script on the GPU box:  scripts/gp.sh <timeout> <script> 
set -e
make -C /root/repo/depthg_b200/csrc -j8 2>&1 | grep -v "^nvcc\|^make\|^g++\|^    -" || true
test -f /root/repo/depthg_b200/libdepthg_b200.so || { echo "BUILD FAILED: no .so"; exit 1; }
python -c "import ctypes; ctypes.CDLL('/root/repo/depthg_b200/libdepthg_b200.so')" || { echo "BUILD FAILED: .so does not load"; exit 1; }
/usr/local/graft/bin/gpurun --timeout $1 -- "bash $2" 2>&1 | tail -${3:-60}
