#!/bin/bash
# scripts/gpr.sh <timeout> <script> [tail-lines] [--gpus N]: like gp.sh, but retries while the pod answers "busy / transient"
set -e
make -C /root/repo/depthg_b200/csrc -j8 2>&1 | grep -v "^nvcc\|^make\|^g++\|^    -" || true
test -f /root/repo/depthg_b200/libdepthg_b200.so || { echo "BUILD FAILED: no .so"; exit 1; }
python -c "import ctypes; ctypes.CDLL('/root/repo/depthg_b200/libdepthg_b200.so')" || { echo "BUILD FAILED: .so does not load"; exit 1; }
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $4 $5 --timeout $1 -- "bash $2" > /tmp/gpr_$$.log 2>&1 || true
  if grep -q "status=transient\|retry in a few minutes\|rc=3" /tmp/gpr_$$.log && ! grep -q "status=ok" /tmp/gpr_$$.log; then
    echo "[gpr] attempt $i: busy, retrying in 90 s"; sleep 90; continue
  fi
  break
done
tail -${3:-60} /tmp/gpr_$$.log
