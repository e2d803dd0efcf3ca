#!/bin/bash
mkdir -p gpurun_out
for cfg in "32 11" "8 11"; do timeout 300 python scripts/pipe_account.py $cfg 2>&1 | grep -v Warn | tail -4; done
