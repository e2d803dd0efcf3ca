#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/pipe_phases.py 32 2>&1 | grep -v Warn | grep "CTA   0\|CTA  40\|CTA  75\|CTA  76\|CTA 100\|CTA 147\|span\|corr_pipe"
timeout 300 python scripts/pipe_account.py 32 11 2>&1 | grep -v Warn | tail -4
