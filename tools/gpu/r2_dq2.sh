#!/bin/bash
mkdir -p gpurun_out/r2_dq2
timeout 600 python tools/check_bwd_dq2.py > gpurun_out/r2_dq2/check.log 2>&1
echo "rc=$?"; grep -v "watchdog: block 0 thread [0-9]*[1-9] " gpurun_out/r2_dq2/check.log | tail -40
