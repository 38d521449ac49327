#!/bin/bash
# fused backward: first correctness check + timing (watchdog build so a protocol bug traps instead of hanging)
mkdir -p gpurun_out/r2_bwdf1
timeout 600 python tools/check_bwd_fused.py > gpurun_out/r2_bwdf1/check.log 2>&1
echo "rc=$?"; tail -30 gpurun_out/r2_bwdf1/check.log
