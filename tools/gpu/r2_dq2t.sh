#!/bin/bash
mkdir -p gpurun_out/r2_dq2t
timeout 300 python tools/bwd_trace.py dq 13 14 > gpurun_out/r2_dq2t/trace.txt 2>&1
echo "rc=$?"; head -100 gpurun_out/r2_dq2t/trace.txt
