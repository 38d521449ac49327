#!/bin/bash
mkdir -p gpurun_out/r2_bwdf2
timeout 300 python tools/bwd_trace.py fused 20 24 > gpurun_out/r2_bwdf2/trace.txt 2>&1
echo "rc=$?"; head -120 gpurun_out/r2_bwdf2/trace.txt
