#!/bin/bash
set -u
OUT=gpurun_out/r2_s2_f; mkdir -p $OUT
AULE_LIBRARY_PATH=$PWD/experiments/ab_trace/libaule.so timeout 300 python tools/bwd_trace.py fused2 100 103 > $OUT/trace_fused2.txt 2>&1; echo "rc=$?"
head -2 $OUT/trace_fused2.txt
