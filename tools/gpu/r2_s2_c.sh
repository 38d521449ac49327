#!/bin/bash
# session 2, call C: pipeline traces of the D = 64 backward kernels on config B (CTA 0 = heaviest key / query block)
set -u
OUT=gpurun_out/r2_s2_c; mkdir -p $OUT
export AULE_LIBRARY_PATH=$PWD/experiments/ab_trace/libaule.so AULE_TRACE_SHAPE=4,32,32,2048,64
timeout 300 python tools/bwd_trace.py dkv 0 15 > $OUT/trace_dkv_B.txt 2>&1; echo "dkv rc=$?"
timeout 300 python tools/bwd_trace.py dq 0 15 > $OUT/trace_dq_B.txt 2>&1; echo "dq rc=$?"
head -3 $OUT/trace_dkv_B.txt; grep -c . $OUT/trace_dkv_B.txt
