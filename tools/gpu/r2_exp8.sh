#!/bin/bash
set -u
OUT=gpurun_out/r2_exp8; mkdir -p $OUT
timeout 1700 python -m pytest tests/test_gpu_r2.py tests/test_gpu_fp32.py tests/test_gpu_paged.py -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -40 $OUT/pytest_gpu.log
timeout 300 python tools/ab_bwd_streams.py 2>&1 | tee $OUT/ab_bwd_streams.log
