#!/bin/bash
set -u
OUT=gpurun_out/r2_exp15; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sm100.py tests/test_gpu_r2.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
for cfg in C B D8; do
  AULE_SWEEP_CFG=$cfg timeout 300 python tools/sweep_variants.py 20 7 0,16384 > $OUT/sweep_$cfg.log 2>&1; echo "sweep $cfg rc=$?"; cat $OUT/sweep_$cfg.log
done
