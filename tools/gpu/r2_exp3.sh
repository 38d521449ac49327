#!/bin/bash
# round-2 experiment 3: just-in-time work claims, exp2-emulation share, softmax-only phase experiments
set -u
OUT=gpurun_out/r2_exp3; mkdir -p $OUT
timeout 120 python tools/softmax_only.py 4,5,6,7 > $OUT/softmax_only.log 2>&1; echo "so rc=$?"; cat $OUT/softmax_only.log
for cfg in C D8 B; do
  AULE_SWEEP_CFG=$cfg timeout 300 python tools/sweep_variants.py 20 5 0,16384,17,18,19 > $OUT/sweep_$cfg.log 2>&1; echo "sweep $cfg rc=$?"; cat $OUT/sweep_$cfg.log
done
timeout 900 python -m pytest tests/test_gpu_sm100.py -x -q > $OUT/pytest_sm100.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_sm100.log
