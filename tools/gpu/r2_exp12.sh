#!/bin/bash
set -u
OUT=gpurun_out/r2_exp12; mkdir -p $OUT
timeout 120 python tools/softmax_only.py 4,14 2>&1 | tee $OUT/softmax_only.log
for cfg in C D8 B; do
  AULE_SWEEP_CFG=$cfg timeout 300 python tools/sweep_variants.py 20 7 0,29,31 > $OUT/sweep_$cfg.log 2>&1; echo "sweep $cfg rc=$?"; cat $OUT/sweep_$cfg.log
done
