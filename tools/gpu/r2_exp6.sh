#!/bin/bash
set -u
OUT=gpurun_out/r2_exp6; mkdir -p $OUT
timeout 300 python tools/softmax_only.py 4,5,6,7,8,9,10,11 2>&1 | tee $OUT/softmax_only.log
for cfg in C D8 B; do
  AULE_SWEEP_CFG=$cfg timeout 300 python tools/sweep_variants.py 20 5 0,28,29,30,31 > $OUT/sweep_$cfg.log 2>&1; echo "sweep $cfg rc=$?"; cat $OUT/sweep_$cfg.log
done
