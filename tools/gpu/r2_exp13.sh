#!/bin/bash
set -u
OUT=gpurun_out/r2_exp13; mkdir -p $OUT
for cfg in C D8 B; do
  AULE_SWEEP_CFG=$cfg timeout 300 python tools/sweep_variants.py 20 7 0,29,31 > $OUT/sweep_$cfg.log 2>&1; echo "sweep $cfg rc=$?"; cat $OUT/sweep_$cfg.log
done
