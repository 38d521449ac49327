#!/bin/bash
# round-2 experiment 2: softmax-only floor, softmax loop variants, D8 trace
set -u
OUT=gpurun_out/r2_exp2; mkdir -p $OUT
timeout 120 python tools/softmax_only.py > $OUT/softmax_only.log 2>&1; echo "so rc=$?"; cat $OUT/softmax_only.log
for cfg in C D8 B; do
  AULE_SWEEP_CFG=$cfg timeout 300 python tools/sweep_variants.py 20 5 0,16384,17,18,19 > $OUT/sweep_$cfg.log 2>&1; echo "sweep $cfg rc=$?"; cat $OUT/sweep_$cfg.log
done
timeout 300 python tools/fwd_trace.py D8 2000 260 > $OUT/trace_D8.log 2>&1; echo "trace rc=$?"; head -60 $OUT/trace_D8.log
