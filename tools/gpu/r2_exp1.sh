#!/bin/bash
# round-2 experiment 1: v5 "stream" forward kernel -- correctness (watchdog build), then A/B against v4 + traces
set -u
OUT=gpurun_out/r2_exp1; mkdir -p $OUT
PKG=aule-attention_b200
( make -C $PKG clean && make -C $PKG EXTRA_NVFLAGS="-DAULE_WATCHDOG=1 -DAULE_TUNING_VARIANTS" ) > $OUT/build_wd.log 2>&1 || { tail -30 $OUT/build_wd.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_sm100.py -x -q > $OUT/pytest_sm100.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_sm100.log
tail -15 $OUT/pytest_sm100.log
( make -C $PKG clean && make -C $PKG EXTRA_NVFLAGS="-DAULE_TUNING_VARIANTS" ) > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
for cfg in C D8 B; do
  AULE_SWEEP_CFG=$cfg timeout 300 python tools/sweep_variants.py 20 5 0,16384,17,18 > $OUT/sweep_$cfg.log 2>&1; echo "sweep $cfg rc=$?"; cat $OUT/sweep_$cfg.log
done
timeout 300 python tools/fwd_trace.py C 2000 200 > $OUT/trace_C.log 2>&1; echo "trace rc=$?"; head -80 $OUT/trace_C.log
timeout 300 python tools/fwd_trace.py B 1000 200 > $OUT/trace_B.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json | cut -c1-1500
