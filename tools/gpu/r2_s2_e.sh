#!/bin/bash
# session 2: fused backward in 64-query half steps (path bit 26): parity under the watchdog build, A/B timing, trace
set -u
OUT=gpurun_out/r2_s2_e; mkdir -p $OUT
AULE_FUSED_BIT=26 AULE_LIBRARY_PATH=$PWD/experiments/ab_wd/libaule.so timeout 300 python tools/check_bwd_fused.py --no-time > $OUT/check.log 2>&1; echo "check rc=$?"; grep -c " ok" $OUT/check.log; grep "FAIL\|CHECK\|watchdog\|rror" $OUT/check.log | head
AULE_FUSED_BIT=26 AULE_EXTRA_PATH=268435456 AULE_LIBRARY_PATH=$PWD/experiments/ab_wd/libaule.so timeout 300 python tools/check_bwd_fused.py --no-time > $OUT/check_cf.log 2>&1; echo "check (fence helper warp) rc=$?"; grep "FAIL\|CHECK\|watchdog\|rror" $OUT/check_cf.log | head
if grep -q "CHECK PASS" $OUT/check.log; then
  AULE_FUSED_BIT=26 timeout 300 python tools/check_bwd_fused.py --units 8,0 > $OUT/ab.log 2>&1; echo "ab rc=$?"; tail -3 $OUT/ab.log
  AULE_FUSED_BIT=26 AULE_EXTRA_PATH=268435456 timeout 300 python tools/check_bwd_fused.py --units 8,0 > $OUT/ab_cf.log 2>&1; echo "ab (fence helper warp) rc=$?"; tail -3 $OUT/ab_cf.log
  AULE_EXTRA_PATH=268435456 AULE_LIBRARY_PATH=$PWD/experiments/ab_trace/libaule.so timeout 300 python tools/bwd_trace.py fused2 100 102 > $OUT/trace_fused2.txt 2>&1; echo "trace rc=$?"
fi
