#!/bin/bash
# session 2, last call: fused2 built with deeper Q / dO rings and ONE staging tile (-DAULE_F2_NQ=4 -DAULE_F2_NDO=3 -DAULE_F2_STG=1)
set -u
OUT=gpurun_out/r2_s2_k; mkdir -p $OUT
AULE_FUSED_BIT=26 AULE_LIBRARY_PATH=$PWD/experiments/ab_f2b/libaule.so timeout 100 python tools/check_bwd_fused.py --units 8 > $OUT/ab_deep_rings.log 2>&1; echo "rc=$?"; grep "CHECK\|FAIL" $OUT/ab_deep_rings.log; tail -2 $OUT/ab_deep_rings.log
