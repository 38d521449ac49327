#!/bin/bash
# backward P phase: interleaved FFMA2 (default) vs scale/offset pass of its own; plus the window backward tests
set -u
OUT=gpurun_out/r2_exp11; mkdir -p $OUT
PKG=aule-attention_b200
( make -C $PKG clean && make -C $PKG ) > $OUT/build0.log 2>&1 || { tail -20 $OUT/build0.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_r2.py -x -q -k "window_backward or sdpa or rope" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
echo "--- XPASS=0"; timeout 300 python tools/ab_bwd_streams.py 2>&1 | grep "two streams" | tee $OUT/bwd_xpass0.log
( make -C $PKG clean && make -C $PKG EXTRA_NVFLAGS="-DAULE_BWD_XPASS=1" ) > $OUT/build1.log 2>&1 || { tail -20 $OUT/build1.log; exit 1; }
echo "--- XPASS=1"; timeout 300 python tools/ab_bwd_streams.py 2>&1 | grep "two streams" | tee $OUT/bwd_xpass1.log
timeout 600 python -m pytest tests/test_gpu_sm100.py -x -q -k "backward" > $OUT/pytest_bwd_xpass1.log 2>&1; echo "pytest bwd (xpass=1) rc=$?"; tail -3 $OUT/pytest_bwd_xpass1.log
( make -C $PKG clean && make -C $PKG ) > $OUT/build2.log 2>&1
echo "--- XPASS=0 again"; timeout 300 python tools/ab_bwd_streams.py 2>&1 | grep "two streams" | tee $OUT/bwd_xpass0b.log
