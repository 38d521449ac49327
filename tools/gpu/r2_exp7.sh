#!/bin/bash
# round-2 experiment 7: full GPU test-suite on a watchdog build, then the release build + bench.py
set -u
OUT=gpurun_out/r2_exp7; mkdir -p $OUT
PKG=aule-attention_b200
( make -C $PKG clean && make -C $PKG EXTRA_NVFLAGS="-DAULE_WATCHDOG=1" ) > $OUT/build_wd.log 2>&1 || { tail -30 $OUT/build_wd.log; exit 1; }
timeout 1700 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
( make -C $PKG clean && make -C $PKG ) > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -5 $OUT/bench.err; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(json.dumps({k:d[k] for k in ('value','ms_per_step','clocks','e2e','roofline','cpu_baseline')},indent=1)[:3000]);print(json.dumps(d['secondary'],indent=1)[:6000])"
