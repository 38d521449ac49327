#!/bin/bash
# session 2, final validation: fused2 tests under the watchdog build first (bounded), then the release build -- full GPU suite,
# smoke, bench (N=1) + reference arm, launch list of the bench command
set -u
OUT=gpurun_out/r2_s2_final2; mkdir -p $OUT
AULE_LIBRARY_PATH=$PWD/experiments/ab_wd/libaule.so timeout 240 python -m pytest tests/test_gpu_r2.py -x -q -k "fused" > $OUT/pytest_fused_wd.log 2>&1; rc=$?; echo "pytest fused (watchdog build) rc=$rc"; tail -2 $OUT/pytest_fused_wd.log
if [ $rc -ne 0 ]; then grep -n "aule\]\|Error\|FAIL" $OUT/pytest_fused_wd.log | head -10; fi
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench reference rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_s2_final2/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "clocks")}, d["roofline"]["frac"], d["e2e"]["value"])
for k, v in d["secondary"].items():
    if isinstance(v, dict):
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("tflops", "bwd_tflops", "fwd_tflops", "GBps", "tflops_total", "ms", "bwd_ms", "error")})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-secondary > $OUT/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
