#!/bin/bash
# session 2, call D: second S^T buffer in the D = 64 dK/dV kernel: parity under the watchdog build first, then interleaved A/B
# (path bit 25 = one buffer) with the release build
set -u
OUT=gpurun_out/r2_s2_d; mkdir -p $OUT
AULE_LIBRARY_PATH=$PWD/experiments/ab_wd/libaule.so timeout 600 python -m pytest tests/test_gpu_sm100.py tests/test_gpu_r2.py tests/test_gpu_paged.py -x -q -k "backward or bwd or autograd or golden" > $OUT/pytest_wd.log 2>&1; rc=$?; echo "pytest (watchdog build) rc=$rc"; tail -5 $OUT/pytest_wd.log
if [ $rc -ne 0 ]; then grep -n "aule\]\|Error\|error" $OUT/pytest_wd.log | head -20; exit 1; fi
for i in 1 2 3; do
  AULE_PATH=33554432 AULE_SHAPES=B,E,B_gqa,B_long timeout 200 python tools/bwd_time.py >> $OUT/ab.jsonl 2>> $OUT/ab.err
  AULE_PATH=0 AULE_SHAPES=B,E,B_gqa,B_long timeout 200 python tools/bwd_time.py >> $OUT/ab.jsonl 2>> $OUT/ab.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_s2_d/ab.jsonl"):
    d = json.loads(l)
    print(("one S^T buffer " if d["path"] else "two S^T buffers"), {k: (v["ms_median"], v["tflops_median"]) for k, v in d.items() if k not in ("lib", "path")}, [round(x, 3) for x in d["B"]["checksum"]])
PY
