#!/bin/bash
# session 2: per-buffer P^T barriers in the D = 64 dK/dV kernel (protocol fix found by tools/bwd_protocol_sim.py): backward
# parity under the watchdog build, then the whole GPU suite on the release build and a timing check (path bit 25 = one buffer)
set -u
OUT=gpurun_out/r2_s2_i; mkdir -p $OUT
AULE_LIBRARY_PATH=$PWD/experiments/ab_wd/libaule.so timeout 300 python -m pytest tests/test_gpu_sm100.py tests/test_gpu_r2.py tests/test_gpu_paged.py -x -q -k "backward or bwd or autograd or golden" > $OUT/pytest_wd.log 2>&1; echo "pytest (watchdog) rc=$?"; tail -2 $OUT/pytest_wd.log
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
for i in 1 2; do
  AULE_PATH=33554432 AULE_SHAPES=B,E,B_long timeout 120 python tools/bwd_time.py >> $OUT/ab.jsonl 2>> $OUT/ab.err
  AULE_PATH=0 AULE_SHAPES=B,E,B_long timeout 120 python tools/bwd_time.py >> $OUT/ab.jsonl 2>> $OUT/ab.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_s2_i/ab.jsonl"):
    d = json.loads(l)
    print(("one S^T buffer " if d["path"] else "two S^T buffers"), {k: (v["ms_median"], v["tflops_median"]) for k, v in d.items() if k not in ("lib", "path")}, [round(x, 3) for x in d["B"]["checksum"]])
PY
