#!/bin/bash
# session 2, call B: vectorised Delta pre-pass + read-only TMA-store wait in the dK/dV epilogue: parity (backward tests) and
# interleaved A/B against the previous build (experiments/ab_old/libaule.so)
set -u
OUT=gpurun_out/r2_s2_b; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sm100.py tests/test_gpu_r2.py tests/test_gpu_paged.py -x -q -k "backward or bwd or autograd or golden or rope" > $OUT/pytest_bwd.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_bwd.log
for i in 1 2 3; do
  AULE_LIBRARY_PATH=$PWD/experiments/ab_old/libaule.so timeout 300 python tools/bwd_time.py >> $OUT/ab.jsonl 2>> $OUT/ab.err
  timeout 300 python tools/bwd_time.py >> $OUT/ab.jsonl 2>> $OUT/ab.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_s2_b/ab.jsonl"):
    d = json.loads(l)
    print(("old" if "ab_old" in d["lib"] else "new"), {k: (v["ms_median"], v["ms_min"], v["tflops_median"]) for k, v in d.items() if k != "lib"}, [round(x, 2) for x in d["B"]["checksum"]])
PY
