#!/bin/bash
# release build as the driver sees it: full GPU test-suite, smoke, reference arm, bench
set -u
OUT=gpurun_out/r2_final1; mkdir -p $OUT
timeout 1700 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python -c "
import json;d=json.load(open('$OUT/bench.json'));print(json.dumps({k:d[k] for k in ('value','ms_per_step','clocks','roofline','cpu_baseline')})[:1500]);print(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['copy_ceiling']['e2e_frac_of_ceiling']);s=d['secondary'];print({k:(v.get('tflops') or v.get('bwd_tflops') or v.get('GBps') or v.get('tflops_total')) if isinstance(v,dict) else v for k,v in s.items()})"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-400 $OUT/bench_ref.json
timeout 600 python tools/bench_extra.py > $OUT/extra.jsonl 2> $OUT/extra.err; echo "extra rc=$?"; cut -c1-700 $OUT/extra.jsonl
