#!/bin/bash
# session 2 validation: release build -- full GPU suite, smoke, bench (N=1) + reference arm, ncu captures of the dK/dV kernel
# (D = 128 on config C/2, D = 64 on config B) and the launch list of the bench command
set -u
OUT=gpurun_out/r2_s2_final; mkdir -p $OUT
timeout 1700 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench reference rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_s2_final/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "clocks")}, d["roofline"]["frac"], d["e2e"]["value"])
for k, v in d["secondary"].items():
    if isinstance(v, dict):
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("tflops", "bwd_tflops", "fwd_tflops", "GBps", "tflops_total", "ms", "bwd_ms", "exact_fp32_cuda_core_tflops", "torch_sdpa_tflops", "torch_sdpa_fp32_tflops", "max_err_vs_exact_rel_to_scale", "error")})
r = json.load(open("gpurun_out/r2_s2_final/bench_reference.json"))
print("reference arm:", r.get("value"), r.get("cpu_baseline"))
PY
cat > /tmp/one_bwd.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "aule-attention_b200", "python"))
from aule import cuda_flash, ffi
lib = ffi.ensure_init()
lib.aule_set_kernel_path(1 << 16)      # one stream: the profiler serialises kernels anyway
for (B, Hq, Hkv, S, D) in ((4, 32, 8, 4096, 128), (4, 32, 32, 2048, 64)):
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    o, lse = cuda_flash.forward_with_lse(q, k, v, causal=True)
    do = torch.randn_like(o)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    for _ in range(3):
        rc = lib.aule_attention_backward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                              B, Hq, Hkv, S, S, D, ffi.DTYPE_BF16, 0.0, 1, 0, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aule_bwd_dkvt_sm100_bf16_d128 -s 1 -c 1 -o $OUT/bwd_dkvt_d128 -f python /tmp/one_bwd.py > $OUT/ncu_dkvt128.log 2>&1; echo "ncu dkvt128 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aule_bwd_dkvt_sm100_bf16_d64 -s 1 -c 1 -o $OUT/bwd_dkvt_d64 -f python /tmp/one_bwd.py > $OUT/ncu_dkvt64.log 2>&1; echo "ncu dkvt64 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:aule_bwd_delta_bf16 -s 1 -c 1 -o $OUT/bwd_delta -f python /tmp/one_bwd.py > $OUT/ncu_delta.log 2>&1; echo "ncu delta rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-secondary > $OUT/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
ls -la $OUT
