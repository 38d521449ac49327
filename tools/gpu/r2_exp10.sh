#!/bin/bash
# D=64 forward with the row sum folded into the P V MMA: correctness, A/B, ncu summary
set -u
OUT=gpurun_out/r2_exp10; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sm100.py tests/test_gpu_r2.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
AULE_SWEEP_CFG=B timeout 300 python tools/sweep_variants.py 30 7 0,16384 2>&1 | tee $OUT/sweep_B.log
AULE_SWEEP_CFG=C timeout 300 python tools/sweep_variants.py 20 5 0,16384 2>&1 | tee $OUT/sweep_C.log
cat > /tmp/one_b.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "aule-attention_b200", "python"))
from aule import cuda_flash, ffi
lib = ffi.ensure_init()
g = torch.Generator(device="cuda").manual_seed(42)
B, Hq, Hkv, S, D = 4, 32, 32, 2048, 64
q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
for _ in range(4):
    cuda_flash.forward_with_lse(q, k, v, causal=True)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aule_fwd_sm100_bf16_d64 -s 2 -c 1 -o $OUT/fwd_b -f python /tmp/one_b.py > $OUT/ncu_b.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu_b.log
