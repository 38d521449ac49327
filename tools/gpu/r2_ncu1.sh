#!/bin/bash
set -u
OUT=gpurun_out/r2_ncu1; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aule_fwd_sm100_bf16_d128_e4 -c 1 -o $OUT/so_e4 -f python tools/softmax_only.py 4 > $OUT/ncu_so.log 2>&1; echo "ncu so rc=$?"; tail -3 $OUT/ncu_so.log
cat > /tmp/one_c.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "aule-attention_b200", "python"))
from aule import cuda_flash, ffi
lib = ffi.ensure_init()
g = torch.Generator(device="cuda").manual_seed(42)
B, Hq, Hkv, S, D = 8, 32, 8, 4096, 128
q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
for _ in range(4):
    cuda_flash.forward_with_lse(q, k, v, causal=True)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aule_fwd_sm100_bf16_d128 -s 2 -c 1 -o $OUT/fwd_c -f python /tmp/one_c.py > $OUT/ncu_c.log 2>&1; echo "ncu c rc=$?"; tail -3 $OUT/ncu_c.log
ls -la $OUT
