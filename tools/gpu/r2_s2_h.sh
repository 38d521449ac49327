#!/bin/bash
# session 2: ncu --set full of the shipped forward kernels (config C d128, config B d64)
set -u
OUT=gpurun_out/r2_s2_h; mkdir -p $OUT
cat > /tmp/one_fwd.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "aule-attention_b200", "python"))
from aule import cuda_flash
for (B, Hq, Hkv, S, D) in ((8, 32, 8, 4096, 128), (4, 32, 32, 2048, 64)):
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    for _ in range(3):
        cuda_flash.forward_with_lse(q, k, v, causal=True)
    torch.cuda.synchronize()
PY
timeout 240 ncu --set full --clock-control none --import-source on -k regex:aule_fwd_sm100_bf16_d128 -s 1 -c 1 -o $OUT/fwd_d128 -f python /tmp/one_fwd.py > $OUT/ncu_d128.log 2>&1; echo "ncu d128 rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:aule_fwd_sm100_bf16_d64 -s 1 -c 1 -o $OUT/fwd_d64 -f python /tmp/one_fwd.py > $OUT/ncu_d64.log 2>&1; echo "ncu d64 rc=$?"
