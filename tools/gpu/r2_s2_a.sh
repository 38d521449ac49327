#!/bin/bash
# session 2, call A: sanity of the rebuilt tree (full GPU suite), reduction-rate microbenchmark, per-kernel times of the
# config-B (D = 64) and config-C/2 backward from an ncu launch list
set -u
OUT=gpurun_out/r2_s2_a; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/red_rate tools/microbench/red_rate.cu && timeout 300 /tmp/red_rate > $OUT/red_rate.log 2>&1; echo "red_rate rc=$?"; cat $OUT/red_rate.log
cat > /tmp/bwd_b.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "aule-attention_b200", "python"))
import aule
for shape in ((4, 32, 32, 2048, 64), (4, 32, 8, 4096, 128), (2, 16, 16, 1024, 64)):
    B, Hq, Hkv, S, D = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g).requires_grad_()
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g).requires_grad_()
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g).requires_grad_()
    o = aule.flash_attention(q, k, v, causal=True)
    do = torch.randn_like(o)
    for _ in range(3):
        q.grad = k.grad = v.grad = None
        o.backward(do, retain_graph=True)
    torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aule_ --csv --log-file $OUT/bwd_launches.csv python /tmp/bwd_b.py > $OUT/bwd_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" $OUT/bwd_launches.csv | awk -F'","' '{print $5, $(NF)}' | tail -40
