#!/bin/bash
mkdir -p gpurun_out/r2_tf32
timeout 900 python -m pytest tests/test_gpu_r2.py -q -m gpu -k "tf32" > gpurun_out/r2_tf32/pytest.log 2>&1
echo "pytest rc=$?"; grep -v "watchdog: block [0-9]* thread [0-9]*[1-9] " gpurun_out/r2_tf32/pytest.log | tail -30
timeout 300 python - > gpurun_out/r2_tf32/time.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, "aule-attention_b200/python")
from aule import cuda_flash
for (B,H,S,D) in [(4,32,2048,64),(8,32,4096,64),(2,8,512,64)]:
    q = torch.randn(B,H,S,D,device="cuda"); k = torch.randn_like(q); v = torch.randn_like(q)
    fl = 4.0*B*H*D*(S*(S+1)/2)
    for tf in (True, False):
        for _ in range(3): cuda_flash.forward_with_lse(q,k,v,causal=True,allow_tf32=tf)
        torch.cuda.synchronize()
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10 if tf else 2
        e0.record()
        for _ in range(n): cuda_flash.forward_with_lse(q,k,v,causal=True,allow_tf32=tf)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/n
        print(f"fp32 [{B},{H},{S},{D}] causal {'tf32 tensor cores' if tf else 'exact CUDA cores '}: {ms:.4f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    with torch.backends.cuda.sdp_kernel(enable_flash=True, enable_math=True, enable_mem_efficient=True):
        torch.backends.cuda.matmul.allow_tf32 = True
        f = torch.nn.functional.scaled_dot_product_attention
        for _ in range(3): f(q,k,v,is_causal=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5): f(q,k,v,is_causal=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/5
        print(f"     torch SDPA fp32 (allow_tf32): {ms:.4f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
PY
cat gpurun_out/r2_tf32/time.log | tail -12
