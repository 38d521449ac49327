"""Time the tuning variants of the headline kernel on config C and check each against the
CUDA-core kernel (GPU box). usage: python tools/sweep_variants.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = torch.Generator(device="cuda").manual_seed(42)
B, Hq, Hkv, S, D = 8, 32, 8, 4096, 128
q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
flops = 4.0 * B * Hq * D * (S * (S + 1) / 2)
lib.aule_set_kernel_path(1)
ref, ref_lse = cuda_flash.forward_with_lse(q[:1, :, :1024], k[:1, :, :1024], v[:1, :, :1024], causal=True)
res = {}
for var in [16, 17, 18, 19, 0]:
    lib.aule_set_kernel_path(var)
    o, lse = cuda_flash.forward_with_lse(q[:1, :, :1024].contiguous(), k[:1, :, :1024].contiguous(), v[:1, :, :1024].contiguous(), causal=True)
    err = (o.float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
    lerr = (lse - ref_lse).abs().max().item()
    for _ in range(5):
        cuda_flash.forward_with_lse(q, k, v, causal=True)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            cuda_flash.forward_with_lse(q, k, v, causal=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    name = lib.aule_last_kernel().decode()
    print(f"path={var:2d} {name:32s} {best:.4f} ms  {flops / best / 1e9:8.1f} TFLOP/s  rel_err_vs_cudacore={err:.2e} lse_err={lerr:.2e}", flush=True)
lib.aule_set_kernel_path(0)
