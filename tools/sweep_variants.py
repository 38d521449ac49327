"""Interleaved A/B timing of the tuning variants of the headline kernel on config C, each checked
against the CUDA-core kernel (GPU box).  Variants are measured round-robin so that power/thermal
drift hits all of them equally; the median over rounds is reported.
usage: python tools/sweep_variants.py [steps] [rounds] [paths comma-separated]"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
paths = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 16, 17, 18]
g = torch.Generator(device="cuda").manual_seed(42)
B, Hq, Hkv, S, D = {"B": (4, 32, 32, 2048, 64), "D8": (1, 4, 4, 32768, 128), "C": (8, 32, 8, 4096, 128)}[os.environ.get("AULE_SWEEP_CFG", "C")]
q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
flops = 4.0 * B * Hq * D * (S * (S + 1) / 2)
qs, ks, vs = (t[:1, :, :1024].contiguous() for t in (q, k, v))
lib.aule_set_kernel_path(1)
ref, ref_lse = cuda_flash.forward_with_lse(qs, ks, vs, causal=True)
names, errs = {}, {}
for var in paths:
    lib.aule_set_kernel_path(var)
    o, lse = cuda_flash.forward_with_lse(qs, ks, vs, causal=True)
    names[var] = lib.aule_last_kernel().decode()
    errs[var] = ((o.float() - ref.float()).abs().max().item() / ref.float().abs().max().item(), (lse - ref_lse).abs().max().item())
    for _ in range(3):
        cuda_flash.forward_with_lse(q, k, v, causal=True)
torch.cuda.synchronize()
times = {p: [] for p in paths}
for r in range(rounds):
    for var in paths:
        lib.aule_set_kernel_path(var)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            cuda_flash.forward_with_lse(q, k, v, causal=True)
        e1.record()
        torch.cuda.synchronize()
        times[var].append(e0.elapsed_time(e1) / steps)
for var in paths:
    med, best = statistics.median(times[var]), min(times[var])
    print(f"path={var:2d} {names[var]:30s} median {med:.4f} ms {flops / med / 1e9:7.1f} TFLOP/s | best {best:.4f} ms "
          f"{flops / best / 1e9:7.1f} | rel_err={errs[var][0]:.2e} lse_err={errs[var][1]:.2e}", flush=True)
lib.aule_set_kernel_path(0)
