"""Softmax-only microbenchmark of the forward kernel (tuning builds, GPU box): the variants compiled with VAR bit 256 run
ONLY the softmax warps' per-block loop (TMEM S -> max -> exp2 -> P -> TMEM) 400 times with every barrier removed and no
MMA / TMA / epilogue work, two warps per SM sub-partition as in the real kernel.  Prints cycles per 128-key block: the
floor the softmax stage puts under the tensor pipe's 2048 cycles per block pair (D=128; 1024 at D=64).
usage: python tools/softmax_only.py [variants comma-separated, default 4,5,6,7]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()
variants = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4, 5, 6, 7]
q = torch.randn(1, 2, 256, 128, device="cuda", dtype=torch.bfloat16)
k = torch.randn(1, 1, 256, 128, device="cuda", dtype=torch.bfloat16)
v = torch.randn(1, 1, 256, 128, device="cuda", dtype=torch.bfloat16)
for var in variants:
    lib.aule_set_kernel_path(16 + var)
    buf = torch.zeros(4 * 4096, dtype=torch.int64, device="cuda")
    lib.aule_set_trace_buffer(buf.data_ptr())
    res = []
    for _ in range(3):
        buf.zero_()
        cuda_flash.forward_with_lse(q, k, v, causal=True)
        torch.cuda.synchronize()
        res.append(buf[:2].cpu().tolist())
    lib.aule_set_trace_buffer(0)
    print(f"variant {var} {lib.aule_last_kernel().decode()}: cycles per block (tile0, tile1) = "
          + ", ".join(f"({a / 400:.0f}, {b / 400:.0f})" for a, b in res), flush=True)
lib.aule_set_kernel_path(0)
