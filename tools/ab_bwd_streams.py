"""A/B on the GPU box: backward with the dQ kernel on a second stream (default) against both kernels back to back on the
caller's stream (aule_set_kernel_path bit 16).  Interleaved rounds, medians.  usage: python tools/ab_bwd_streams.py"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()
for name, (B, Hq, Hkv, S, D) in {"C/2": (4, 32, 8, 4096, 128), "B": (4, 32, 32, 2048, 64), "E": (2, 16, 16, 1024, 64)}.items():
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    o, lse = cuda_flash.forward_with_lse(q, k, v, causal=True)
    do = torch.randn_like(o)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    stream = torch.cuda.current_stream().cuda_stream

    def call():
        rc = lib.aule_attention_backward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                              dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, Hq, Hkv, S, S, D, ffi.DTYPE_BF16, 0.0, 1, 0, stream)
        assert rc == 0, ffi.last_error()
    res = {}
    times = {0: [], 65536: []}
    for path in times:
        lib.aule_set_kernel_path(path)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        res[path] = (dq.clone(), dk.clone(), dv.clone())
    for r in range(7):
        for path in times:
            lib.aule_set_kernel_path(path)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                call()
            e1.record()
            torch.cuda.synchronize()
            times[path].append(e0.elapsed_time(e1) / 10)
    lib.aule_set_kernel_path(0)
    fl = 2.5 * 4.0 * B * Hq * D * (S * (S + 1) / 2)
    same = all(torch.equal(a, b) for a, b in zip(res[0], res[65536]))
    for path, label in ((0, "two streams"), (65536, "one stream ")):
        m = statistics.median(times[path])
        print(f"{name:4s} {label}: median {m:.4f} ms  {fl / m / 1e9:7.1f} TFLOP/s   best {min(times[path]):.4f} ms | gradients identical: {same}", flush=True)
