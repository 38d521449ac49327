"""Median backward time (CUDA events, device-pointer C ABI) of config C/2, config B and config E, one JSON line.
A/B between two builds: run it under different AULE_LIBRARY_PATH values, interleaved (tools/gpu/r2_s2_b.sh)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()
stream = torch.cuda.current_stream().cuda_stream
path = int(os.environ.get("AULE_PATH", "0"))
lib.aule_set_kernel_path(path)
out = {"lib": os.environ.get("AULE_LIBRARY_PATH", "default"), "path": path}
shapes = os.environ.get("AULE_SHAPES", "C_half,B,E,D96").split(",")
for name, (B, Hq, Hkv, S, D), reps in (("C_half", (4, 32, 8, 4096, 128), 15), ("B", (4, 32, 32, 2048, 64), 30), ("E", (2, 16, 16, 1024, 64), 100),
                                       ("D96", (4, 32, 8, 4096, 96), 10), ("B_gqa", (4, 32, 8, 2048, 64), 30), ("B_long", (1, 32, 32, 8192, 64), 15)):
    if name not in shapes:
        continue
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    o, lse = cuda_flash.forward_with_lse(q, k, v, causal=True)
    do = torch.randn(o.shape, device="cuda", dtype=torch.bfloat16, generator=g)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)

    def call():
        rc = lib.aule_attention_backward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                              dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, Hq, Hkv, S, S, D, ffi.DTYPE_BF16, 0.0, 1, 0, stream)
        assert rc == 0, ffi.last_error()

    for _ in range(5):
        call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    flops = 2.5 * 4 * B * Hq * D * S * (S + 1) / 2
    out[name] = {"ms_median": round(ts[len(ts) // 2], 5), "ms_min": round(ts[0], 5), "tflops_median": round(flops / ts[len(ts) // 2] / 1e9, 1),
                 "checksum": [float(dq.float().sum()), float(dk.float().sum()), float(dv.float().sum())]}
print(json.dumps(out))
