"""dQ kernel v2 (default) against v1 (aule_set_kernel_path bit 23, tuning builds) and a torch fp32 reference, then an interleaved
A/B timing of the whole backward on configs C/2, B and E.  usage: python tools/check_bwd_dq2.py"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from check_bwd_fused import backward, ref_grads, lib, ffi, cuda_flash  # noqa: E402

V1 = 1 << 23
cases = [  # B, Hq, Hkv, Sq, Sk, D, causal, dtype
    (1, 1, 1, 128, 128, 128, True, torch.bfloat16), (2, 4, 2, 512, 512, 128, True, torch.bfloat16), (1, 4, 1, 384, 384, 64, False, torch.bfloat16),
    (1, 2, 2, 200, 200, 64, True, torch.bfloat16), (2, 4, 2, 1000, 1000, 128, True, torch.float16), (1, 2, 1, 300, 520, 128, False, torch.bfloat16),
    (1, 8, 2, 2048, 2048, 128, True, torch.bfloat16), (3, 4, 4, 520, 520, 64, True, torch.float16), (1, 2, 2, 700, 130, 64, False, torch.bfloat16)]
ok = True
for (B, Hq, Hkv, Sq, Sk, D, causal, dt) in cases:
    g = torch.Generator(device="cuda").manual_seed(Sq + Hq)
    q = torch.randn(B, Hq, Sq, D, device="cuda", dtype=dt, generator=g)
    k = torch.randn(B, Hkv, Sk, D, device="cuda", dtype=dt, generator=g)
    v = torch.randn(B, Hkv, Sk, D, device="cuda", dtype=dt, generator=g)
    o, lse = cuda_flash.forward_with_lse(q, k, v, causal=causal)
    do = torch.randn_like(o)
    code = ffi.DTYPE_BF16 if dt == torch.bfloat16 else ffi.DTYPE_F16

    def bw(path):
        dq, dk, dv = torch.full_like(q, float("nan")), torch.full_like(k, float("nan")), torch.full_like(v, float("nan"))
        lib.aule_set_kernel_path(path)
        rc = lib.aule_attention_backward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                              dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, Hq, Hkv, Sq, Sk, D, code, 0.0,
                                              1 if causal else 0, 0, torch.cuda.current_stream().cuda_stream)
        lib.aule_set_kernel_path(0)
        assert rc == 0, ffi.last_error()
        torch.cuda.synchronize()
        return dq, dk, dv
    new, old = bw(0), bw(V1)
    ref = ref_grads(q, k, v, do, causal)
    scale = ref[0].abs().max().item()
    e_n = (new[0].float() - ref[0]).abs().max().item() / scale
    e_o = (old[0].float() - ref[0]).abs().max().item() / scale
    good = bool(torch.isfinite(new[0]).all()) and e_n <= max(1.5 * e_o, 1e-2)
    ok &= good
    print(f"[{B},{Hq}({Hkv}),{Sq}/{Sk},{D}] causal={causal} {str(dt)[6:]}: dq v2 {e_n:.2e} v1 {e_o:.2e} {'ok' if good else 'FAIL'}"
          f"{' == v1' if torch.equal(new[0], old[0]) else ''}", flush=True)
print("CHECK", "PASS" if ok else "FAIL", flush=True)
if ok:
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
        paths = {"dQ v2 (two streams)": 0, "dQ v1 + old polls": V1 | (1 << 24), "dQ v2 only": 2 << 10, "dQ v1 only": V1 | (2 << 10), "dK/dV only": 1 << 10,
                 "dK/dV only, old polls": (1 << 10) | (1 << 24)}
        times = {n_: [] for n_ in paths}
        for n_, pth in paths.items():
            lib.aule_set_kernel_path(pth)
            for _ in range(3):
                call()
            torch.cuda.synchronize()
        for r in range(5):
            for n_, pth in paths.items():
                lib.aule_set_kernel_path(pth)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    call()
                e1.record()
                torch.cuda.synchronize()
                times[n_].append(e0.elapsed_time(e1) / 10)
        lib.aule_set_kernel_path(0)
        fl = 2.5 * 4.0 * B * Hq * D * (S * (S + 1) / 2)
        for n_ in paths:
            m = statistics.median(times[n_])
            print(f"{name:4s} {n_:22s}: median {m:.4f} ms  {fl / m / 1e9:7.1f} TFLOP/s (whole-backward FLOPs)  best {min(times[n_]):.4f} ms", flush=True)
