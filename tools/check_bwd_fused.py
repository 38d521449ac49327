"""Fused backward kernel (aule_set_kernel_path bit 17) against the deterministic two-kernel backward and a torch fp32
reference, then an interleaved A/B timing on config C/2.  usage: python tools/check_bwd_fused.py [--no-time] [--units 4,8,16]"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

FUSED = (1 << int(os.environ.get("AULE_FUSED_BIT", "17"))) | int(os.environ.get("AULE_EXTRA_PATH", "0"))      # 17: first fused kernel, 26: 64-query half steps + TMA bulk reductions
lib = ffi.ensure_init()


def ref_grads(q, k, v, do, causal):
    qf, kf, vf = (t.float().detach().requires_grad_() for t in (q, k, v))
    g = q.shape[1] // k.shape[1]
    kk, vv = kf.repeat_interleave(g, 1), vf.repeat_interleave(g, 1)
    s = qf @ kk.transpose(-1, -2) / q.shape[-1] ** 0.5
    if causal:
        Sq, Sk = q.shape[2], k.shape[2]
        m = torch.arange(Sq, device=q.device)[:, None] >= torch.arange(Sk, device=q.device)[None, :]
        s = s.masked_fill(~m, float("-inf"))
    o = torch.softmax(s, -1) @ vv
    o.backward(do.float())
    return qf.grad, kf.grad, vf.grad


def backward(q, k, v, o, lse, do, causal, path, dtype_code):
    B, Hq, Sq, D = q.shape
    Hkv, Sk = k.shape[1], k.shape[2]
    dq, dk, dv = torch.full_like(q, float("nan")), torch.full_like(k, float("nan")), torch.full_like(v, float("nan"))
    lib.aule_set_kernel_path(path)
    rc = lib.aule_attention_backward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                          dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, Hq, Hkv, Sq, Sk, D, dtype_code, 0.0,
                                          1 if causal else 0, 0, torch.cuda.current_stream().cuda_stream)
    lib.aule_set_kernel_path(0)
    assert rc == 0, ffi.last_error()
    torch.cuda.synchronize()
    return dq, dk, dv


def check():
    cases = [  # B, Hq, Hkv, Sq, Sk, causal, dtype
        (1, 1, 1, 128, 128, True, torch.bfloat16),
        (1, 2, 1, 256, 256, True, torch.bfloat16),
        (2, 4, 2, 512, 512, True, torch.bfloat16),
        (1, 4, 1, 384, 384, False, torch.bfloat16),
        (1, 2, 2, 200, 200, True, torch.bfloat16),
        (2, 4, 2, 1000, 1000, True, torch.float16),
        (1, 2, 1, 300, 520, False, torch.bfloat16),
        (1, 8, 2, 2048, 2048, True, torch.bfloat16),
    ]
    ok = True
    for (B, Hq, Hkv, Sq, Sk, causal, dt) in cases:
        g = torch.Generator(device="cuda").manual_seed(Sq + Hq)
        q = torch.randn(B, Hq, Sq, 128, device="cuda", dtype=dt, generator=g)
        k = torch.randn(B, Hkv, Sk, 128, device="cuda", dtype=dt, generator=g)
        v = torch.randn(B, Hkv, Sk, 128, device="cuda", dtype=dt, generator=g)
        o, lse = cuda_flash.forward_with_lse(q, k, v, causal=causal)
        do = torch.randn_like(o)
        code = ffi.DTYPE_BF16 if dt == torch.bfloat16 else ffi.DTYPE_F16
        two = backward(q, k, v, o, lse, do, causal, 0, code)
        fus = backward(q, k, v, o, lse, do, causal, FUSED, code)
        assert lib.aule_last_kernel().decode().startswith("aule_bwd_dq_convert"), lib.aule_last_kernel()
        ref = ref_grads(q, k, v, do, causal)
        line = f"[{B},{Hq}({Hkv}),{Sq}/{Sk},128] causal={causal} {str(dt)[6:]}:"
        for name, a, b_, r in zip(("dq", "dk", "dv"), fus, two, ref):
            scale = r.abs().max().item()
            e_f = (a.float() - r).abs().max().item() / scale
            e_t = (b_.float() - r).abs().max().item() / scale
            good = bool(torch.isfinite(a).all()) and e_f <= max(2.0 * e_t, 1e-2)
            ok &= good
            line += f"  {name} fused {e_f:.2e} two-kernel {e_t:.2e} {'ok' if good else 'FAIL'}{' ==' if torch.equal(a, b_) else ''}"
        print(line, flush=True)
    return ok


def timing(units):
    B, Hq, Hkv, S, D = 4, 32, 8, 4096, 128
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
    paths = {"two-kernel": 0}
    for u in units:
        paths[f"fused U={u}"] = FUSED | (u << 18)
    times = {n: [] for n in paths}
    for n, pth in paths.items():
        lib.aule_set_kernel_path(pth)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
    for r in range(5):
        for n, pth in paths.items():
            lib.aule_set_kernel_path(pth)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                call()
            e1.record()
            torch.cuda.synchronize()
            times[n].append(e0.elapsed_time(e1) / 10)
    lib.aule_set_kernel_path(0)
    fl = 2.5 * 4.0 * B * Hq * D * (S * (S + 1) / 2)
    for n in paths:
        m = statistics.median(times[n])
        print(f"C/2 {n:14s}: median {m:.4f} ms  {fl / m / 1e9:7.1f} TFLOP/s   best {min(times[n]):.4f} ms", flush=True)


if __name__ == "__main__":
    good = check()
    print("CHECK", "PASS" if good else "FAIL", flush=True)
    if "--no-time" not in sys.argv and good:
        units = [4, 8, 16, 0]
        if "--units" in sys.argv:
            units = [int(x) for x in sys.argv[sys.argv.index("--units") + 1].split(",")]
        timing(units)
