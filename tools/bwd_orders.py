"""Backward timing split on the GPU box: aule_attention_backward_dptr called directly (no autograd overhead) with the
hooks of aule_set_kernel_path: mode m is the path value m << 10.
  0: whole backward       1: Delta + dK/dV kernel only       2: Delta + dQ kernel only   (bits 10-11; partial gradients)
  +8 (path bit 13): the v3 dK/dV kernel (P, dS staged through shared memory) instead of the shipped transposed v4
  32+m: mode m with path bit 9 set (L2-run CTA order hook; unused by the shipped backward)
Interleaved rounds, median.  TFLOP/s are quoted on the whole backward's 2.5x-forward FLOPs, so only the mode-0 figure
is a throughput; modes 1/2 show where the time goes.  For modes 8/16 the deviation of dq/dk/dv from mode 0 is reported.
usage: AULE_BWD_ORDERS=0,1,2,8 python tools/bwd_orders.py [rounds]"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()
ROUNDS = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ORDERS = [int(x) for x in os.environ.get("AULE_BWD_ORDERS", "0,1,2").split(",")]
# modes 32+m: mode m with the L2-run CTA order of the dK/dV kernel disabled (path bit 9), for A/B
RAW_OR = {32 + m: 512 - (32 << 10) for m in (0, 1, 2)}


def run(name, B, Hq, Hkv, S, D, reps):
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    o, lse = cuda_flash.forward_with_lse(q, k, v, causal=True)
    do = torch.randn_like(o)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    stream = torch.cuda.current_stream().cuda_stream
    fl = 2.5 * 4.0 * B * Hq * D * (S * (S + 1) / 2)

    def call():
        rc = lib.aule_attention_backward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(),
                                              lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, Hq, Hkv, S, S, D,
                                              ffi.DTYPE_BF16, 0.0, 1, 0, stream)
        assert rc == 0, ffi.last_error()

    times = {o_: [] for o_ in ORDERS}
    ref = None
    same = True
    dev = {}
    for r in range(ROUNDS):
        for o_ in ORDERS:
            lib.aule_set_kernel_path((o_ << 10) + RAW_OR.get(o_, 0))
            call()
            torch.cuda.synchronize()
            if o_ in (8, 16) and ref is not None and r == 0:   # polynomial-exp2 variants: deviation from the MUFU result
                for nm, a, b_ in zip(("dq", "dk", "dv"), ref, (dq, dk, dv)):
                    dev[f"emu{o_ // 8}_{nm}_rel"] = round(((a.float() - b_.float()).abs().max() / a.float().abs().max()).item(), 6)
            if o_ == 0:                                  # the whole backward is bit-reproducible run to run
                cur = (dq.clone(), dk.clone(), dv.clone())
                if ref is None:
                    ref = cur
                else:
                    same = same and all(torch.equal(a, b) for a, b in zip(ref, cur))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                call()
            e1.record()
            torch.cuda.synchronize()
            times[o_].append(e0.elapsed_time(e1) / reps)
    lib.aule_set_kernel_path(0)
    res = {"config": name, "bit_identical_run_to_run": same, **dev}
    for o_ in ORDERS:
        t = statistics.median(times[o_])
        res[f"mode{o_}_ms"] = round(t, 4)
        if o_ == 0:
            res["tflops"] = round(fl / t / 1e9, 1)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    run("C/2 bwd bf16 GQA [4,32,4096,128]", 4, 32, 8, 4096, 128, 5)
    run("B bwd bf16 MHA [4,32,2048,64]", 4, 32, 32, 2048, 64, 10)
    run("E bwd bf16 [2,16,1024,64]", 2, 16, 16, 1024, 64, 20)
