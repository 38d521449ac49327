"""A/B of a forward-kernel path bit on the GPU box: interleaved rounds, median.  usage: python tools/ab_fwd_path.py <path_bits_B> [rounds]
(path A = 0; e.g. 32768 disables the cross-item prefetch of the next work item's first Q K^T)"""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi
lib = ffi.ensure_init()
PB = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 7
for name, B, Hq, Hkv, S, D, reps in (("C [8,32(8),4096,128]", 8, 32, 8, 4096, 128, 10), ("B [4,32,2048,64]", 4, 32, 32, 2048, 64, 30),
                                     ("D/8 [1,4,32768,128]", 1, 4, 4, 32768, 128, 10), ("MHA [8,16,1024,128]", 8, 16, 16, 1024, 128, 30)):
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    fl = 4.0 * B * Hq * D * (S * (S + 1) / 2)
    ts = {0: [], PB: []}
    outs = {}
    for r in range(ROUNDS):
        for path in (0, PB):
            lib.aule_set_kernel_path(path)
            o, _ = cuda_flash.forward_with_lse(q, k, v, causal=True)
            torch.cuda.synchronize()
            outs[path] = o
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                cuda_flash.forward_with_lse(q, k, v, causal=True)
            e1.record()
            torch.cuda.synchronize()
            ts[path].append(e0.elapsed_time(e1) / reps)
    lib.aule_set_kernel_path(0)
    a, b = statistics.median(ts[0]), statistics.median(ts[PB])
    print(json.dumps({"config": name, "path0_ms": round(a, 4), "path0_tflops": round(fl / a / 1e9, 1), f"path{PB}_ms": round(b, 4),
                      f"path{PB}_tflops": round(fl / b / 1e9, 1), "bit_identical": bool(torch.equal(outs[0], outs[PB]))}), flush=True)
