"""Secondary measurements on the GPU box (context, not the headline): backward throughput of this engine,
other configs of BASELINE.json, and yardsticks available in the image (torch SDPA, flash_attn 2.8) on identical
inputs.  FLOPs: causal-exact 4*B*Hq*D*S(S+1)/2 forward, x2.5 backward."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
import aule  # noqa: E402
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()


def timeit(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


_REF_SCRIPT = r"""
import json, sys, torch
sys.path.insert(0, sys.argv[1])
import aule                                       # the REFERENCE package (baseline/_ref), unmodified
from aule.triton_flash import flash_attention_triton
B, Hq, Hkv, S, D = (int(x) for x in sys.argv[2:7])
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
with torch.no_grad():
    for _ in range(3):
        flash_attention_triton(q, k, v, causal=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        flash_attention_triton(q, k, v, causal=True)
    e1.record()
    torch.cuda.synchronize()
print("MS", e0.elapsed_time(e1) / 5)
"""


def reference_triton(B, Hq, Hkv, S, D, flops):
    """The reference's own GPU path (python/aule/triton_flash.py:529 flash_attention_triton), recompiled by Triton for this
    GPU, as the yardstick BASELINE.md 4 / SURVEY 2.3 ask for.  Needs the pip-installed reference under baseline/_ref."""
    import subprocess
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "aule")):
        return "unavailable: baseline/_ref not installed"
    try:
        env = dict(os.environ); env.pop("PYTHONPATH", None)
        p = subprocess.run([sys.executable, "-c", _REF_SCRIPT, ref, str(B), str(Hq), str(Hkv), str(S), str(D)],
                           capture_output=True, text=True, timeout=900, env=env)
        ms = [float(ln.split()[1]) for ln in p.stdout.splitlines() if ln.startswith("MS ")]
        if not ms:
            return "failed: " + (p.stderr.strip().splitlines() or ["no output"])[-1][:200]
        return flops / ms[-1] / 1e9
    except Exception as e:
        return f"failed: {type(e).__name__}"


def run(name, B, Hq, Hkv, S, D, dtype=torch.bfloat16, bwd=False, yard=True):
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=dtype, generator=g)
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=dtype, generator=g)
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=dtype, generator=g)
    fl = 4.0 * B * Hq * D * (S * (S + 1) / 2)
    res = {"config": name, "shape": [B, Hq, Hkv, S, D], "dtype": str(dtype)}
    t = timeit(lambda: cuda_flash.forward_with_lse(q, k, v, causal=True))
    res["aule_fwd_ms"], res["aule_fwd_tflops"] = t, fl / t / 1e9
    res["aule_fwd_kernel"] = lib.aule_last_kernel().decode()
    if bwd:
        qg, kg, vg = (x.clone().requires_grad_() for x in (q, k, v))
        o = aule.flash_attention(qg, kg, vg, causal=True)
        do = torch.randn_like(o)

        def step():
            qg.grad = kg.grad = vg.grad = None
            o.backward(do, retain_graph=True)
        t = timeit(step, steps=5, warm=2)
        res["aule_bwd_ms"], res["aule_bwd_tflops"] = t, 2.5 * fl / t / 1e9
        lib.aule_set_kernel_path(1)
        try:
            t = timeit(step, steps=2, warm=1)
        finally:
            lib.aule_set_kernel_path(0)
        res["aule_bwd_cudacore_ms"], res["aule_bwd_cudacore_tflops"] = t, 2.5 * fl / t / 1e9
    if yard:
        try:
            t = timeit(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=(Hq != Hkv)))
            res["torch_sdpa_fwd_tflops"] = fl / t / 1e9
        except Exception as e:
            res["torch_sdpa_fwd_tflops"] = f"unavailable: {type(e).__name__}"
        res["reference_triton_fwd_tflops"] = reference_triton(B, Hq, Hkv, S, D, fl)
        try:
            from flash_attn import flash_attn_func
            qf, kf, vf = (x.transpose(1, 2).contiguous() for x in (q, k, v))
            t = timeit(lambda: flash_attn_func(qf, kf, vf, causal=True))
            res["flash_attn_2_8_fwd_tflops"] = fl / t / 1e9
        except Exception as e:
            res["flash_attn_2_8_fwd_tflops"] = f"unavailable: {type(e).__name__}"
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    only_bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
    if not only_bwd:
        run("B bf16 MHA [4,32,2048,64]", 4, 32, 32, 2048, 64)
        run("C bf16 GQA [8,32,4096,128]", 8, 32, 8, 4096, 128)
        run("D/8 bf16 [1,4,32768,128] (one GPU's head shard of config D)", 1, 4, 4, 32768, 128, yard=False)
    run("E bf16 fwd+bwd [2,16,1024,64]", 2, 16, 16, 1024, 64, bwd=True, yard=not only_bwd)
    run("C/2 bwd bf16 GQA [4,32,4096,128]", 4, 32, 8, 4096, 128, bwd=True, yard=False)
    if not only_bwd:
        run("C fp16 GQA [8,32,4096,128]", 8, 32, 8, 4096, 128, dtype=torch.float16, yard=False)
