"""Paged-KV decode throughput on the GPU box: achieved HBM GB/s of aule_paged_sm100_* against the
measured copy bandwidth in MEASURED_PEAKS.json.  Algorithmic bytes per launch = live K+V bytes
(sum(context_lens) * Hkv * D * 2 tensors * 2 B) + q + out; the cache is sized well past the 126 MB L2 and pages are
shuffled, so every launch streams from HBM.  usage: python tools/bench_paged.py [reps]"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
import aule  # noqa: E402
from aule import ffi  # noqa: E402

lib = ffi.ensure_init()
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 20
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        PEAK, PEAK_SRC = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
except Exception:
    PEAK, PEAK_SRC = 6500.0, "fallback (B200_PROFILING.md)"


def run(name, B, Hq, Hkv, D, bs, ctx, window=-1, ragged=False, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(3)
    mb = (ctx + bs - 1) // bs
    nb = B * mb
    q = torch.randn(B, Hq, D, device="cuda", dtype=dtype, generator=g)
    kc = torch.randn(nb, bs, Hkv, D, device="cuda", dtype=dtype, generator=g)
    vc = torch.randn(nb, bs, Hkv, D, device="cuda", dtype=dtype, generator=g)
    bt = torch.randperm(nb, device="cuda", generator=g).reshape(B, mb).to(torch.int32)
    if ragged:
        cl = torch.randint(ctx // 4, ctx + 1, (B,), device="cuda", generator=g).to(torch.int32)
    else:
        cl = torch.full((B,), ctx, dtype=torch.int32, device="cuda")
    live = cl.clamp(max=window) if window > 0 else cl
    bytes_alg = int(live.sum().item()) * Hkv * D * 2 * 2 + 2 * B * Hq * D * 2

    def call():
        return aule.flash_attention_paged(q, kc, vc, bt, cl, window_size=window, max_context_len=ctx)

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ts = []
    inner = 10                                   # back-to-back launches per sample: host launch latency is hidden
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            call()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    t = statistics.median(ts)
    res = {"config": name, "shape": {"B": B, "Hq": Hq, "Hkv": Hkv, "D": D, "block_size": bs, "context": ctx, "window": window,
                                     "ragged": ragged}, "kv_cache_MiB": round(2 * kc.numel() * 2 / 2**20, 1),
           "ms": round(t, 4), "ms_min": round(min(ts), 4), "GBps": round(bytes_alg / t / 1e6, 1),
           "frac_of_hbm_peak": round(bytes_alg / t / 1e6 / PEAK, 3), "peak_GBps": PEAK, "peak_source": PEAK_SRC,
           "kernel": lib.aule_last_kernel().decode()}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    if os.environ.get("AULE_PAGED_SWEEP"):
        for ns in (1, 2, 3, 4, 7, 15):
            os.environ["AULE_PAGED_NSPLIT"] = str(ns)
            run(f"nsplit={ns} B=32 ctx=8192", 32, 32, 8, 128, 16, 8192)
            run(f"nsplit={ns} ragged B=64 ctx<=8192", 64, 32, 8, 128, 16, 8192, ragged=True)
        sys.exit(0)
    if os.environ.get("AULE_PAGED_ONE"):
        run("llama3-8b decode B=32 ctx=8192", 32, 32, 8, 128, 16, 8192)
        sys.exit(0)
    run("llama3-8b decode B=32 ctx=8192", 32, 32, 8, 128, 16, 8192)
    run("llama3-8b decode B=8 ctx=32768", 8, 32, 8, 128, 16, 32768)
    run("llama3-8b decode B=128 ctx=2048", 128, 32, 8, 128, 16, 2048)
    run("single sequence ctx=131072", 1, 32, 8, 128, 16, 131072)
    run("ragged B=64 ctx<=8192", 64, 32, 8, 128, 16, 8192, ragged=True)
    run("block_size 128, B=32 ctx=8192", 32, 32, 8, 128, 128, 8192)
    run("MHA d64 B=32 ctx=4096", 32, 32, 32, 64, 16, 4096)
    run("window 1024, B=64 ctx=8192", 64, 32, 8, 128, 16, 8192, window=1024)
