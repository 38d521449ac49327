"""Sustained-load comparison of kernel variants: each variant runs back-to-back for `secs` seconds while
nvidia-smi samples SM clock and power; reports TFLOP/s, median clock, and FLOP per clock per SM (pipeline
efficiency independent of the power state).  usage: python tools/sustained.py [secs] [paths]"""
import os
import random
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
paths = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 18]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(42)
B, Hq, Hkv, S, D = 8, 32, 8, 4096, 128
q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
flops = 4.0 * B * Hq * D * (S * (S + 1) / 2)

samples = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                        stdout=subprocess.PIPE, text=True)


def reader():
    for line in proc.stdout:
        try:
            c, pw = line.split(",")
            samples.append((time.perf_counter(), float(c), float(pw)))
        except ValueError:
            pass


threading.Thread(target=reader, daemon=True).start()
res = {p: [] for p in paths}
order = []
for r in range(reps):
    ps = paths[:]
    random.Random(r).shuffle(ps)
    order += ps
for var in order:
    lib.aule_set_kernel_path(var)
    for _ in range(3):
        cuda_flash.forward_with_lse(q, k, v, causal=True)
    torch.cuda.synchronize()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(20):
            cuda_flash.forward_with_lse(q, k, v, causal=True)
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / n
    clk = [c for (ts, c, pw) in samples if t0 + 0.3 <= ts <= t1]
    pw = [pw for (ts, c, pw) in samples if t0 + 0.3 <= ts <= t1]
    mc = statistics.median(clk) if clk else float("nan")
    tf = flops / ms / 1e9
    res[var].append((tf, mc, statistics.median(pw) if pw else float("nan"), flops / (ms * 1e-3) / (mc * 1e6 * 148) if clk else float("nan")))
    time.sleep(0.5)
proc.terminate()
lib.aule_set_kernel_path(0)
for var in paths:
    tf = statistics.median(x[0] for x in res[var]); mc = statistics.median(x[1] for x in res[var])
    pw = statistics.median(x[2] for x in res[var]); fpc = statistics.median(x[3] for x in res[var])
    print(f"path={var:3d} sustained {tf:7.1f} TFLOP/s  sm_clock {mc:6.0f} MHz  power {pw:6.0f} W  flop/clk/SM {fpc:6.0f} ({100 * fpc / 8192:.1f}% of 8192)  runs={[round(x[0]) for x in res[var]]}", flush=True)
