#!/usr/bin/env python
"""bench.py -- attention forward TFLOPS & % of tensor-core peak on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C]

A "step" is one pass of the hot path (fused attention forward) over one synthetic batch.
Default workload = BASELINE.json configs[2], the configuration the metric is quoted on
(S=4096, D=128, bf16): GQA 32q/8kv [8,32,4096,128] causal, PER GPU (weak scaling: the
batch x heads space shards with no data-path collective).

  value         whole-job TFLOP/s, inputs resident in HBM, K back-to-back launches between
                CUDA events on the launch stream, max over ranks.  FLOPs are causal-exact:
                4*B*Hq*D*Sq(Sq+1)/2; the reference's 4BHS^2D convention is
                reported beside it in config.
  e2e           same metric through the C-ABI host-buffer call (aule_attention_forward_host)
                from PINNED HOST memory: H2D of q,k,v and D2H of o inside the timed region.
  roofline      tensor-bound: achieved = FLOPs per launch / mean launch time; peak = measured
                cuBLAS bf16 (MEASURED_PEAKS.json), else the profiling guide's fallback.
  cpu_baseline  the reference's NumPy path (oracle port of python/aule/__init__.py:247-271)
                on the host cores over a bounded sample of (batch, head) slices of the same
                workload; rank 0, N=1 only.
  --impl reference   times that CPU path alone, all host cores, same metric/config.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "aule-attention_b200", "python")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: (B, Hq, Hkv, S, D, description)
    "B": (4, 32, 32, 2048, 64, "bf16 causal MHA [4,32,2048,64] (BASELINE.json configs[1])"),
    "C": (8, 32, 8, 4096, 128, "bf16 GQA 32q/8kv [8,32,4096,128] causal, Llama-3-8B shape (BASELINE.json configs[2])"),
    "D": (1, 32, 32, 32768, 128, "bf16 causal long-seq [1,32,32768,128] (BASELINE.json configs[3]), heads sharded across ranks"),
}
METRIC = "attention fwd TFLOPS at S=4096 D=128 bf16 (causal-exact FLOPs)"


def causal_flops(B, Hq, S, D):
    return 4.0 * B * Hq * D * (S * (S + 1) / 2.0)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 1590.0, "fallback (B200_PROFILING.md: 1.59 PFLOP/s)"


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, gpu_index):
        self.idx, self.samples, self.proc, self.thread = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz, mx, reasons, power = [], None, set(), []
        for ts, line in self.samples:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk = float(f[0]); mx = float(f[1])
            except ValueError:
                continue
            in_region = (t0 is None or (ts >= t0 - 0.05 and ts <= t1 + 0.15))
            if in_region:
                mhz.append(clk)
                try:
                    power.append(float(f[6]))
                except ValueError:
                    pass
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------ CPU reference arm
def _cpu_slice_worker(args):
    """One (batch, head) slice [1,1,S,D] fp32 through the reference NumPy path (oracle port)."""
    S, D, seed = args
    import numpy as np
    from oracle.attention_oracle import cpu_attention
    rng = np.random.RandomState(seed)
    q, k, v = (rng.randn(1, 1, S, D).astype(np.float32) for _ in range(3))
    t = time.perf_counter()
    o = cpu_attention(q, k, v, causal=True)
    return time.perf_counter() - t, float(o[0, 0, -1, 0])


def cpu_reference_run(S, D, steps, warmup, budget_s, max_workers=None):
    """Times `steps` steps; a step = `nslices` slices processed by a pool of host processes."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    workers = max(1, min(cores, max_workers or 32))
    t_slice, _ = _cpu_slice_worker((S, D, 0))                       # calibration, 1 core
    per_step = budget_s / max(1, steps + warmup)
    rounds = max(1, min(4, int(per_step / max(t_slice * 1.5, 1e-3))))
    nslices = workers * rounds
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for w in range(warmup):
            pool.map(_cpu_slice_worker, [(S, D, 1000 + i) for i in range(nslices)])
        times = []
        for s in range(steps):
            t0 = time.perf_counter()
            pool.map(_cpu_slice_worker, [(S, D, 2000 + s * nslices + i) for i in range(nslices)])
            times.append(time.perf_counter() - t0)
    total = sum(times)
    flops = causal_flops(1, 1, S, D) * nslices * steps
    return {"tflops": flops / total / 1e12, "ms_per_step": 1e3 * total / steps, "cores": workers, "nslices": nslices,
            "t_slice_1core_s": t_slice, "cores_available": cores}


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="aule", choices=["aule", "reference"])
    ap.add_argument("--workload", default="C", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B, Hq, Hkv, S, D, desc = WORKLOADS[args.workload]
    sharded_heads = args.workload == "D"
    if sharded_heads:                      # strong scaling: heads split across ranks
        assert Hq % world == 0
        Hq_r, Hkv_r = Hq // world, Hkv // world
    else:                                  # weak scaling: the full config per GPU
        Hq_r, Hkv_r = Hq, Hkv
    flops_rank = causal_flops(B, Hq_r, S, D)
    config = {"workload": desc, "per_gpu_shape": {"q": [B, Hq_r, S, D], "kv": [B, Hkv_r, S, D]}, "causal": True,
              "parallelism": f"batch x heads sharded over {world} GPU(s), no data-path collective",
              "l2": "inputs (>= 640 MiB per GPU) exceed the 126 MB L2; no flush needed",
              "flops_convention": "causal-exact 4*B*Hq*D*S(S+1)/2",
              "flops_per_step_per_gpu": flops_rank, "reference_convention_4BHS2D": 4.0 * B * Hq_r * S * S * D}

    if args.impl == "reference":
        # The reference's own CPU implementation of the path (NumPy, python/aule/__init__.py:247-271),
        # restated in oracle/attention_oracle.py::cpu_attention; rank 0 only.
        if rank != 0:
            return
        r = cpu_reference_run(S, D, args.steps, args.warmup, budget_s=150.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["tflops"], "unit": "TFLOP/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if sharded_heads else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": r["tflops"], "unit": "TFLOP/s", "cores": r["cores"], "kind": "port",
                                 "sample": f"{r['nslices']} (batch,head) slices [1,1,{S},{D}] fp32 causal per step, one per "
                                           f"process over {r['cores']} of {r['cores_available']} host cores; the NumPy path is "
                                           f"single-threaded per slice ({r['t_slice_1core_s']:.2f} s/slice on 1 core)"},
                "e2e": {"value": r["tflops"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import aule
    from aule import cuda_flash, ffi
    if aule.get_available_backends() != ["cuda"]:
        raise SystemExit("aule CUDA backend unavailable: " + str(aule.get_backend_errors()))
    lib = ffi.load_library()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        # the image exports NCCL_DEBUG=VERSION, which makes NCCL print a banner on stdout; stdout is ONE JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    g = torch.Generator(device=dev).manual_seed(42 + rank)
    q = torch.randn(B, Hq_r, S, D, device=dev, dtype=torch.bfloat16, generator=g)
    k = torch.randn(B, Hkv_r, S, D, device=dev, dtype=torch.bfloat16, generator=g)
    v = torch.randn(B, Hkv_r, S, D, device=dev, dtype=torch.bfloat16, generator=g)
    o = torch.empty_like(q)
    lse = torch.empty(B, Hq_r, S, device=dev, dtype=torch.float32)
    stream = torch.cuda.current_stream(dev)

    def launch():
        rc = lib.aule_attention_forward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
                                             B, Hq_r, Hkv_r, S, S, D, ffi.DTYPE_BF16, 0.0, 1, -1, local_rank,
                                             stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(ffi.last_error())

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        launch()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    n0 = lib.aule_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_start = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        launch()
    e1.record(stream)
    sync_all()
    t_end = time.perf_counter()
    launches = lib.aule_launch_count() - n0
    kernel = lib.aule_last_kernel().decode()
    ms_total = e0.elapsed_time(e1)
    # keep the GPU busy a little longer so the clock sampler sees load even for short regions
    t_busy = time.perf_counter()
    while time.perf_counter() - t_busy < 1.0:
        launch()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop(t_start, time.perf_counter())
    if dist is not None:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * flops_rank / (ms_step * 1e-3) / 1e12

    # ---------------- e2e: host buffers through the C ABI, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        hq = torch.empty(q.shape, dtype=torch.bfloat16).pin_memory(); hq.copy_(q)
        hk = torch.empty(k.shape, dtype=torch.bfloat16).pin_memory(); hk.copy_(k)
        hv = torch.empty(v.shape, dtype=torch.bfloat16).pin_memory(); hv.copy_(v)
        ho = torch.empty(q.shape, dtype=torch.bfloat16).pin_memory()

        def host_step():
            rc = lib.aule_attention_forward_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), None,
                                                 B, Hq_r, Hkv_r, S, S, D, ffi.DTYPE_BF16, 0.0, 1, -1, local_rank)
            if rc != 0:
                raise RuntimeError(ffi.last_error())
        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            host_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            host_step()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert torch.equal(ho.to(dev), o), "host-buffer path and device-pointer path disagree"
        e2e = {"value": world * flops_rank * e_steps / dt / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": world * 2 * (q.numel() + k.numel() + v.numel()),
               "d2h_bytes_per_step": world * 2 * q.numel(), "steps": e_steps, "ms_per_step": 1e3 * dt / e_steps,
               "api": "aule_attention_forward_host (pinned host buffers, chunked H2D/compute/D2H overlap on three streams)"}

    if rank != 0:
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    per_launch_tflops = flops_rank / (ms_step * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "fwd_traffic_bytes.json")
    if os.path.exists(tpath) and args.workload == "C":
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    line = {"metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if sharded_heads else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "kernel": kernel,
            "roofline": {"bound": "tensor", "achieved": per_launch_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": per_launch_tflops / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_flops_per_launch": flops_rank,
                         "compulsory_hbm_bytes_per_launch": 2 * (2 * q.numel() + k.numel() + v.numel())},
            "pct_of_nominal_2250_tflops": 100.0 * per_launch_tflops / 2250.0}
    if args.gpus == 1 and not args.no_cpu:
        r = cpu_reference_run(S, D, steps=1, warmup=0, budget_s=20.0, max_workers=1)
        line["cpu_baseline"] = {"value": r["tflops"], "unit": "TFLOP/s", "cores": 1, "kind": "port",
                                "sample": f"{r['nslices']} (batch,head) slice(s) [1,1,{S},{D}] fp32 causal of the same workload "
                                          f"through the reference NumPy path (single-threaded einsum); full step = {B * Hq_r} slices"}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
