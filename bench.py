#!/usr/bin/env python
"""bench.py -- attention forward TFLOPS & % of tensor-core peak on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C] [--no-secondary] [--no-e2e] [--no-cpu]

A "step" is one pass of the hot path (fused attention forward) over one synthetic batch.
Default workload = BASELINE.json configs[2], the configuration the metric is quoted on
(S=4096, D=128, bf16): GQA 32q/8kv [8,32,4096,128] causal, PER GPU (weak scaling: the
batch x heads space shards with no data-path collective).

  value         whole-job TFLOP/s, inputs resident in HBM, K back-to-back launches between CUDA events on the launch
                stream, max over ranks.  FLOPs are causal-exact: 4*B*Hq*D*Sq(Sq+1)/2; the reference's
                4BHS^2D convention is reported beside it in config.
  e2e           same metric through the C-ABI host-buffer call (aule_attention_forward_host) from PINNED HOST memory
                (allocated with the rank's CPU affinity set to the GPU's local cores): H2D of q,k,v and D2H of o inside
                the timed region.  `copy_ceiling` beside it = the same bytes moved by bare, overlapped cuMemcpy calls
                on all ranks at once: what the host side of the box allows.
  roofline      tensor-bound: achieved = FLOPs per launch / mean launch time; peak = measured cuBLAS bf16
                (MEASURED_PEAKS.json: burst, with the sustained fraction beside it), else the profiling guide's fallback.
                traffic = DRAM bytes per launch from this round's ncu capture (profiles/fwd_traffic_bytes.json), or null.
  secondary     measured in the same run (rank-local unless stated): config B forward, backward (config C/2 and E),
                forward+backward end to end from host buffers (config E), paged-KV decode GB/s, config D (long
                sequence, heads sharded over the ranks: strong scaling), the native spanning call (one call whose
                tensors live on GPU 0, computed by all N GPUs: scatter / kernel / gather), torch SDPA (cuDNN) beside it.
  cpu_baseline  the reference's own NumPy path (unmodified reference package under baseline/_ref when installed, else
                the oracle port) on the host cores over a bounded sample of (batch, head) slices; rank 0, N=1 only.
  --impl reference   times that CPU path alone, all host cores, same metric/config.
"""
import argparse
import ctypes
import json
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
        return {"bf16": float(d["bf16_tflops"]), "bf16_sustained": float(d.get("bf16_tflops_sustained", 0)) or None,
                "hbm": float(d["hbm_gbs"]), "src": "measured (MEASURED_PEAKS.json)"}
    return {"bf16": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md: 1.59 PFLOP/s, 6.65 TB/s)"}


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
def cpu_reference_run(S, D, steps, warmup, budget_s, max_workers=None):
    """Times `steps` steps; a step = one slice [1,1,S,D] per worker process (baseline/run_ref_cpu.py), all workers at once.
    Returns TFLOP/s over the slices actually processed and the full-step time extrapolated from it."""
    cores = len(os.sched_getaffinity(0))
    workers = max(1, min(cores, max_workers or 64))
    worker = os.path.join(ROOT, "baseline", "run_ref_cpu.py")

    def one_round(n_per_worker, seed0):
        t0 = time.perf_counter()
        procs = [subprocess.Popen([sys.executable, worker, str(S), str(D), str(n_per_worker), str(seed0 + i)],
                                  stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for i in range(workers)]
        outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
        wall = time.perf_counter() - t0
        return wall, outs

    # calibration: one slice on one worker (process start + imports are NOT part of the timed compute below)
    cal = json.loads(subprocess.run([sys.executable, worker, str(S), str(D), "1", "0"], capture_output=True, text=True)
                     .stdout.strip().splitlines()[-1])
    t_slice = cal["seconds"]
    per_step = budget_s / max(1, steps + warmup)
    n_per = max(1, min(4, int(per_step / max(t_slice * 2.0, 1e-3))))
    for w in range(warmup):
        one_round(n_per, 1000 + 100 * w)
    comp = []
    for s in range(steps):
        _, outs = one_round(n_per, 2000 + 100 * s)
        comp.append(max(o["seconds"] for o in outs))          # slowest worker's compute time (imports excluded)
    nslices = workers * n_per
    total = sum(comp)
    flops = causal_flops(1, 1, S, D) * nslices * steps
    return {"tflops": flops / total / 1e12, "sample_ms": 1e3 * total / steps, "cores": workers, "nslices": nslices,
            "t_slice_1core_s": t_slice, "cores_available": cores, "kind": cal["kind"]}


# ------------------------------------------------------------------ helpers
def gpu_local_cpus(torch, dev_index):
    """CPU list of the GPU's NUMA node (sysfs), or None."""
    try:
        p = torch.cuda.get_device_properties(dev_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
        return sorted(cpus), node
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="aule", choices=["aule", "reference"])
    ap.add_argument("--workload", default="C", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
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
        # The reference's own CPU implementation of the path (NumPy, python/aule/__init__.py:247-271): the unmodified
        # reference package under baseline/_ref when present, else its restatement in oracle/; rank 0 only.
        if rank != 0:
            return
        r = cpu_reference_run(S, D, args.steps, args.warmup, budget_s=150.0)
        full_step_ms = r["sample_ms"] * (B * Hq_r) / r["nslices"]
        line = {"impl": "reference", "metric": METRIC, "value": r["tflops"], "unit": "TFLOP/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": full_step_ms, "higher_is_better": True,
                "scaling": "strong" if sharded_heads else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": r["tflops"], "unit": "TFLOP/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": f"each timed step = {r['nslices']} of the workload's {B * Hq_r} (batch,head) slices "
                                           f"[1,1,{S},{D}] fp32 causal ({r['sample_ms']:.0f} ms), one worker process per core over "
                                           f"{r['cores']} of {r['cores_available']} host cores; ms_per_step is that time scaled to "
                                           f"the full {B * Hq_r}-slice step; the NumPy path is single-threaded per slice "
                                           f"({r['t_slice_1core_s']:.2f} s/slice on 1 core)"},
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
        # stdout is ONE JSON line: the image exports NCCL_DEBUG=VERSION, which makes NCCL print its banner on stdout (so does
        # WARN).  Drop the banner-only level and send whatever NCCL still logs (a level the caller chose) to stderr.
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")     # host-side barrier for the single-process spanning section
    BF16 = ffi.DTYPE_BF16
    stream = torch.cuda.current_stream(dev)
    pk = peaks()

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def mk(shape_q, shape_kv, seed):
        g_ = torch.Generator(device=dev).manual_seed(seed)
        return (torch.randn(*shape_q, device=dev, dtype=torch.bfloat16, generator=g_),
                torch.randn(*shape_kv, device=dev, dtype=torch.bfloat16, generator=g_),
                torch.randn(*shape_kv, device=dev, dtype=torch.bfloat16, generator=g_))

    def fwd_call(q_, k_, v_, o_, lse_):
        b_, hq_, s_, d_ = q_.shape
        rc = lib.aule_attention_forward_dptr(q_.data_ptr(), k_.data_ptr(), v_.data_ptr(), o_.data_ptr(),
                                             lse_.data_ptr() if lse_ is not None else 0, b_, hq_, k_.shape[1], s_, k_.shape[2],
                                             d_, BF16, 0.0, 1, -1, local_rank, stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(ffi.last_error())

    def time_events(fn, steps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps

    # ---------------- primary: resident inputs, device-timed
    q, k, v = mk((B, Hq_r, S, D), (B, Hkv_r, S, D), 42 + rank)
    o = torch.empty_like(q)
    lse = torch.empty(B, Hq_r, S, device=dev, dtype=torch.float32)

    def launch():
        fwd_call(q, k, v, o, lse)

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
    launches = lib.aule_launch_count() - n0
    kernel = lib.aule_last_kernel().decode()
    ms_total = e0.elapsed_time(e1)
    # keep the GPU busy a little longer so the clock sampler sees load even for short regions
    t_busy = time.perf_counter()
    while time.perf_counter() - t_busy < 1.0:
        launch()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop(t_start, time.perf_counter())
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = world * flops_rank / (ms_step * 1e-3) / 1e12

    # ---------------- e2e: host buffers through the C ABI, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        cpus, node = gpu_local_cpus(torch, local_rank)
        old_aff = os.sched_getaffinity(0)
        bound = False
        if cpus and set(cpus) & old_aff:
            try:
                os.sched_setaffinity(0, set(cpus) & old_aff)          # first-touch of the pinned pages on the GPU's NUMA node
                bound = True
            except OSError:
                pass
        hq = torch.empty(q.shape, dtype=torch.bfloat16).pin_memory(); hq.copy_(q)
        hk = torch.empty(k.shape, dtype=torch.bfloat16).pin_memory(); hk.copy_(k)
        hv = torch.empty(v.shape, dtype=torch.bfloat16).pin_memory(); hv.copy_(v)
        ho = torch.empty(q.shape, dtype=torch.bfloat16).pin_memory(); ho.zero_()

        def host_step():
            rc = lib.aule_attention_forward_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), None,
                                                 B, Hq_r, Hkv_r, S, S, D, BF16, 0.0, 1, -1, local_rank)
            if rc != 0:
                raise RuntimeError(ffi.last_error())
        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            host_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            host_step()
        dt = max_over_ranks(time.perf_counter() - t0)
        assert torch.equal(ho.to(dev), o), "host-buffer path and device-pointer path disagree"
        # what the host side allows: the same bytes as bare async copies, H2D and D2H overlapped on two streams, all ranks at once
        s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        dq_, dk_, dv_ = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)

        def bare_copy():
            with torch.cuda.stream(s_h2d):
                dq_.copy_(hq, non_blocking=True); dk_.copy_(hk, non_blocking=True); dv_.copy_(hv, non_blocking=True)
            with torch.cuda.stream(s_d2h):
                ho.copy_(o, non_blocking=True)
            s_h2d.synchronize(); s_d2h.synchronize()
        for _ in range(2):
            bare_copy()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            bare_copy()
        dt_copy = max_over_ranks(time.perf_counter() - t0)
        del dq_, dk_, dv_
        os.sched_setaffinity(0, old_aff)
        h2d = 2 * (q.numel() + k.numel() + v.numel())
        d2h = 2 * q.numel()
        e2e = {"value": world * flops_rank * e_steps / dt / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": world * h2d, "d2h_bytes_per_step": world * d2h, "steps": e_steps,
               "ms_per_step": 1e3 * dt / e_steps,
               "api": "aule_attention_forward_host (pinned host buffers, chunked H2D/compute/D2H overlap on three streams)",
               "copy_ceiling": {"ms_per_step": 1e3 * dt_copy / e_steps,
                                "per_gpu_h2d_GBps": h2d * e_steps / dt_copy / 1e9, "per_gpu_d2h_GBps": d2h * e_steps / dt_copy / 1e9,
                                "what": "bare overlapped cuMemcpyAsync of the same bytes (no kernel), all ranks at once",
                                "e2e_frac_of_ceiling": dt_copy / dt},
               "numa": {"gpu_numa_node": node, "bound_to_local_cpus": bound, "local_cpus": len(cpus) if cpus else None}}
        del hq, hk, hv, ho

    # ---------------- secondary measurements (same run, same box)
    secondary = None
    if not args.no_secondary:
        secondary = {}
        sec_steps = max(5, min(args.steps, 10))
        del q, k, v, o, lse
        torch.cuda.empty_cache()

        # config B forward
        Bb, Hb, Hkb, Sb, Db, _ = WORKLOADS["B"]
        qb, kb_, vb = mk((Bb, Hb, Sb, Db), (Bb, Hkb, Sb, Db), 7)
        ob = torch.empty_like(qb); lb = torch.empty(Bb, Hb, Sb, device=dev, dtype=torch.float32)
        ms = time_events(lambda: fwd_call(qb, kb_, vb, ob, lb), 50, 10)
        fl = causal_flops(Bb, Hb, Sb, Db)
        secondary["config_B_fwd"] = {"shape": [Bb, Hb, Hkb, Sb, Db], "ms": ms, "tflops": fl / ms / 1e9, "frac_of_peak": fl / ms / 1e9 / pk["bf16"],
                                     "kernel": lib.aule_last_kernel().decode(), "l2": "inputs 128 MiB ~ L2 size (126 MB): partly L2-resident between launches"}
        try:
            import torch.nn.functional as F
            msy = time_events(lambda: F.scaled_dot_product_attention(qb, kb_, vb, is_causal=True), 30, 5)
            secondary["config_B_fwd"]["torch_sdpa_tflops"] = fl / msy / 1e9
        except Exception as ex:
            secondary["config_B_fwd"]["torch_sdpa_tflops"] = f"unavailable: {type(ex).__name__}"
        del qb, kb_, vb, ob, lb

        def bwd_block(name, Bx, Hx, Hkx, Sx, Dx, steps_):
            qx, kx, vx = mk((Bx, Hx, Sx, Dx), (Bx, Hkx, Sx, Dx), 11)
            ox = torch.empty_like(qx); lx = torch.empty(Bx, Hx, Sx, device=dev, dtype=torch.float32)
            fwd_call(qx, kx, vx, ox, lx)
            dox = torch.randn_like(ox)
            dqx, dkx, dvx = torch.empty_like(qx), torch.empty_like(kx), torch.empty_like(vx)

            def bwd():
                rc = lib.aule_attention_backward_dptr(qx.data_ptr(), kx.data_ptr(), vx.data_ptr(), ox.data_ptr(), dox.data_ptr(),
                                                      lx.data_ptr(), dqx.data_ptr(), dkx.data_ptr(), dvx.data_ptr(), Bx, Hx, Hkx,
                                                      Sx, Sx, Dx, BF16, 0.0, 1, local_rank, stream.cuda_stream)
                if rc != 0:
                    raise RuntimeError(ffi.last_error())
            n_before = lib.aule_launch_count()
            ms_b = time_events(bwd, steps_, 3)
            n_launch = (lib.aule_launch_count() - n_before) // (steps_ + 3)
            ms_f = time_events(lambda: fwd_call(qx, kx, vx, ox, lx), steps_, 3)
            flx = causal_flops(Bx, Hx, Sx, Dx)
            res = {"shape": [Bx, Hx, Hkx, Sx, Dx], "bwd_ms": ms_b, "bwd_tflops": 2.5 * flx / ms_b / 1e9,
                   "bwd_frac_of_peak": 2.5 * flx / ms_b / 1e9 / pk["bf16"], "fwd_ms": ms_f, "fwd_tflops": flx / ms_f / 1e9,
                   "fwd_bwd_tflops": 3.5 * flx / (ms_b + ms_f) / 1e9, "kernels_per_backward": int(n_launch),
                   "flops": "backward = 2.5 x forward (5 GEMMs counted; the dQ kernel recomputes S and dP: 7 issued)"}
            secondary[name] = res
            return qx, kx, vx, ox, lx, dox

        t_ = bwd_block("config_C_half_bwd", 4, 32, 8, 4096, 128, sec_steps)
        del t_
        torch.cuda.empty_cache()
        t_ = bwd_block("config_B_bwd", 4, 32, 32, 2048, 64, sec_steps)
        del t_
        torch.cuda.empty_cache()
        # tensor-core backward outside the (64, 128) x (causal, full) grid: a padded head dim and a sliding window
        t_ = bwd_block("padded_D96_bwd", 4, 32, 8, 4096, 96, sec_steps)
        secondary["padded_D96_bwd"]["note"] = "head_dim 96 runs the D = 128 kernels (TMA zero-fills / clips the padding): FLOPs counted at D = 96"
        del t_
        torch.cuda.empty_cache()

        # fp32 inputs: tf32 tensor-core forward (opt-in) against the exact CUDA-core fp32 kernel and torch SDPA
        try:
            Bt, Ht, St, Dt = 4, 32, 2048, 64
            gt = torch.Generator(device=dev).manual_seed(5)
            qt_, kt_, vt_ = (torch.randn(Bt, Ht, St, Dt, device=dev, generator=gt) for _ in range(3))
            ot_ = torch.empty_like(qt_); lt_ = torch.empty(Bt, Ht, St, device=dev, dtype=torch.float32)

            def f32_call(code):
                rc = lib.aule_attention_forward_dptr(qt_.data_ptr(), kt_.data_ptr(), vt_.data_ptr(), ot_.data_ptr(), lt_.data_ptr(),
                                                     Bt, Ht, Ht, St, St, Dt, code, 0.0, 1, -1, local_rank, stream.cuda_stream)
                if rc != 0:
                    raise RuntimeError(ffi.last_error())
            flt = causal_flops(Bt, Ht, St, Dt)
            ms_t = time_events(lambda: f32_call(3), 30, 5)
            kern_t = lib.aule_last_kernel().decode()
            o_tf32 = ot_.clone()
            ms_x = time_events(lambda: f32_call(0), 3, 1)
            err = ((o_tf32 - ot_).abs().max() / ot_.abs().max()).item()
            secondary["fp32_tf32_fwd"] = {"shape": [Bt, Ht, Ht, St, Dt], "dtype": "fp32 tensors, tf32 tensor-core math (dtype code 3)", "ms": ms_t,
                                          "tflops": flt / ms_t / 1e9, "kernel": kern_t, "exact_fp32_cuda_core_ms": ms_x,
                                          "exact_fp32_cuda_core_tflops": flt / ms_x / 1e9, "max_err_vs_exact_rel_to_scale": err}
            try:
                import torch.nn.functional as F
                torch.backends.cuda.matmul.allow_tf32 = True
                msy = time_events(lambda: F.scaled_dot_product_attention(qt_, kt_, vt_, is_causal=True), 10, 3)
                secondary["fp32_tf32_fwd"]["torch_sdpa_fp32_tflops"] = flt / msy / 1e9
            except Exception as ex:
                secondary["fp32_tf32_fwd"]["torch_sdpa_fp32_tflops"] = f"unavailable: {type(ex).__name__}"
            del qt_, kt_, vt_, ot_, lt_, o_tf32
        except Exception as ex:
            secondary["fp32_tf32_fwd"] = {"error": f"{type(ex).__name__}: {ex}"}
        torch.cuda.empty_cache()
        qe, ke, ve, oe, le, doe = bwd_block("config_E_fwd_bwd", 2, 16, 16, 1024, 64, 50)
        # config E end to end from pinned host buffers: forward (+LSE) then backward, all copies inside
        hs = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (qe, ke, ve, oe, doe, le)]
        for h_, t_ in zip(hs, (qe, ke, ve, oe, doe, le)):
            h_.copy_(t_)
        hq_, hk_, hv_, ho_, hdo_, hl_ = hs
        hdq, hdk, hdv = (torch.empty(t.shape, dtype=torch.bfloat16).pin_memory() for t in (qe, ke, ve))
        fp = ctypes.POINTER(ctypes.c_float)

        def e_host():
            rc = lib.aule_attention_forward_host(hq_.data_ptr(), hk_.data_ptr(), hv_.data_ptr(), ho_.data_ptr(),
                                                 ctypes.cast(hl_.data_ptr(), fp), 2, 16, 16, 1024, 1024, 64, BF16, 0.0, 1, -1, local_rank)
            if rc == 0:
                rc = lib.aule_attention_backward_host(hq_.data_ptr(), hk_.data_ptr(), hv_.data_ptr(), ho_.data_ptr(), hdo_.data_ptr(),
                                                      ctypes.cast(hl_.data_ptr(), fp), hdq.data_ptr(), hdk.data_ptr(), hdv.data_ptr(),
                                                      2, 16, 16, 1024, 1024, 64, BF16, 0.0, 1, local_rank)
            if rc != 0:
                raise RuntimeError(ffi.last_error())
        for _ in range(3):
            e_host()
        t0 = time.perf_counter()
        for _ in range(20):
            e_host()
        dt_e = (time.perf_counter() - t0) / 20
        fle = causal_flops(2, 16, 1024, 64)
        nb = lambda t: t.numel() * t.element_size()   # noqa: E731
        secondary["config_E_fwd_bwd"]["e2e_host"] = {
            "ms_per_step": 1e3 * dt_e, "tflops": 3.5 * fle / dt_e / 1e12,
            "h2d_bytes_per_step": sum(nb(t) for t in (qe, ke, ve)) + sum(nb(t) for t in (qe, ke, ve, oe, doe, le)),
            "d2h_bytes_per_step": nb(oe) + nb(le) + sum(nb(t) for t in (qe, ke, ve)),
            "api": "aule_attention_forward_host (+LSE) then aule_attention_backward_host, pinned host buffers"}
        del qe, ke, ve, oe, le, doe, hs, hdq, hdk, hdv

        # paged-KV decode (Llama-3-8B decode shape, 1 GiB of shuffled 16-token pages): HBM-bound
        Bp, Hqp, Hkp, Dp, bs, ctx = 32, 32, 8, 128, 16, 8192
        g_ = torch.Generator(device=dev).manual_seed(3)
        mb = ctx // bs
        nbk = Bp * mb
        pq = torch.randn(Bp, Hqp, Dp, device=dev, dtype=torch.bfloat16, generator=g_)
        kc = torch.randn(nbk, bs, Hkp, Dp, device=dev, dtype=torch.bfloat16, generator=g_)
        vc = torch.randn(nbk, bs, Hkp, Dp, device=dev, dtype=torch.bfloat16, generator=g_)
        bt = torch.randperm(nbk, device=dev, generator=g_).reshape(Bp, mb).to(torch.int32)
        cl = torch.full((Bp,), ctx, dtype=torch.int32, device=dev)
        po = torch.empty_like(pq)

        def paged():
            rc = lib.aule_attention_paged_decode_dptr(pq.data_ptr(), kc.data_ptr(), vc.data_ptr(), bt.data_ptr(), cl.data_ptr(),
                                                      po.data_ptr(), Bp, Hqp, Hkp, Dp, nbk, bs, mb, ctx, BF16, 0.0, -1,
                                                      local_rank, stream.cuda_stream)
            if rc != 0:
                raise RuntimeError(ffi.last_error())
        ms_p = time_events(paged, 50, 10)
        bytes_alg = Bp * ctx * Hkp * Dp * 2 * 2 + 2 * Bp * Hqp * Dp * 2
        secondary["paged_decode"] = {"shape": {"B": Bp, "Hq": Hqp, "Hkv": Hkp, "D": Dp, "block_size": bs, "context": ctx},
                                     "kv_cache_MiB": 2 * kc.numel() * 2 / 2**20, "ms": ms_p, "GBps": bytes_alg / ms_p / 1e6,
                                     "frac_of_hbm_peak": bytes_alg / ms_p / 1e6 / pk["hbm"], "hbm_peak_GBps": pk["hbm"],
                                     "algorithmic_bytes": bytes_alg, "bound": "hbm", "kernel": lib.aule_last_kernel().decode()}
        del pq, kc, vc, bt, cl, po
        torch.cuda.empty_cache()

        # config D: [1,32,32768,128] causal, heads sharded over the ranks (strong scaling), resident inputs
        Bd, Hd, Hkd, Sd, Dd, _ = WORKLOADS["D"]
        if Hd % world == 0:
            hs_ = Hd // world
            qd, kd, vd = mk((Bd, hs_, Sd, Dd), (Bd, hs_, Sd, Dd), 100 + rank)
            od = torch.empty_like(qd); ld = torch.empty(Bd, hs_, Sd, device=dev, dtype=torch.float32)
            for _ in range(2):
                fwd_call(qd, kd, vd, od, ld)
            sync_all()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nD = 3 if world == 1 else 5
            ea.record(stream)
            for _ in range(nD):
                fwd_call(qd, kd, vd, od, ld)
            eb.record(stream)
            sync_all()
            msd = max_over_ranks(ea.elapsed_time(eb) / nD)
            fld = causal_flops(Bd, Hd, Sd, Dd)
            secondary["config_D_head_sharded"] = {"shape": [Bd, Hd, Hkd, Sd, Dd], "heads_per_gpu": hs_, "n_gpus": world, "ms": msd,
                                                  "tflops_total": fld / msd / 1e9, "frac_of_peak_per_gpu": fld / msd / 1e9 / world / pk["bf16"],
                                                  "scaling": "strong (fixed total work; efficiency = tflops_total(N) / (N * tflops_total(1)))"}
            del qd, kd, vd, od, ld
            torch.cuda.empty_cache()

        # spanning call (native, one process): config D's tensors live on GPU 0, all `world` GPUs compute
        if world > 1:
            # The other ranks must not have an NCCL kernel spinning on their GPU while rank 0's process drives all GPUs:
            # they wait on a host-side (gloo) barrier.
            sync_all()
            dist.barrier(group=cpu_group)
            if rank == 0:
                try:
                    qs, ks, vs = mk((Bd, Hd, Sd, Dd), (Bd, Hkd, Sd, Dd), 5)
                    os_ = torch.empty_like(qs); ls = torch.empty(Bd, Hd, Sd, device=dev, dtype=torch.float32)
                    devs = (ctypes.c_int32 * world)(*range(world))
                    tm = (ctypes.c_float * 4)()

                    def span(chunks):
                        rc = lib.aule_attention_forward_spanning_dptr(qs.data_ptr(), ks.data_ptr(), vs.data_ptr(), os_.data_ptr(), ls.data_ptr(),
                                                                      Bd, Hd, Hkd, Sd, Sd, Dd, BF16, 0.0, 1, -1, local_rank, stream.cuda_stream,
                                                                      devs, world, chunks, tm)
                        if rc != 0:
                            raise RuntimeError(ffi.last_error())
                        return [float(x) for x in tm]
                    span(1); span(1)
                    serial = [span(1) for _ in range(3)]
                    span(4)
                    over = [span(4) for _ in range(3)]
                    med = lambda rows, i: statistics.median(r[i] for r in rows)   # noqa: E731
                    fld = causal_flops(Bd, Hd, Sd, Dd)
                    local = torch.empty_like(qs)
                    fwd_call(qs, ks, vs, local, None)
                    torch.cuda.synchronize(dev)
                    secondary["spanning_call"] = {
                        "what": f"one call, tensors of config D resident on GPU 0, computed by {world} GPUs "
                                "(aule_attention_forward_spanning_dptr: cuMemcpyPeerAsync scatter of Q/K/V head slabs over NVLink, kernels, gather of O/LSE)",
                        "serial_phases_ms": {"scatter": med(serial, 0), "kernel": med(serial, 1), "gather": med(serial, 2), "total": med(serial, 3)},
                        "overlapped_4_chunks_ms": {"scatter": med(over, 0), "kernel": med(over, 1), "gather_tail": med(over, 2), "total": med(over, 3)},
                        "tflops_total_overlapped": fld / med(over, 3) / 1e9,
                        "bytes_scattered_from_gpu0": int(3 * qs.numel() * 2 * (world - 1) / world),
                        "bytes_gathered_to_gpu0": int((qs.numel() * 2 + ls.numel() * 4) * (world - 1) / world),
                        "bit_identical_to_single_gpu": bool(torch.equal(local, os_))}
                    del qs, ks, vs, os_, ls, local
                except Exception as ex:               # never lose the headline line to a secondary measurement
                    secondary["spanning_call"] = {"error": f"{type(ex).__name__}: {ex}"}
            dist.barrier(group=cpu_group)
            sync_all()

        # yardstick beside the headline, same run: torch SDPA (cuDNN) on the primary workload (N=1 only: keeps N>1 runs short)
        if world == 1:
            try:
                import torch.nn.functional as F
                qy, ky, vy = mk((B, Hq_r, S, D), (B, Hkv_r, S, D), 42)
                msy = time_events(lambda: F.scaled_dot_product_attention(qy, ky, vy, is_causal=True, enable_gqa=(Hq_r != Hkv_r)), 10, 3)
                secondary["torch_sdpa_primary_workload"] = {"ms": msy, "tflops": flops_rank / msy / 1e9,
                                                            "what": "torch.nn.functional.scaled_dot_product_attention (cuDNN / flash backend chosen by torch) on the same shape"}
                del qy, ky, vy
            except Exception as ex:
                secondary["torch_sdpa_primary_workload"] = f"unavailable: {type(ex).__name__}"

    if rank != 0:
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return

    per_launch_tflops = flops_rank / (ms_step * 1e-3) / 1e12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "fwd_traffic_bytes.json")
    if os.path.exists(tpath) and args.workload == "C":
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    line = {"metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if sharded_heads else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "kernel": kernel,
            "roofline": {"bound": "tensor", "achieved": per_launch_tflops, "peak": pk["bf16"], "unit": "TFLOP/s",
                         "frac": per_launch_tflops / pk["bf16"],
                         "frac_of_sustained_peak": (per_launch_tflops / pk["bf16_sustained"]) if pk["bf16_sustained"] else None,
                         "peak_sustained": pk["bf16_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": pk["src"] + " bf16_tflops (burst)",
                         "algorithmic_flops_per_launch": flops_rank,
                         "compulsory_hbm_bytes_per_launch": 2 * (2 * B * Hq_r * S * D + 2 * B * Hkv_r * S * D)},
            "pct_of_nominal_2250_tflops": 100.0 * per_launch_tflops / 2250.0,
            "secondary": secondary}
    if args.gpus == 1 and not args.no_cpu:
        r = cpu_reference_run(S, D, steps=1, warmup=0, budget_s=20.0, max_workers=1)
        line["cpu_baseline"] = {"value": r["tflops"], "unit": "TFLOP/s", "cores": 1, "kind": r["kind"],
                                "sample": f"{r['nslices']} (batch,head) slice(s) [1,1,{S},{D}] fp32 causal of the same workload "
                                          f"through the reference NumPy path (single-threaded einsum); full step = {B * Hq_r} slices"}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
