"""Pipeline timeline of the forward kernel (bring-up tool, GPU box): CTA 0 of the traced variant
(aule_fwd_sm100_bf16_d*_e0, tuning builds: make EXTRA_NVFLAGS=-DAULE_TUNING_VARIANTS) records clock64 at its
pipeline events through aule_set_trace_buffer.  Prints (a) per-event-pair statistics (how long each wait took) and
(b) the raw timeline of a window of events.
usage: python tools/fwd_trace.py [C|B|D8] [first_event] [n_events]"""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()
cfg = sys.argv[1] if len(sys.argv) > 1 else "C"
first = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
count = int(sys.argv[3]) if len(sys.argv) > 3 else 160
B, Hq, Hkv, S, D = {"C": (8, 32, 8, 4096, 128), "B": (4, 32, 32, 2048, 64), "D8": (1, 4, 4, 32768, 128)}[cfg]
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
lib.aule_set_kernel_path(16)
for _ in range(3):
    cuda_flash.forward_with_lse(q, k, v, causal=True)
torch.cuda.synchronize()
NREG = 4
buf = torch.zeros(NREG * 4096, dtype=torch.int64, device="cuda")
lib.aule_set_trace_buffer(buf.data_ptr())
cuda_flash.forward_with_lse(q, k, v, causal=True)
torch.cuda.synchronize()
print("kernel:", lib.aule_last_kernel().decode())
lib.aule_set_trace_buffer(0)
lib.aule_set_kernel_path(0)
ev = buf.cpu().numpy().astype("uint64")
names = {1: "iss: wait P0", 2: "iss: P0 ok -> PV0[0:6]", 3: "iss: P0b ok -> PV0[6:8]",
         17: "iss: wait P1", 18: "iss: P1 ok -> PV1[0:6]", 19: "iss: P1b ok -> PV1[6:8]",
         4: "iss: QK wait sfree", 5: "iss: sfree ok -> QK", 6: "iss: wait KV stage", 7: "iss: KV ok",
         8: "tma: wait stage empty", 9: "tma: stage empty ok -> load",
         10: "smx: wait S", 11: "smx: S ok", 12: "smx: S in regs (sfree)", 13: "smx: max/rescale done",
         14: "smx: half exps done, wait pvdone", 15: "smx: pvdone ok", 16: "smx: P[0:96] published", 17 + 256: "",
         }
smx_names = {10: "wait S", 11: "S ok", 12: "S in regs", 13: "max done", 14: "exp half, wait pvdone", 15: "pvdone ok",
             16: "P[0:96] pub", 17: "P[96:128] pub"}
rows = []
per_region = defaultdict(list)
for region in range(NREG):
    for x in ev[region * 4096:(region + 1) * 4096]:
        if x == 0:
            break
        tag, t = int(x >> 48), int(x & ((1 << 48) - 1))
        rows.append((t, region, tag >> 8, tag & 255))
        per_region[region].append((t, tag >> 8, tag & 255))
rows.sort()
t0 = rows[0][0]
print(f"{len(rows)} events; span {rows[-1][0] - t0} cycles; per region: {[len(per_region[r]) for r in range(NREG)]}")


def label(region, code):
    if region in (1, 2):
        return f"smx{region - 1}: " + smx_names.get(code, str(code))
    return names.get(code, str(code))


# (a) statistics of consecutive-event gaps per region, keyed by (code_a -> code_b)
for region in range(NREG):
    evs = per_region[region]
    gaps = defaultdict(list)
    for (ta, ca, _), (tb, cb, _) in zip(evs, evs[1:]):
        gaps[(ca, cb)].append(tb - ta)
    print(f"--- region {region}: gap statistics (cycles): count / median / p90 / sum")
    tot = sum(sum(v_) for v_ in gaps.values())
    for (ca, cb), v_ in sorted(gaps.items(), key=lambda kv: -sum(kv[1])):
        v_ = sorted(v_)
        print(f"  {label(region, ca):38s} -> {label(region, cb):38s} n={len(v_):5d} med={v_[len(v_) // 2]:6d} p90={v_[int(len(v_) * 0.9)]:6d} "
              f"sum={sum(v_):9d} ({100.0 * sum(v_) / tot:4.1f}%)")
# (b) raw window
print(f"--- timeline, events [{first}, {first + count})")
for t, region, code, step in rows[first:first + count]:
    print(f"{t - t0:10d}  r{region} {step:3d}  {label(region, code)}")
