"""Per-step pipeline timeline of the backward kernels (bring-up tool): CTA 0 records clock64 at its pipeline events
through aule_set_trace_buffer; this prints, per step, the event times relative to the step's first event.
The tracer is compiled in only with  make -C aule-attention_b200 EXTRA_NVFLAGS=-DAULE_BWD_TRACE=1  (rebuild without it
afterwards: the checks cost instructions in issue-bound kernels).
usage: python tools/bwd_trace.py [dq|dkv|fused] [first_step] [last_step]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from aule import cuda_flash, ffi  # noqa: E402

lib = ffi.ensure_init()
which = sys.argv[1] if len(sys.argv) > 1 else "dq"
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 8
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 14
B, Hq, Hkv, S, D = (int(x) for x in os.environ.get("AULE_TRACE_SHAPE", "4,32,8,4096,128").split(","))   # config B: 4,32,32,2048,64
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


serial = len(sys.argv) > 4 and sys.argv[4] == "serial"
lib.aule_set_kernel_path(((1 << 26) | int(os.environ.get("AULE_EXTRA_PATH", "0"))) if which == "fused2" else (1 << 17) if which == "fused" else (((1 if which == "dkv" else 2) << 10) | ((1 << 12) if serial else 0)))
for _ in range(3):
    call()
torch.cuda.synchronize()
buf = torch.zeros(3 * 4096, dtype=torch.int64, device="cuda")
lib.aule_set_trace_buffer(buf.data_ptr())
call()
torch.cuda.synchronize()
lib.aule_set_trace_buffer(0)
lib.aule_set_kernel_path(0)
ev = buf.cpu().numpy().astype("uint64")
rows = []
for region in range(3):
    for x in ev[region * 4096:(region + 1) * 4096]:
        if x == 0:
            break
        tag, t = int(x >> 48), int(x & ((1 << 48) - 1))
        rows.append((t, region, tag >> 8, tag & 255))
rows.sort()
t0 = rows[0][0]
print(f"{len(rows)} events; total span {rows[-1][0] - t0} cycles")
names = {16: "iss: dK(i-1) done -> load Q(i+1)", 17: "iss: wait P", 18: "iss: P ok -> dV", 19: "iss: dV done -> load dO(i+1)",
         28: "cmp: P math done (wait dV(i-1))", 10: "iss: wait sfree", 11: "iss: sfree ok", 12: "iss: dO ok -> dP", 13: "iss: dS(j-1) ok", 14: "iss: Q|K ok -> S(j+1)", 15: "iss: dQ(j-1) issued",
         20: "cmp: wait S", 21: "cmp: S ok", 22: "cmp: S in regs", 23: "cmp: P done, wait dP", 24: "cmp: dP ok", 25: "cmp: dS math done",
         26: "cmp: dS cols free", 27: "cmp: dS stored", 30: "iss: dQ(j-1) DONE", 31: "iss: dP issued", 32: "iss: dP DONE",
         33: "iss: S(j+1) issued", 34: "iss: S(j+1) DONE"}
if which == "fused":
    names = {10: "iss: wait dS", 11: "iss: dS ok -> dQ^T, dK", 12: "iss: dO(s+1) ok", 13: "iss: P half0 ok -> dV even", 14: "iss: dQ drained -> dP(s+1)",
             15: "iss: P half1 ok -> dV odd", 16: "iss: Q(s+2) ok -> S(s+2)",
             20: "cmp: wait dP", 21: "cmp: dP ok", 22: "cmp: dS math done", 23: "cmp: dS stored+fenced", 24: "cmp: S(s+1) ok",
             25: "cmp: P half0 published", 26: "cmp: dQ^T ok", 27: "cmp: drained"}
if which == "fused2":
    names = {10: "iss: wait dS", 11: "iss: dS ok -> dQ^T, dK", 12: "iss: dO(h+1) ok -> dP(h+1)", 13: "iss: P(h+1) ok -> dV", 16: "iss: Q(h+2) ok -> S(h+2)",
             20: "cmp: wait dP", 21: "cmp: dP ok", 22: "cmp: dS math done, wait tile free", 23: "cmp: dS stored+fenced", 25: "cmp: P(h+1) published",
             30: "drn: wait dQ^T", 26: "drn: dQ^T ok", 28: "cmp: dQ^T in regs, wait staging free", 29: "cmp: staging free", 27: "cmp: staged"}
if which == "dq":
    names.update({10: "iss: step top (dQ(j-1) issued)", 16: "iss: sfree ok", 11: "iss: K(j+1) ok -> S(j+1)", 18: "iss: S(j+1) issued", 17: "iss: dpfree ok",
                  12: "iss: V(j+1) ok -> dP(j+1)", 19: "iss: dP(j+1) issued", 13: "iss: dS(j) ok -> dQ(j)"})
for t, region, code, step in rows:
    if lo <= step <= hi:
        print(f"{t - t0:9d}  r{region} step {step:3d}  {code:3d} {names.get(code, '')}")
