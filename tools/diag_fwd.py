"""Bring-up diagnostics for the tcgen05 forward kernel (run on the GPU box).
Separates the two GEMM paths: LSE depends only on Q K^T (+mask); with Q = 0 the output is a
prefix mean of V (exercises P V and the V layout only). Prints error maps per 32-row / 16-col block."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
from oracle import attention_oracle as orc  # noqa: E402  (diagnostic tool, not product)
from aule import cuda_flash, ffi  # noqa: E402


def blockmap(err, rb=32, cb=16):
    R, C = err.shape
    return np.array([[err[r:r + rb, c:c + cb].max() for c in range(0, C, cb)] for r in range(0, R, rb)])


def run(tag, q, k, v, causal):
    out, lse = cuda_flash.forward_with_lse(q, k, v, causal=causal)
    torch.cuda.synchronize()
    kern = ffi.load_library().aule_last_kernel().decode()
    fq, fk, fv = (t.float().cpu().numpy() for t in (q, k, v))
    exp, el = orc.attention_ref(fq, fk, fv, causal=causal)
    o = out.float().cpu().numpy()
    e_o = orc.rel_err_to_scale(o, exp)
    e_l = float(np.abs(lse.cpu().numpy() - el).max())
    print(f"[{tag}] kernel={kern} out_rel={e_o:.3e} lse_abs={e_l:.3e} finite={np.isfinite(o).all()}")
    if e_o > 1e-2 or e_l > 1e-2:
        np.set_printoptions(precision=2, linewidth=250, suppress=False)
        print("  out err map (32-row x 16-col blocks), head 0:")
        print(blockmap(np.abs(o - exp)[0, 0]))
        print("  lse err per 32 rows, head 0:", np.abs(lse.cpu().numpy() - el)[0, 0].reshape(-1, 32).max(axis=1))
        print("  out[0,0,:4,:8]\n", o[0, 0, :4, :8], "\n  exp\n", exp[0, 0, :4, :8])
    return e_o, e_l


def main():
    torch.manual_seed(0)
    dev = "cuda"
    for D in (128, 64):
        for dt in (torch.bfloat16,):
            S = 256
            q = torch.randn(1, 1, S, D, device=dev).to(dt)
            k = torch.randn(1, 1, S, D, device=dev).to(dt)
            v = torch.randn(1, 1, S, D, device=dev).to(dt)
            print(f"===== D={D} S={S}")
            run("qk-only (lse) + full", q, k, v, False)
            run("Q=0 -> mean(V) (PV path)", torch.zeros_like(q), k, v, False)
            vi = torch.zeros_like(v)
            vi[0, 0, torch.arange(min(S, D)), torch.arange(min(S, D))] = 1.0
            run("V=I slab (P readout)", q, k, vi, False)
            run("causal", q, k, v, True)
            S2 = 1024
            q2, k2, v2 = (torch.randn(2, 4, S2, D, device=dev).to(dt) for _ in range(3))
            run("multi-tile causal S=1024", q2, k2, v2, True)


if __name__ == "__main__":
    main()
