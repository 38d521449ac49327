"""Golden vectors from the UNMODIFIED reference *Triton* kernels, executed on the CPU.

Run ONCE in the build container (needs /root/reference and the `triton` package; the GPU box has
no /root/reference):   python tests/golden/gen_golden_triton.py

The reference's GPU path (python/aule/triton_flash.py: _flash_attn_fwd_kernel :62-235,
_flash_attn_bwd_kernel :242-350, _compute_delta_kernel :353-379, FlashAttentionTritonFunc :386-526) and its
paged decode kernel (python/aule/triton_flash_amd.py:544-740) are plain Triton; with TRITON_INTERPRET=1 Triton runs
the very same kernel source on CPU tensors through its NumPy interpreter -- no GPU, no edits to the reference.
The two files are loaded by path (not through `import aule`, whose __init__ would disable Triton without CUDA,
triton_flash.py:617-618).  This pins the parts of the oracle the NumPy reference path cannot express: GQA/MQA,
explicit scale, LSE, cross attention, sliding window, the backward (dQ/dK/dV incl. the GQA group sum) and the paged
decode.  Inputs AND outputs are stored (fp32, .npz) so the fixtures do not depend on torch's RNG stream.
"""
import importlib.util
import json
import os
import sys

os.environ["TRITON_INTERPRET"] = "1"

import numpy as np  # noqa: E402
import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


ref = _load("ref_triton_flash", "/root/reference/python/aule/triton_flash.py")
ref_amd = _load("ref_triton_flash_amd", "/root/reference/python/aule/triton_flash_amd.py")

# name, B, Hq, Hkv, Sq, Sk, D, causal, scale, window, with_backward
FWD_CASES = [
    ("mha_causal_1x8x64x64", 1, 8, 8, 64, 64, 64, True, None, -1, False),          # python/tests/test_triton.py:35-46
    ("mha_noncausal_1x8x64x64", 1, 8, 8, 64, 64, 64, False, None, -1, False),      # :48-58
    ("mha_causal_1x8x64x128", 1, 8, 8, 64, 64, 128, True, None, -1, False),        # :128-139
    ("gqa_12to2_1x12x64x64", 1, 12, 2, 64, 64, 64, True, None, -1, False),         # :96-110
    ("mqa_8to1_1x8x64x64", 1, 8, 1, 64, 64, 64, True, None, -1, False),            # :112-126
    ("cross_1x4_q16_k32_d64", 1, 4, 4, 16, 32, 64, False, None, -1, False),        # tests/test_cross_attn.py:12-60
    ("ragged_gqa_2x4x100x64", 2, 4, 2, 100, 100, 64, True, None, -1, False),       # bounds masks, triton_flash.py:198,:227
    ("scale_0p2_1x4x48x64", 1, 4, 4, 48, 48, 64, True, 0.2, -1, False),            # explicit scale, :394-395
    ("window_24_1x4x96x64", 1, 4, 4, 96, 96, 64, True, None, 24, False),           # sliding window, :190-193
    ("bwd_mha_1x4x96x64", 1, 4, 4, 96, 96, 64, True, None, -1, True),              # :78-94 (backward vs autograd)
    ("bwd_gqa_1x4x80x64", 1, 4, 2, 80, 80, 64, True, None, -1, True),              # GQA group sum, triton_flash.py:345-347
    ("bwd_noncausal_1x2x64x128", 1, 2, 2, 64, 64, 128, False, None, -1, True),
]
# name, B, Hq, Hkv, D, block_size, num_blocks, max_blocks, context_lens, window
PAGED_CASES = [
    ("paged_mha_bs16", 2, 4, 4, 64, 16, 12, 4, [37, 64], -1),
    ("paged_gqa_bs16_d128", 3, 8, 2, 128, 16, 24, 6, [96, 1, 50], -1),
    ("paged_mqa_bs32", 2, 8, 1, 64, 32, 10, 4, [128, 33], -1),
    ("paged_window_20", 2, 4, 2, 64, 16, 12, 4, [37, 64], 20),
]


def main():
    arrays, meta = {}, {"triton": __import__("triton").__version__, "torch": torch.__version__, "fwd": {}, "paged": {}}
    for (name, B, Hq, Hkv, Sq, Sk, D, causal, scale, window, bwd) in FWD_CASES:
        torch.manual_seed(42)                                                       # python/tests/test_triton.py:37
        q = torch.randn(B, Hq, Sq, D, requires_grad=bwd)
        k = torch.randn(B, Hkv, Sk, D, requires_grad=bwd)
        v = torch.randn(B, Hkv, Sk, D, requires_grad=bwd)
        out = ref.FlashAttentionTritonFunc.apply(q, k, v, causal, scale, window)             # forward(ctx,q,k,v,causal,scale,window_size,...) :388
        arrays[name + ".q"], arrays[name + ".k"], arrays[name + ".v"] = (t.detach().numpy() for t in (q, k, v))
        arrays[name + ".out"] = out.detach().numpy()
        if bwd:
            lse = out.grad_fn.saved_tensors[4]                                      # ctx.save_for_backward(q,k,v,out,L) :466
            arrays[name + ".lse"] = lse.detach().numpy()
            do = torch.randn_like(out)                                              # test_triton.py:88
            out.backward(do)
            arrays[name + ".do"] = do.numpy()
            arrays[name + ".dq"], arrays[name + ".dk"], arrays[name + ".dv"] = (t.grad.numpy() for t in (q, k, v))
        meta["fwd"][name] = {"shape": [B, Hq, Hkv, Sq, Sk, D], "causal": causal, "scale": scale, "window": window,
                             "backward": bwd, "out_sum": float(out.detach().double().sum())}
        print(name, "ok", flush=True)
    for (name, B, Hq, Hkv, D, bs, nb, mb, lens, window) in PAGED_CASES:
        torch.manual_seed(7)
        q = torch.randn(B, Hq, D)
        kc, vc = torch.randn(nb, bs, Hkv, D), torch.randn(nb, bs, Hkv, D)
        bt = torch.randperm(nb)[:B * mb].reshape(B, mb).to(torch.int32)
        cl = torch.tensor(lens, dtype=torch.int32)
        out = ref_amd.flash_attention_paged_amd(q, kc, vc, bt, cl, window_size=window)
        for key, t in (("q", q), ("k_cache", kc), ("v_cache", vc), ("block_tables", bt), ("context_lens", cl), ("out", out)):
            arrays[f"{name}.{key}"] = t.numpy()
        meta["paged"][name] = {"shape": [B, Hq, Hkv, D], "block_size": bs, "num_blocks": nb, "max_blocks": mb,
                               "context_lens": lens, "window": window, "out_sum": float(out.double().sum())}
        print(name, "ok", flush=True)
    np.savez_compressed(os.path.join(HERE, "reference_triton_path.npz"), **arrays)
    with open(os.path.join(HERE, "reference_triton_path.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(arrays), "arrays")


if __name__ == "__main__":
    sys.exit(main())
