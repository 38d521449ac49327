"""Round-2 golden vectors from the UNMODIFIED reference Triton kernels run on the CPU (TRITON_INTERPRET=1), same
method as gen_golden_triton.py (which see).  Run ONCE in the build container:  python tests/golden/gen_golden_r2.py
Adds what round 1 did not pin:
  * (NOT pinned here: the bidirectional sliding window.  The reference's Triton kernel visits only the K/V blocks within
    window//2 of the TILE CENTRE (triton_flash.py:140-144) and then masks |i-j| <= W per row (:190-194), so rows near a
    tile edge lose keys their own window allows -- the result depends on the tile size.  This ABI follows the Vulkan
    shader's per-row rule |i-j| <= window//2 (attention_f32.comp:180-183), pinned by the oracle.)
  * head dims that are not a power of two (the reference pads to BLOCK_K = next_power_of_2(D), triton_flash.py:446),
  * RoPE + attention: the reference's torch formulation apply_rope_separate (:680-703) followed by its attention kernel,
    and -- when it runs under the interpreter -- its fused flash_attention_rope (:561-603) on the same inputs,
  * fp32 [.,.,.,64] forward (dtype policy :405-411).
Inputs AND outputs are stored (fp32, .npz)."""
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

# name, B, Hq, Hkv, Sq, Sk, D, causal, scale, window
FWD_CASES = [
    ("d40_causal_1x4x96x40", 1, 4, 4, 96, 96, 40, True, None, -1),
    ("d80_gqa_1x4x80x80", 1, 4, 2, 80, 80, 80, True, None, -1),
    ("d96_noncausal_1x2x64x96", 1, 2, 2, 64, 64, 96, False, None, -1),
    ("d32_causal_1x4x64x32", 1, 4, 4, 64, 64, 32, True, None, -1),
]
ROPE_CASES = [
    ("rope_gqa_1x4x64x64", 1, 4, 2, 64, 64, 64, True),
    ("rope_mha_1x2x48x128", 1, 2, 2, 48, 48, 128, True),
]


def main():
    arrays, meta = {}, {"triton": __import__("triton").__version__, "torch": torch.__version__, "fwd": {}, "rope": {}}
    for (name, B, Hq, Hkv, Sq, Sk, D, causal, scale, window) in FWD_CASES:
        torch.manual_seed(42)
        q, k, v = torch.randn(B, Hq, Sq, D), torch.randn(B, Hkv, Sk, D), torch.randn(B, Hkv, Sk, D)
        out = ref.FlashAttentionTritonFunc.apply(q, k, v, causal, scale, window)
        for key, t in (("q", q), ("k", k), ("v", v), ("out", out)):
            arrays[f"{name}.{key}"] = t.detach().numpy()
        meta["fwd"][name] = {"shape": [B, Hq, Hkv, Sq, Sk, D], "causal": causal, "scale": scale, "window_triton": window,
                             "out_sum": float(out.double().sum())}
        print(name, "ok", flush=True)
    for (name, B, Hq, Hkv, Sq, Sk, D, causal) in ROPE_CASES:
        torch.manual_seed(42)
        q, k, v = torch.randn(B, Hq, Sq, D), torch.randn(B, Hkv, Sk, D), torch.randn(B, Hkv, Sk, D)
        cos, sin = ref.precompute_rope_frequencies(max(Sq, Sk), D, device="cpu")
        q_rot, k_rot = ref.apply_rope_separate(q, k, cos, sin) if Sq == Sk else (None, None)
        out = ref.FlashAttentionTritonFunc.apply(q_rot, k_rot, v, causal, None, -1)
        fused = None
        try:
            fused = ref.flash_attention_rope(q, k, v, cos, sin, causal=causal)
            fused_note = "ran"
        except Exception as ex:                       # the fused branch joins the rotated halves with tl.join (:131,:180)
            fused_note = f"failed under the interpreter: {type(ex).__name__}: {str(ex)[:120]}"
        for key, t in (("q", q), ("k", k), ("v", v), ("cos", cos), ("sin", sin), ("q_rot", q_rot), ("k_rot", k_rot), ("out", out)):
            arrays[f"{name}.{key}"] = t.detach().numpy()
        if fused is not None:
            arrays[f"{name}.out_fused"] = fused.detach().numpy()
        meta["rope"][name] = {"shape": [B, Hq, Hkv, Sq, Sk, D], "causal": causal, "fused_kernel": fused_note,
                              "fused_vs_separate_max_abs": None if fused is None else float((fused - out).abs().max()),
                              "out_sum": float(out.double().sum())}
        print(name, "ok", fused_note, meta["rope"][name]["fused_vs_separate_max_abs"], flush=True)
    np.savez_compressed(os.path.join(HERE, "reference_triton_path_r2.npz"), **arrays)
    with open(os.path.join(HERE, "reference_triton_path_r2.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(arrays), "arrays")


if __name__ == "__main__":
    sys.exit(main())
