"""Generate golden vectors from the UNMODIFIED reference Python path.

Run ONCE in the build container (needs /root/reference; the GPU box does not
have it):   python tests/golden/gen_golden.py

Imports /root/reference/python/aule and calls the reference's own public entry
point `aule.flash_attention` with NumPy inputs, which (no libaule.so, no CUDA)
routes to `_cpu_attention` (python/aule/__init__.py:188-193,:240-271).

Inputs are NOT stored: they are regenerated from the reference's own fixture
recipe (python/tests/conftest.py:8-16: np.random.seed(42); q,k,v = successive
randn(B,H,S,D).astype(float32)).  Outputs are stored as float32 .npz.
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/python")
import aule  # noqa: E402  (the reference package)

HERE = os.path.dirname(os.path.abspath(__file__))

# (name, B, H, S, D, causal) -- shapes of python/tests/test_cpu.py:20-64,
# python/tests/test_triton.py:35-64 and BASELINE.json configs[0] (config A).
CASES = [
    ("cpu_causal_1x4x32x64", 1, 4, 32, 64, True),        # test_cpu.py:20-29
    ("cpu_noncausal_1x4x32x64", 1, 4, 32, 64, False),    # test_cpu.py:31-39
    ("cpu_batch_4x8x64x64", 4, 8, 64, 64, True),         # test_cpu.py:41-49
    ("cpu_d32_1x4x32x32", 1, 4, 32, 32, True),           # test_cpu.py:51-64
    ("cpu_d128_1x4x32x128", 1, 4, 32, 128, True),        # test_cpu.py:51-64
    ("triton_shape_1x8x64x64_causal", 1, 8, 64, 64, True),     # test_triton.py:35-46
    ("triton_shape_1x8x64x64_noncausal", 1, 8, 64, 64, False), # test_triton.py:48-58
    ("triton_shape_1x8x64x128_causal", 1, 8, 64, 128, True),   # test_triton.py:128-139
    ("configA_1x8x256x64_causal", 1, 8, 256, 64, True),  # BASELINE.json configs[0]
]


def make_inputs(B, H, S, D):
    np.random.seed(42)                                   # conftest.py:10
    q = np.random.randn(B, H, S, D).astype(np.float32)   # conftest.py:12-14
    k = np.random.randn(B, H, S, D).astype(np.float32)
    v = np.random.randn(B, H, S, D).astype(np.float32)
    return q, k, v


def main():
    assert aule.get_available_backends() == ["cpu"], aule.get_available_backends()
    outs = {}
    meta = {"reference_version": aule.__version__, "numpy": np.__version__, "cases": {}}
    for name, B, H, S, D, causal in CASES:
        q, k, v = make_inputs(B, H, S, D)
        o = aule.flash_attention(q, k, v, causal=causal)
        assert o.dtype == np.float32 and o.shape == q.shape
        outs[name] = o
        meta["cases"][name] = {
            "shape": [B, H, S, D], "causal": causal,
            "sum": float(o.sum(dtype=np.float64)), "max_abs": float(np.abs(o).max()),
            "sha256_16": hashlib.sha256(o.tobytes()).hexdigest()[:16],
        }
    # GQA semantic pin (tests/test_gqa_unit.py:20-55): the reference NumPy path has
    # no GQA, so the golden is the reference path applied to repeat_interleave'd K/V.
    np.random.seed(42)
    q = np.random.randn(1, 4, 16, 64).astype(np.float32)
    k = np.random.randn(1, 1, 16, 64).astype(np.float32)
    v = np.random.randn(1, 1, 16, 64).astype(np.float32)
    o = aule.flash_attention(q, np.repeat(k, 4, axis=1), np.repeat(v, 4, axis=1), causal=True)
    outs["gqa_mqa_4to1_1x4x16x64"] = o
    meta["cases"]["gqa_mqa_4to1_1x4x16x64"] = {"shape": [1, 4, 16, 64], "kv_heads": 1, "causal": True,
                                                "sum": float(o.sum(dtype=np.float64))}
    # cross attention Sq=16, Sk=32 non-causal (tests/test_cross_attn.py:12-60)
    np.random.seed(42)
    q = np.random.randn(1, 4, 16, 64).astype(np.float32)
    k = np.random.randn(1, 4, 32, 64).astype(np.float32)
    v = np.random.randn(1, 4, 32, 64).astype(np.float32)
    o = aule.flash_attention(q, k, v, causal=False)
    outs["cross_1x4_q16_k32_d64"] = o
    meta["cases"]["cross_1x4_q16_k32_d64"] = {"shape_q": [1, 4, 16, 64], "seq_k": 32, "causal": False,
                                               "sum": float(o.sum(dtype=np.float64))}
    np.savez_compressed(os.path.join(HERE, "reference_numpy_path.npz"), **outs)
    with open(os.path.join(HERE, "reference_numpy_path.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta["cases"]["configA_1x8x256x64_causal"]))


if __name__ == "__main__":
    main()
