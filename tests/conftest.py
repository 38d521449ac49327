"""Shared fixtures. Mirrors the reference's fixture recipe
(/root/reference/python/tests/conftest.py:8-31) so the parity tests read like
the reference's own tests."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_PY = os.path.join(ROOT, "aule-attention_b200", "python")
for p in (ROOT, PKG_PY):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def ref_inputs(B, H, S, D, Hkv=None, Sk=None, seed=42):
    """python/tests/conftest.py:8-16 : seed 42, q,k,v successive randn fp32."""
    np.random.seed(seed)
    q = np.random.randn(B, H, S, D).astype(np.float32)
    k = np.random.randn(B, Hkv or H, Sk or S, D).astype(np.float32)
    v = np.random.randn(B, Hkv or H, Sk or S, D).astype(np.float32)
    return q, k, v


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_numpy_path.npz")
    return dict(np.load(path))
