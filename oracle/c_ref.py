"""ctypes loader for the C restatement of src/attention_ref.zig (oracle/attention_ref.c).
TEST INFRASTRUCTURE ONLY -- see attention_ref.c for the file:line map."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libaule_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        for name in ("aule_ref_forward", "aule_ref_forward_causal"):
            f = getattr(L, name)
            f.argtypes = [fp, fp, fp, fp] + [ctypes.c_uint32] * 4
            f.restype = ctypes.c_int
        L.aule_ref_compare_arrays.argtypes = [fp, fp, ctypes.c_size_t, ctypes.c_float]
        L.aule_ref_compare_arrays.restype = ctypes.c_int
        L.aule_ref_max_abs_diff.argtypes = [fp, fp, ctypes.c_size_t]
        L.aule_ref_max_abs_diff.restype = ctypes.c_float
        L.aule_ref_mean_abs_diff.argtypes = [fp, fp, ctypes.c_size_t]
        L.aule_ref_mean_abs_diff.restype = ctypes.c_float
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def forward(q, k, v, causal):
    """AttentionRef.forward / forwardCausal on fp32 [B,H,S,D] MHA, Sq == Sk."""
    q = np.ascontiguousarray(q, np.float32)
    k = np.ascontiguousarray(k, np.float32)
    v = np.ascontiguousarray(v, np.float32)
    assert q.shape == k.shape == v.shape and q.ndim == 4
    B, H, S, D = q.shape
    out = np.empty_like(q)
    fn = lib().aule_ref_forward_causal if causal else lib().aule_ref_forward
    rc = fn(_p(q), _p(k), _p(v), _p(out), B, H, S, D)
    if rc != 0:
        raise MemoryError("oracle scratch allocation failed")
    return out


def max_abs_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).ravel()
    b = np.ascontiguousarray(b, np.float32).ravel()
    return float(lib().aule_ref_max_abs_diff(_p(a), _p(b), a.size))
