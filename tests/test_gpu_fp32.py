"""GPU parity, fp32 path (CUDA-core kernels) -- reads like the reference's own tests
(python/tests/test_cpu.py, test_vulkan.py, tests/test_gqa_unit.py, test_cross_attn.py),
with the oracle / committed reference golden vectors as the expected values."""
import ctypes

import numpy as np
import pytest

from conftest import ref_inputs
from oracle import attention_oracle as orc
from oracle import c_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def aule():
    import aule
    assert aule.get_available_backends() == ["cuda"], aule.get_backend_errors()
    return aule


def test_launch_path_smoke_multiply(aule):
    """tests/test_multiply.zig analogue: module load -> launch -> copies."""
    from aule import ffi
    lib = ffi.ensure_init()
    x = np.arange(1000, dtype=np.float32)
    y = np.empty_like(x)
    fp = ctypes.POINTER(ctypes.c_float)
    assert lib.aule_smoke_multiply(x.ctypes.data_as(fp), y.ctypes.data_as(fp), x.size) == 0, ffi.last_error()
    assert np.array_equal(y, 2 * x)
    assert lib.aule_last_kernel() == b"aule_smoke_multiply"


GOLD = [("cpu_causal_1x4x32x64", (1, 4, 32, 64), True), ("cpu_noncausal_1x4x32x64", (1, 4, 32, 64), False),
        ("cpu_batch_4x8x64x64", (4, 8, 64, 64), True), ("cpu_d32_1x4x32x32", (1, 4, 32, 32), True),
        ("cpu_d128_1x4x32x128", (1, 4, 32, 128), True), ("triton_shape_1x8x64x64_causal", (1, 8, 64, 64), True),
        ("triton_shape_1x8x64x64_noncausal", (1, 8, 64, 64), False),
        ("triton_shape_1x8x64x128_causal", (1, 8, 64, 128), True),
        ("configA_1x8x256x64_causal", (1, 8, 256, 64), True)]


@pytest.mark.parametrize("name,shape,causal", GOLD)
def test_numpy_entry_matches_reference_golden(aule, golden, name, shape, causal):
    """aule.flash_attention(numpy) vs outputs of the unmodified reference; tolerance of
    python/tests/test_cpu.py:29 (rtol=1e-4, atol=1e-4)."""
    q, k, v = ref_inputs(*shape)
    out = aule.flash_attention(q, k, v, causal=causal)
    assert out.shape == q.shape and out.dtype == np.float32
    np.testing.assert_allclose(out, golden[name], rtol=1e-4, atol=1e-4)
    assert orc.max_abs_diff(out, golden[name]) < 2e-5


@pytest.mark.parametrize("name,shape,causal", GOLD[:4])
def test_ffi_class_and_handle_api(aule, golden, name, shape, causal):
    """Aule.attention and the tensor-handle path (vulkan.py:613-815), test_vulkan.py tolerance 1e-3."""
    q, k, v = ref_inputs(*shape)
    with aule.Aule() as a:
        out = a.attention(q, k, v, causal=causal)
        np.testing.assert_allclose(out, golden[name], rtol=1e-3, atol=1e-3)
        n0 = a.tensor_count
        tq, tk, tv, to = (a.tensor(shape) for _ in range(4))
        assert a.tensor_count == n0 + 4
        tq.upload(q); tk.upload(k); tv.upload(v)
        a.attention_gpu(tq, tk, tv, to, causal=causal)
        np.testing.assert_allclose(to.download(), golden[name], rtol=1e-3, atol=1e-3)
        np.testing.assert_array_equal(tq.download(), q)
        for t in (tq, tk, tv, to):
            t.destroy()
        assert a.tensor_count == n0
        assert a.vendor == "nvidia" and a.supports_backward and "B200" in a.device_name


def test_zig_known_answers(aule):
    """attention_ref.zig:250-298 and tests/test_attention.zig:158-270 on the GPU."""
    q = np.full((1, 1, 2, 4), 0.5, np.float32)
    v = np.array([[1, 2, 3, 4], [5, 6, 7, 8]], np.float32).reshape(1, 1, 2, 4)
    out = aule.flash_attention(q, q, v, causal=False)
    assert np.abs(out[0, 0] - np.array([3, 4, 5, 6], np.float32)).max() < 1e-3
    q = np.full((1, 1, 4, 8), 0.5, np.float32)
    v = (np.arange(4)[:, None] * 8 + np.arange(8)[None, :]).astype(np.float32).reshape(1, 1, 4, 8)
    out = aule.flash_attention(q, q, v, causal=False)
    assert np.abs(out[0, 0] - v[0, 0].mean(axis=0)[None]).max() < 0.01
    S = D = 8
    q = np.zeros((1, 1, S, D), np.float32)
    q[0, 0, np.arange(S), np.arange(S)] = 10.0
    v = (0.1 * np.arange(S * D, dtype=np.float32)).reshape(1, 1, S, D)
    assert np.abs(aule.flash_attention(q, q, v, causal=False) - v).max() < 0.1


def test_zig_shapes_vs_c_restatement(aule):
    """tests/test_attention.zig:18-31,80-156: shapes up to [2,8,64,64], uniform(-0.5,0.5),
    tolerance max_abs<1e-4 OR max_rel<1e-3 (:60-77), expected = C restatement of attention_ref.zig."""
    for seed, shape in ((42, (1, 1, 16, 16)), (123, (1, 4, 32, 32)), (456, (2, 8, 64, 64)), (42, (1, 2, 48, 64))):
        rng = np.random.RandomState(seed)
        q, k, v = (rng.uniform(-0.5, 0.5, shape).astype(np.float32) for _ in range(3))
        for causal in (False, True):
            out = aule.flash_attention(q, k, v, causal=causal)
            exp = c_ref.forward(q, k, v, causal)
            assert orc.max_abs_diff(out, exp) < 1e-4 or orc.max_rel_diff(out, exp) < 1e-3


def test_stability_and_batch_independence(aule):
    """tests/test_attention.zig:272-384."""
    rng = np.random.RandomState(7)
    q, k, v = (rng.uniform(-5, 5, (2, 4, 64, 32)).astype(np.float32) for _ in range(3))
    o2 = aule.flash_attention(q, k, v, causal=False)
    assert np.isfinite(o2).all()
    o1 = aule.flash_attention(q[:1], k[:1], v[:1], causal=False)
    assert np.abs(o1[0] - o2[0]).max() < 1e-5


@pytest.mark.parametrize("B,Hq,Hkv,Sq,Sk,D,causal", [
    (1, 4, 1, 16, 16, 64, True),      # tests/test_gqa_unit.py (MQA 4/1)
    (1, 12, 2, 64, 64, 64, True),     # test_triton.py:96-110 (GQA 12/2)
    (1, 8, 1, 64, 64, 64, True),      # test_triton.py:112-126 (MQA 8/1)
    (1, 4, 4, 16, 32, 64, False),     # tests/test_cross_attn.py
    (2, 6, 3, 37, 53, 48, False),     # ragged everything
    (2, 6, 2, 70, 70, 20, True),      # D not a multiple of 32
    (1, 2, 2, 4, 256, 32, True),      # top-left causal with Sq << Sk
    (1, 2, 1, 300, 129, 128, True),   # Sq > Sk causal
])
def test_gqa_cross_ragged_vs_oracle(aule, B, Hq, Hkv, Sq, Sk, D, causal):
    import torch
    q, k, v = ref_inputs(B, Hq, Sq, D, Hkv=Hkv, Sk=Sk)
    exp, exp_lse = orc.attention_ref(q, k, v, causal=causal)
    tq, tk, tv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    out = aule.flash_attention(tq, tk, tv, causal=causal)
    assert out.dtype == torch.float32 and out.shape == tq.shape
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=1e-3, atol=1e-5)     # north_star: <=1e-3 rel fp32
    from aule import cuda_flash, ffi
    o2, lse = cuda_flash.forward_with_lse(tq, tk, tv, causal=causal)
    assert ffi.load_library().aule_last_kernel() == b"aule_fwd_simt_f32"
    np.testing.assert_allclose(lse.cpu().numpy(), exp_lse, rtol=1e-4, atol=1e-4)
    # numpy entry agrees with the device-pointer entry
    np.testing.assert_allclose(aule.flash_attention(q, k, v, causal=causal), o2.cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_scale_argument_is_honoured(aule):
    """triton_flash.py:394-395 / SURVEY 8a: the GPU path honours `scale`."""
    import torch
    q, k, v = ref_inputs(1, 2, 40, 32)
    exp, _ = orc.attention_ref(q, k, v, causal=True, scale=0.05)
    out = aule.flash_attention(*(torch.from_numpy(x).cuda() for x in (q, k, v)), causal=True, scale=0.05)
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("window", [1, 8, 33])
def test_sliding_window(aule, window):
    """window semantics of attention_f32.comp:176-178 (keep i - j < W), tests/test_sliding_window.py tol 1e-2."""
    import torch
    q, k, v = ref_inputs(1, 2, 96, 32)
    exp, _ = orc.attention_ref(q, k, v, causal=True, window=window)
    out = aule.flash_attention(*(torch.from_numpy(x).cuda() for x in (q, k, v)), causal=True, window_size=window)
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("B,Hq,Hkv,Sq,Sk,D,causal", [
    (1, 8, 8, 64, 64, 64, True),      # test_triton.py:66-94
    (2, 4, 2, 50, 50, 32, True),
    (1, 4, 1, 33, 65, 48, False),
    (1, 2, 2, 40, 40, 128, True),
])
def test_backward_vs_oracle(aule, B, Hq, Hkv, Sq, Sk, D, causal):
    """dQ/dK/dV through autograd of aule.flash_attention vs the analytic oracle;
    tolerance of python/tests/test_triton.py:92-94 is 1e-2, fp32 does far better."""
    import torch
    q, k, v = ref_inputs(B, Hq, Sq, D, Hkv=Hkv, Sk=Sk)
    do = np.random.RandomState(1).randn(B, Hq, Sq, D).astype(np.float32)
    dq, dk, dv, _, _ = orc.attention_bwd_ref(q, k, v, do, causal=causal)
    tq, tk, tv = (torch.from_numpy(x).cuda().requires_grad_() for x in (q, k, v))
    out = aule.flash_attention(tq, tk, tv, causal=causal)
    out.backward(torch.from_numpy(do).cuda())
    for g, e in ((tq.grad, dq), (tk.grad, dk), (tv.grad, dv)):
        assert g.shape == e.shape
        np.testing.assert_allclose(g.cpu().numpy(), e, rtol=1e-3, atol=2e-5)


def test_legacy_training_abi(aule):
    """vulkan.py:824-962: forward_with_lse + backward on host fp32 arrays."""
    q, k, v = ref_inputs(2, 4, 48, 64)
    do = np.random.RandomState(3).randn(*q.shape).astype(np.float32)
    dq, dk, dv, o, lse = orc.attention_bwd_ref(q, k, v, do, causal=True)
    with aule.Aule() as a:
        out, l = a.attention_forward_with_lse(q, k, v, causal=True)
        np.testing.assert_allclose(out, o, rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(l, lse, rtol=1e-4, atol=1e-4)
        gq, gk, gv = a.attention_backward(q, k, v, out, do, l, causal=True)
    for g, e in ((gq, dq), (gk, dk), (gv, dv)):
        np.testing.assert_allclose(g, e, rtol=1e-3, atol=2e-5)


def test_errors_are_loud(aule):
    import torch
    q = torch.zeros(1, 2, 8, 16, device="cuda")
    with pytest.raises(ValueError, match="head_dim"):
        aule.flash_attention(torch.zeros(1, 2, 8, 130, device="cuda"), torch.zeros(1, 2, 8, 130, device="cuda"),
                             torch.zeros(1, 2, 8, 130, device="cuda"))
    with aule.Aule() as a:
        with pytest.raises(ValueError):
            a.tensor((1, 2, 3))
        t = a.tensor((1, 1, 4, 4))
        with pytest.raises(ValueError, match="Shape mismatch"):
            t.upload(np.zeros((1, 1, 4, 8), np.float32))
        t.destroy()
    assert q is not None
