"""GPU parity for what round 2 added to the hot path (all through the C ABI of libaule.so):
padded head dims and bidirectional windows on the tcgen05 kernel, RoPE (both conventions, fused entry, handle ABI),
rows without a visible key, thread safety and pageable host buffers, the native spanning call, and the boundary proof:
the reference's OWN vulkan.py (unmodified, pip-installed under baseline/_ref) driving this library."""
import ctypes
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from conftest import ROOT, ref_inputs
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
BF16_TOL = 1e-2          # north_star: <= 1e-2 rel bf16
FP32_TOL = 1e-3          # north_star: <= 1e-3 rel fp32


@pytest.fixture(scope="module")
def aule():
    import aule
    assert aule.get_available_backends() == ["cuda"], aule.get_backend_errors()
    return aule


@pytest.fixture(scope="module")
def golden_r2():
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_triton_path_r2.npz")))
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_triton_path_r2.json")))
    return d, meta


def _fwd(q, k, v, causal, dtype, window=-1, scale=None):
    import torch
    from aule import cuda_flash, ffi
    td = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[dtype]
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(x)).cuda().to(td) for x in (q, k, v))
    out, lse = cuda_flash.forward_with_lse(tq, tk, tv, causal=causal, scale=scale, window_size=window)
    torch.cuda.synchronize()
    kern = ffi.load_library().aule_last_kernel().decode()
    return out.float().cpu().numpy(), lse.cpu().numpy(), tuple(t.float().cpu().numpy() for t in (tq, tk, tv)), kern


# ------------------------------------------------------------------ padded head dims on the tensor-core kernel
@pytest.mark.parametrize("D", [8, 32, 40, 56, 72, 80, 96, 120])
@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("B,Hq,Hkv,Sq,Sk,causal", [(1, 4, 4, 200, 200, True), (2, 8, 2, 384, 384, True), (1, 2, 2, 100, 333, False)])
def test_padded_head_dim_runs_on_tensor_cores(aule, D, dtype, B, Hq, Hkv, Sq, Sk, causal):
    """The reference's GPU path takes any D <= 128 by padding to BLOCK_K = next_power_of_2(D) with load masks
    (triton_flash.py:446,:111,:161); here the 64-wide TMA boxes zero-fill D up to 64 / 128."""
    q, k, v = ref_inputs(B, Hq, Sq, D, Hkv=Hkv, Sk=Sk)
    out, lse, (rq, rk, rv), kern = _fwd(q, k, v, causal, dtype)
    assert kern == f"aule_fwd_sm100_{dtype}_d{64 if D <= 64 else 128}", kern
    exp, exp_lse = orc.attention_ref(rq, rk, rv, causal=causal)
    assert np.isfinite(out).all()
    assert orc.rel_err_to_scale(out, exp) <= BF16_TOL, orc.rel_err_to_scale(out, exp)
    np.testing.assert_allclose(lse, exp_lse, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("case", ["d40_causal_1x4x96x40", "d80_gqa_1x4x80x80", "d96_noncausal_1x2x64x96", "d32_causal_1x4x64x32"])
def test_padded_head_dim_vs_reference_triton_golden(aule, golden_r2, case):
    """Outputs of the reference's own Triton kernel (TRITON_INTERPRET, tests/golden/gen_golden_r2.py) for D = 40 / 80 / 96 / 32."""
    g, meta = golden_r2
    m = meta["fwd"][case]
    q, k, v, exp = (g[f"{case}.{x}"] for x in ("q", "k", "v", "out"))
    out, _, (rq, rk, rv), kern = _fwd(q, k, v, m["causal"], "bf16")
    assert kern.startswith("aule_fwd_sm100_bf16"), kern
    assert orc.rel_err_to_scale(out, exp) <= 2e-2                       # golden made from the UNROUNDED fp32 inputs
    exp_r, _ = orc.attention_ref(rq, rk, rv, causal=m["causal"])        # ... and against the oracle on what the kernel saw
    assert orc.rel_err_to_scale(out, exp_r) <= BF16_TOL
    out32, _, _, kern32 = _fwd(q, k, v, m["causal"], "f32")             # fp32 path against the same golden
    assert kern32 == "aule_fwd_simt_f32"
    assert orc.rel_err_to_scale(out32, exp) <= FP32_TOL


# ------------------------------------------------------------------ bidirectional sliding window
@pytest.mark.parametrize("window", [1, 2, 12, 100, 256, 300, 5000])
@pytest.mark.parametrize("B,Hq,Hkv,Sq,Sk,D", [(1, 4, 2, 384, 384, 128), (1, 2, 2, 300, 500, 64), (1, 2, 2, 700, 200, 128)])
def test_bidirectional_window_tensor_core(aule, window, B, Hq, Hkv, Sq, Sk, D):
    """Non-causal window keeps |i-j| <= window//2 (attention_f32.comp:180-183), now on the tcgen05 kernel: blocks outside
    the band are skipped, edge blocks are masked on both sides; rows whose band holds no key give O = 0, LSE = -inf."""
    q, k, v = ref_inputs(B, Hq, Sq, D, Hkv=Hkv, Sk=Sk)
    out, lse, (rq, rk, rv), kern = _fwd(q, k, v, False, "bf16", window=window)
    assert kern == f"aule_fwd_sm100_bf16_d{D}", kern
    exp, exp_lse = orc.attention_ref(rq, rk, rv, causal=False, window=window)
    empty = ~np.isfinite(exp_lse)                                       # rows with no visible key (Sq > Sk + window//2)
    assert np.isfinite(out).all()
    assert (out[empty] == 0).all() and np.isneginf(lse[empty]).all()
    exp = np.where(empty[..., None], 0.0, exp)
    assert orc.rel_err_to_scale(out, exp) <= BF16_TOL, orc.rel_err_to_scale(out, exp)
    np.testing.assert_allclose(lse[~empty], exp_lse[~empty], rtol=2e-3, atol=2e-3)
    out32, _, _, kern32 = _fwd(rq, rk, rv, False, "f32", window=window)  # CUDA-core kernel, same convention
    assert kern32 == "aule_fwd_simt_f32"
    assert orc.rel_err_to_scale(out32, exp) <= FP32_TOL


def test_rows_without_visible_key_agree_across_kernels(aule):
    """Causal window with Sq > Sk + window: rows i >= Sk + W - 1 see nothing.  Tensor-core and CUDA-core kernels both
    return O = 0 and LSE = -inf for them (the tensor-core path used to write NaN / junk)."""
    Sq, Sk, W = 600, 130, 40
    q, k, v = ref_inputs(1, 2, Sq, 128, Hkv=2, Sk=Sk)
    out, lse, (rq, rk, rv), kern = _fwd(q, k, v, True, "bf16", window=W)
    assert kern == "aule_fwd_sm100_bf16_d128"
    out32, lse32, _, _ = _fwd(rq, rk, rv, True, "f32", window=W)
    dead = np.arange(Sq) >= Sk + W - 1
    assert dead.any()
    assert np.isfinite(out).all() and (out[:, :, dead] == 0).all() and (out32[:, :, dead] == 0).all()
    assert np.isneginf(lse[:, :, dead]).all() and np.isneginf(lse32[:, :, dead]).all()
    assert orc.rel_err_to_scale(out[:, :, ~dead], out32[:, :, ~dead]) <= BF16_TOL


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("causal,window", [(True, 24), (True, 200), (False, 50)])
def test_sliding_window_backward(aule, dtype, causal, window):
    """backward with a window used to raise.  The reference's own backward ignores the window
    (triton_flash.py:313-319), so the expected values are the oracle's analytic gradients with the mask applied."""
    import torch
    td = {"bf16": torch.bfloat16, "f32": torch.float32}[dtype]
    torch.manual_seed(4)
    q = torch.randn(1, 4, 300, 64, device="cuda").to(td).requires_grad_()
    k = torch.randn(1, 2, 300, 64, device="cuda").to(td).requires_grad_()
    v = torch.randn(1, 2, 300, 64, device="cuda").to(td).requires_grad_()
    out = aule.flash_attention(q, k, v, causal=causal, window_size=window)
    do = torch.randn_like(out)
    out.backward(do)
    rq, rk, rv, rdo = (t.detach().float().cpu().numpy() for t in (q, k, v, do))
    dq, dk, dv, o, _ = orc.attention_bwd_ref(rq, rk, rv, rdo, causal=causal, window=window)
    tol = BF16_TOL if dtype == "bf16" else FP32_TOL
    assert orc.rel_err_to_scale(out.detach().float().cpu().numpy(), o) <= tol
    for g_, e_ in ((q.grad, dq), (k.grad, dk), (v.grad, dv)):
        assert orc.rel_err_to_scale(g_.float().cpu().numpy(), e_) <= tol


@pytest.mark.parametrize("causal,window", [(True, 1), (True, 24), (True, 128), (True, 300), (False, 2), (False, 50), (False, 256), (False, 900)])
@pytest.mark.parametrize("B,Hq,Hkv,Sq,Sk,D", [(1, 4, 2, 700, 700, 128), (2, 2, 2, 520, 520, 64), (1, 2, 1, 300, 900, 128), (1, 2, 2, 900, 260, 80)])
def test_sliding_window_backward_on_tensor_cores(aule, causal, window, B, Hq, Hkv, Sq, Sk, D):
    """Windowed training runs the tcgen05 backward kernels: the visibility band i - win_left <= j <= i + win_right bounds the
    block loops of both kernels and becomes alive-bit masks on the edge blocks.  Shapes cover several 128-blocks per side,
    Sq != Sk (bidirectional only: causal needs Sq == Sk here), query rows that see no key at all (Sq >> Sk with a narrow
    window: zero gradients, LSE = -inf) and a padded head dim."""
    import torch
    from aule import ffi
    if causal and Sq != Sk:
        pytest.skip("causal windows are tested with Sq == Sk")
    q, k, v = ref_inputs(B, Hq, Sq, D, Hkv=Hkv, Sk=Sk)
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(x)).cuda().to(torch.bfloat16).requires_grad_() for x in (q, k, v))
    out = aule.flash_attention(tq, tk, tv, causal=causal, window_size=window)
    do = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(window)).to(torch.bfloat16)
    out.backward(do)
    torch.cuda.synchronize()
    assert ffi.load_library().aule_last_kernel().decode() == f"aule_bwd_dq_sm100_bf16_d{64 if D <= 64 else 128}"
    edq, edk, edv, eo, _ = orc.attention_bwd_ref(*(t.detach().float().cpu().numpy() for t in (tq, tk, tv, do)), causal=causal, window=window)
    assert orc.rel_err_to_scale(out.detach().float().cpu().numpy(), eo) <= BF16_TOL
    for g_, e_ in ((tq.grad, edq), (tk.grad, edk), (tv.grad, edv)):
        g = g_.float().cpu().numpy()
        assert np.isfinite(g).all()
        # (window 1: every row sees only itself, P = 1 and dS = dP - Delta = 0 analytically -- dQ, dK are exactly zero, so the
        #  error is measured against a floor instead of the tensor's own scale)
        assert np.abs(g - e_).max() <= BF16_TOL * max(np.abs(e_).max(), 0.05), (np.abs(g - e_).max(), np.abs(e_).max())


# ------------------------------------------------------------------ padded head dims: tensor-core backward
@pytest.mark.parametrize("D", [8, 40, 56, 80, 96, 120])
@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("B,Hq,Hkv,Sq,Sk,causal", [(1, 4, 2, 200, 200, True), (2, 2, 2, 384, 384, True), (1, 2, 1, 100, 333, False)])
def test_padded_head_dim_backward_on_tensor_cores(aule, D, dtype, B, Hq, Hkv, Sq, Sk, causal):
    """Training with head dims the reference pads to a power of two (triton_flash.py:446): the tensor-core backward runs at
    64 / 128 with the padding columns zero-filled by TMA (Q, K, V, dO), clipped on store (dK, dV) and predicated in the dQ
    kernel's own loads / stores.  Expected values: the oracle's analytic gradients."""
    import torch
    from aule import ffi
    td = {"bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    q, k, v = ref_inputs(B, Hq, Sq, D, Hkv=Hkv, Sk=Sk)
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(x)).cuda().to(td).requires_grad_() for x in (q, k, v))
    out = aule.flash_attention(tq, tk, tv, causal=causal)
    do = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(D)).to(td)
    out.backward(do)
    torch.cuda.synchronize()
    assert ffi.load_library().aule_last_kernel().decode() == f"aule_bwd_dq_sm100_{dtype}_d{64 if D <= 64 else 128}"
    edq, edk, edv, _, _ = orc.attention_bwd_ref(*(t.detach().float().cpu().numpy() for t in (tq, tk, tv, do)), causal=causal)
    for g_, e_ in ((tq.grad, edq), (tk.grad, edk), (tv.grad, edv)):
        g = g_.float().cpu().numpy()
        assert np.isfinite(g).all()
        assert orc.rel_err_to_scale(g, e_) <= BF16_TOL, orc.rel_err_to_scale(g, e_)


# ------------------------------------------------------------------ fp32 inputs on the tensor cores (tf32)
@pytest.mark.parametrize("B,Hq,Hkv,Sq,Sk,D,causal,window", [
    (2, 8, 8, 512, 512, 64, True, -1),          # config A's shape class (fp32 causal MHA, D = 64)
    (4, 32, 32, 2048, 2048, 64, True, -1),      # fp32 [4,32,2048,64] within 1e-3
    (1, 8, 2, 300, 300, 64, True, -1),          # GQA, ragged
    (1, 4, 4, 200, 333, 32, False, -1),         # padded head dim, cross attention
    (1, 2, 1, 130, 700, 40, False, -1),
    (1, 4, 2, 384, 384, 64, True, 100),         # causal window
    (1, 2, 2, 300, 500, 48, False, 64),         # bidirectional window, padded
    (1, 2, 2, 700, 200, 64, False, 2),          # rows without a visible key
])
def test_fp32_inputs_tf32_tensor_core_forward(aule, B, Hq, Hkv, Sq, Sk, D, causal, window):
    """The reference's GPU path runs fp32 inputs through tl.dot, i.e. as tf32 (triton_flash.py:405-411).  With tf32 allowed the
    fp32 forward runs the kind::tf32 tcgen05 kernel (head_dim <= 64): error <= 1e-3 relative to the output scale (north_star's
    fp32 bar) against the fp64 oracle; without it the exact CUDA-core kernel runs."""
    import torch
    from aule import cuda_flash, ffi
    lib = ffi.load_library()
    q, k, v = ref_inputs(B, Hq, Sq, D, Hkv=Hkv, Sk=Sk)
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (q, k, v))
    out, lse = cuda_flash.forward_with_lse(tq, tk, tv, causal=causal, window_size=window, allow_tf32=True)
    torch.cuda.synchronize()
    assert lib.aule_last_kernel().decode() == "aule_fwd_sm100_tf32_d64"
    if B * Hq * Sq * Sk <= 1 << 28:
        exp, exp_lse = orc.attention_ref(q, k, v, causal=causal, window=window)
    else:                                                       # full-size case: the exact fp32 kernel is the yardstick
        e32, l32 = cuda_flash.forward_with_lse(tq, tk, tv, causal=causal, window_size=window, allow_tf32=False)
        assert lib.aule_last_kernel().decode() == "aule_fwd_simt_f32"
        exp, exp_lse = e32.cpu().numpy(), l32.cpu().numpy()
    o = out.cpu().numpy()
    assert np.isfinite(o).all()
    live = np.isfinite(exp_lse)                                  # rows without a visible key: O = 0, LSE = -inf (the oracle has NaN there)
    assert (o[~live] == 0).all()
    assert orc.rel_err_to_scale(o[live], exp[live]) <= FP32_TOL, orc.rel_err_to_scale(o[live], exp[live])
    assert (np.isneginf(lse.cpu().numpy()) == ~live).all()
    np.testing.assert_allclose(lse.cpu().numpy()[live], exp_lse[live], rtol=2e-3, atol=2e-3)


def test_fp32_tf32_is_opt_in(aule):
    """fp32 tensors stay on the exact fp32 kernel unless tf32 is allowed (aule.set_fp32_tf32 / AULE_TF32); head_dim > 64 always does."""
    import torch
    from aule import ffi
    lib = ffi.load_library()
    q = torch.randn(1, 4, 256, 64, device="cuda")
    aule.flash_attention(q, q, q, causal=True)
    assert lib.aule_last_kernel().decode() == "aule_fwd_simt_f32"
    aule.set_fp32_tf32(True)
    try:
        o = aule.flash_attention(q, q, q, causal=True)
        assert lib.aule_last_kernel().decode() == "aule_fwd_sm100_tf32_d64" and o.dtype == torch.float32
        q128 = torch.randn(1, 2, 256, 128, device="cuda")
        aule.flash_attention(q128, q128, q128, causal=True)
        assert lib.aule_last_kernel().decode() == "aule_fwd_simt_f32"
        qg = q.clone().requires_grad_()                        # training: tf32 forward, exact fp32 backward
        aule.flash_attention(qg, qg, qg, causal=True).sum().backward()
        assert torch.isfinite(qg.grad).all()
    finally:
        aule.set_fp32_tf32(None)


# ------------------------------------------------------------------ fused backward (opt-in)
@pytest.mark.parametrize("B,Hq,Hkv,Sq,Sk,causal,dtype", [
    (1, 2, 1, 256, 256, True, "bf16"), (2, 4, 2, 1000, 1000, True, "f16"), (1, 2, 1, 300, 520, False, "bf16"), (1, 8, 2, 1536, 1536, True, "bf16")])
def test_fused_backward_kernel_vs_oracle(aule, B, Hq, Hkv, Sq, Sk, causal, dtype):
    """aule_set_kernel_path bit 17: ONE backward kernel (attn_bwd_fused_sm100.cu) that reduces dQ partials into an fp32
    accumulator with red.global.add, the way the reference does (tl.atomic_add, triton_flash.py:335-339).  dK / dV must be
    bit-identical to the deterministic two-kernel backward (same operands, same accumulation order); dQ within the bf16 bar."""
    import torch
    from aule import cuda_flash, ffi
    lib = ffi.ensure_init()
    td = {"bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    code = {"bf16": ffi.DTYPE_BF16, "f16": ffi.DTYPE_F16}[dtype]
    q, k, v = ref_inputs(B, Hq, Sq, 128, Hkv=Hkv, Sk=Sk)
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(x)).cuda().to(td) for x in (q, k, v))
    o, lse = cuda_flash.forward_with_lse(tq, tk, tv, causal=causal)
    do = torch.randn(o.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)).to(td)
    res = {}
    try:
        for path in (0, 1 << 17):
            dq, dk, dv = (torch.full_like(t, float("nan")) for t in (tq, tk, tv))
            lib.aule_set_kernel_path(path)
            rc = lib.aule_attention_backward_dptr(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                                  dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, Hq, Hkv, Sq, Sk, 128, code, 0.0,
                                                  1 if causal else 0, 0, torch.cuda.current_stream().cuda_stream)
            assert rc == 0, ffi.last_error()
            torch.cuda.synchronize()
            res[path] = (dq, dk, dv, lib.aule_last_kernel().decode())
    finally:
        lib.aule_set_kernel_path(0)
    assert res[0][3].startswith("aule_bwd_dq_sm100") and res[1 << 17][3].startswith("aule_bwd_dq_convert"), (res[0][3], res[1 << 17][3])
    assert torch.equal(res[0][1], res[1 << 17][1]) and torch.equal(res[0][2], res[1 << 17][2])
    edq, edk, edv, _, _ = orc.attention_bwd_ref(*(t.float().cpu().numpy() for t in (tq, tk, tv, do)), causal=causal)
    for g_, e_ in zip(res[1 << 17][:3], (edq, edk, edv)):
        assert orc.rel_err_to_scale(g_.float().cpu().numpy(), e_) <= BF16_TOL


@pytest.mark.parametrize("extra", [0, 1 << 28])
@pytest.mark.parametrize("B,Hq,Hkv,Sq,Sk,causal,dtype", [
    (1, 2, 1, 256, 256, True, "bf16"), (2, 4, 2, 1000, 1000, True, "f16"), (1, 2, 1, 300, 520, False, "bf16"), (1, 2, 2, 200, 200, True, "bf16"),
    (1, 8, 2, 2048, 2048, True, "bf16")])
def test_fused2_backward_kernel_vs_oracle(aule, B, Hq, Hkv, Sq, Sk, causal, dtype, extra):
    """aule_set_kernel_path bit 26 (experimental, opt-in): the fused backward in 64-query half steps whose dQ partials leave
    through TMA bulk reductions (attn_bwd_fused2_sm100.cu); bit 28 adds the fence-helper warp.  dK is bit-identical to the
    two-kernel backward (same operands, same order); dV accumulates in 64-query steps and dQ through fp32 reductions, so both
    are held to the bf16 bar against the fp64 oracle, like the reference's own atomics (triton_flash.py:335-347)."""
    import torch
    from aule import cuda_flash, ffi
    lib = ffi.ensure_init()
    td = {"bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    code = {"bf16": ffi.DTYPE_BF16, "f16": ffi.DTYPE_F16}[dtype]
    q, k, v = ref_inputs(B, Hq, Sq, 128, Hkv=Hkv, Sk=Sk)
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(x)).cuda().to(td) for x in (q, k, v))
    o, lse = cuda_flash.forward_with_lse(tq, tk, tv, causal=causal)
    do = torch.randn(o.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)).to(td)
    fused2 = (1 << 26) | extra
    res = {}
    try:
        for path in (0, fused2):
            dq, dk, dv = (torch.full_like(t, float("nan")) for t in (tq, tk, tv))
            lib.aule_set_kernel_path(path)
            rc = lib.aule_attention_backward_dptr(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                                  dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, Hq, Hkv, Sq, Sk, 128, code, 0.0,
                                                  1 if causal else 0, 0, torch.cuda.current_stream().cuda_stream)
            assert rc == 0, ffi.last_error()
            torch.cuda.synchronize()
            res[path] = (dq, dk, dv, lib.aule_last_kernel().decode())
    finally:
        lib.aule_set_kernel_path(0)
    assert res[0][3].startswith("aule_bwd_dq_sm100") and res[fused2][3].startswith("aule_bwd_dq_convert"), (res[0][3], res[fused2][3])
    assert torch.equal(res[0][1], res[fused2][1])
    edq, edk, edv, _, _ = orc.attention_bwd_ref(*(t.float().cpu().numpy() for t in (tq, tk, tv, do)), causal=causal)
    for g_, e_ in zip(res[fused2][:3], (edq, edk, edv)):
        assert bool(torch.isfinite(g_).all())
        assert orc.rel_err_to_scale(g_.float().cpu().numpy(), e_) <= BF16_TOL


# ------------------------------------------------------------------ RoPE
@pytest.mark.parametrize("case", ["rope_gqa_1x4x64x64", "rope_mha_1x2x48x128"])
def test_rope_vs_reference_golden(aule, golden_r2, case):
    """flash_attention_rope against the reference's own formulation: apply_rope_separate (triton_flash.py:680-703) followed by
    its attention kernel (fixture made by tests/golden/gen_golden_r2.py; the reference's FUSED rope kernel does not run under
    the Triton interpreter -- recorded in the fixture's json)."""
    import torch
    g, meta = golden_r2
    q, k, v, cos, sin, q_rot, k_rot, exp = (g[f"{case}.{x}"] for x in ("q", "k", "v", "cos", "sin", "q_rot", "k_rot", "out"))
    from aule import cuda_flash
    tq, tk, tv, tc, ts = (torch.from_numpy(x).cuda() for x in (q, k, v, cos, sin))
    np.testing.assert_allclose(cuda_flash.apply_rope(tq, tc, ts).cpu().numpy(), q_rot, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cuda_flash.apply_rope(tk, tc, ts).cpu().numpy(), k_rot, rtol=1e-5, atol=1e-5)
    with torch.no_grad():
        out = aule.flash_attention_rope(tq, tk, tv, tc, ts, causal=meta["rope"][case]["causal"])       # one C call
    assert orc.rel_err_to_scale(out.cpu().numpy(), exp) <= FP32_TOL
    with torch.no_grad():
        out16 = aule.flash_attention_rope(tq.bfloat16(), tk.bfloat16(), tv.bfloat16(), tc, ts, causal=True)
    from aule import ffi
    assert ffi.load_library().aule_last_kernel().decode().startswith("aule_fwd_sm100_bf16")
    assert orc.rel_err_to_scale(out16.float().cpu().numpy(), exp) <= 2e-2


@pytest.mark.parametrize("interleaved", [False, True])
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_rope_both_conventions_forward_backward(aule, interleaved, dtype):
    import torch
    td = {"bf16": torch.bfloat16, "f32": torch.float32}[dtype]
    B, Hq, Hkv, Sq, Sk, D = 2, 4, 2, 96, 160, 64                        # Sk > Sq: the same table serves both
    cos, sin = aule.precompute_rope_frequencies(Sk, D)
    torch.manual_seed(0)
    q = torch.randn(B, Hq, Sq, D, device="cuda").to(td).requires_grad_()
    k = torch.randn(B, Hkv, Sk, D, device="cuda").to(td).requires_grad_()
    v = torch.randn(B, Hkv, Sk, D, device="cuda").to(td).requires_grad_()
    out = aule.flash_attention_rope(q, k, v, cos, sin, causal=False, interleaved=interleaved)
    do = torch.randn_like(out)
    out.backward(do)
    cn, sn = cos.cpu().numpy(), sin.cpu().numpy()
    rq, rk, rv, rdo = (t.detach().float().cpu().numpy() for t in (q, k, v, do))
    qr, kr = orc.rope_ref(rq, cn, sn, interleaved), orc.rope_ref(rk, cn, sn, interleaved)
    exp, _ = orc.attention_ref(qr, kr, rv, causal=False)
    tol = 2e-2 if dtype == "bf16" else FP32_TOL                          # bf16: q,k are rounded once more after the rotation
    assert orc.rel_err_to_scale(out.detach().float().cpu().numpy(), exp) <= tol
    dqr, dkr, dv, _, _ = orc.attention_bwd_ref(qr, kr, rv, rdo, causal=False)
    dq, dk = orc.rope_ref(dqr, cn, -sn, interleaved), orc.rope_ref(dkr, cn, -sn, interleaved)
    for g_, e_ in ((q.grad, dq), (k.grad, dk), (v.grad, dv)):
        assert orc.rel_err_to_scale(g_.float().cpu().numpy(), e_) <= tol
    with torch.no_grad():                                                # the fused single-call path gives the same forward
        out2 = aule.flash_attention_rope(q.detach(), k.detach(), v.detach(), cos, sin, causal=False, interleaved=interleaved)
    assert orc.rel_err_to_scale(out2.float().cpu().numpy(), exp) <= tol


def test_rope_table_validation(aule):
    """a [S, D] table or a table shorter than the sequence is an error, never a reinterpretation / OOB read."""
    import torch
    from aule import cuda_flash, ffi
    q = torch.randn(1, 2, 64, 64, device="cuda")
    cos, sin = aule.precompute_rope_frequencies(64, 64)
    with pytest.raises(ValueError, match="head_dim//2"):
        cuda_flash.apply_rope(q, torch.cat([cos, cos], -1), torch.cat([sin, sin], -1))
    with pytest.raises(ValueError, match="rows"):
        cuda_flash.apply_rope(q, cos[:32], sin[:32])
    with pytest.raises(ValueError, match="rows"):                        # K longer than the table
        aule.flash_attention_rope(q, torch.randn(1, 2, 128, 64, device="cuda"), torch.randn(1, 2, 128, 64, device="cuda"), cos, sin)
    lib = ffi.load_library()
    out = torch.empty_like(q)
    rc = lib.aule_rope_dptr(q.data_ptr(), out.data_ptr(), cos.data_ptr(), sin.data_ptr(), 1, 2, 64, 64, 32, 0, ffi.DTYPE_F32, 0, 0,
                            torch.cuda.current_stream().cuda_stream)
    assert rc == -4 and "rows" in ffi.last_error()


def test_rope_through_the_handle_abi(aule):
    """tests/test_rope_unit.py of the reference: Aule().attention_gpu(q, k, v, out, rot_cos, rot_sin) -> lib.zig:496-529,
    the shader's interleaved pairs (attention_f32.comp:98-111); shapes and tolerances of that test (1e-3)."""
    B, H, S, D = 1, 1, 8, 64
    np.random.seed(42)
    q, k, v = (np.random.randn(B, H, S, D).astype(np.float32) for _ in range(3))
    half = D // 2
    freqs = 1.0 / (10000 ** (np.arange(0, half, dtype=np.float32) / half))
    emb = np.outer(np.arange(S, dtype=np.float32), freqs)
    cos, sin = np.cos(emb).astype(np.float32), np.sin(emb).astype(np.float32)
    with aule.Aule() as ctx:
        n0 = ctx.tensor_count
        tq, tk, tv, to, to2 = (ctx.tensor(q.shape) for _ in range(5))
        tc, ts = ctx.tensor((1, 1, S, half)), ctx.tensor((1, 1, S, half))
        for t, a in ((tq, q), (tk, k), (tv, v), (tc, cos.reshape(1, 1, S, half)), (ts, sin.reshape(1, 1, S, half))):
            t.upload(a)
        ctx.attention_gpu(tq, tk, tv, to, rot_cos=tc, rot_sin=ts, causal=False)
        ctx.attention_gpu(tq, tk, tv, to2, rot_cos=None, rot_sin=None, causal=False)
        out_rope, out_plain = to.download(), to2.download()
        with pytest.raises(aule.AuleError):
            ctx.attention_gpu(tq, tk, tv, to, rot_cos=tc, rot_sin=None, causal=False)
        assert ctx.tensor_count == n0 + 7
    assert aule.Aule().tensor_count == n0                               # close() released every handle
    exp_plain, _ = orc.attention_ref(q, k, v, causal=False)
    exp_rope, _ = orc.attention_ref(orc.rope_ref(q, cos, sin, True), orc.rope_ref(k, cos, sin, True), v, causal=False)
    np.testing.assert_allclose(out_plain, exp_plain, atol=1e-3, rtol=1e-3)
    np.testing.assert_allclose(out_rope, exp_rope, atol=1e-3, rtol=1e-3)
    assert np.abs(out_rope - out_plain).max() > 1e-2
    # the public NumPy entry with rot_cos / rot_sin takes the same route
    out_np = aule.flash_attention(q, k, v, rot_cos=cos, rot_sin=sin, causal=False)
    np.testing.assert_allclose(out_np, exp_rope, atol=1e-3, rtol=1e-3)


# ------------------------------------------------------------------ element-wise parity in the reference's form
def test_config_c_samples_elementwise(aule):
    """beside the max-norm gate, the reference's own element-wise form (tests/test_attention.zig:60-77:
    |d| < atol OR |d|/max(|ref|, floor) < rtol) on sampled (batch, head, row-block) triples of config C."""
    import torch
    from aule import cuda_flash
    B, Hq, Hkv, S, D = 8, 32, 8, 4096, 128
    g = torch.Generator(device="cuda").manual_seed(42)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    out, lse = cuda_flash.forward_with_lse(q, k, v, causal=True)
    for (b, h, r0) in [(0, 0, 0), (3, 17, 1920), (7, 31, 3968), (5, 6, 2048)]:
        hk = h // (Hq // Hkv)
        qs = q[b, h].float().cpu().numpy()[None, None]
        ks, vs = (t[b, hk].float().cpu().numpy()[None, None] for t in (k, v))
        exp, exp_lse = orc.attention_rows(qs, ks, vs, 0, 0, r0, 128, causal=True)
        got = out[b, h, r0:r0 + 128].float().cpu().numpy()
        ok, nbad, worst_abs, worst_rel = orc.allclose_zig(got, exp, atol=1e-2, rtol=1e-2)   # bf16 output: half an ulp at |o| in [2,4) is 7.8e-3
        assert ok, (b, h, r0, nbad, worst_abs, worst_rel)
        np.testing.assert_allclose(lse[b, h, r0:r0 + 128].cpu().numpy(), exp_lse, rtol=2e-3, atol=2e-3)


# ------------------------------------------------------------------ host-side behaviour
def test_pageable_and_pinned_host_buffers_agree(aule):
    """aule_attention_forward_host: pinned buffers are copied directly, pageable ones through the library's pinned bounce
    buffers (double-buffered per chunk).  Same bits either way, including LSE, several chunks, odd unit counts."""
    import torch
    from aule import ffi
    lib = ffi.ensure_init()
    B, Hq, Hkv, S, D = 3, 6, 3, 300, 64
    torch.manual_seed(1)
    q, k, v = torch.randn(B, Hq, S, D).bfloat16(), torch.randn(B, Hkv, S, D).bfloat16(), torch.randn(B, Hkv, S, D).bfloat16()
    fp = ctypes.POINTER(ctypes.c_float)
    res = []
    for pinned in (False, True):
        t = [x.pin_memory() if pinned else x.clone() for x in (q, k, v)]
        o = torch.empty(B, Hq, S, D, dtype=torch.bfloat16)
        lse = torch.empty(B, Hq, S, dtype=torch.float32)
        if pinned:
            o, lse = o.pin_memory(), lse.pin_memory()
        rc = lib.aule_attention_forward_host(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), o.data_ptr(),
                                             ctypes.cast(lse.data_ptr(), fp), B, Hq, Hkv, S, S, D, ffi.DTYPE_BF16, 0.0, 1, -1, 0)
        assert rc == 0, ffi.last_error()
        res.append((o.clone(), lse.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    exp, _ = orc.attention_ref(q.float().numpy(), k.float().numpy(), v.float().numpy(), causal=True)
    assert orc.rel_err_to_scale(res[0][0].float().numpy(), exp) <= BF16_TOL


def test_concurrent_host_calls_from_two_threads(aule):
    """ctypes releases the GIL; two threads on NumPy inputs share the staging buffers and streams.  The
    host-pointer entries are serialised per device, so every result must be right."""
    shapes = [(1, 4, 150, 64), (2, 2, 260, 32)]
    data = [ref_inputs(*s, seed=7 + i) for i, s in enumerate(shapes)]
    exps = [orc.attention_ref(*d, causal=True)[0] for d in data]
    errs = []

    def work(i):
        try:
            for _ in range(25):
                out = aule.flash_attention(*data[i], causal=True)
                e = orc.rel_err_to_scale(out, exps[i])
                if not (e <= FP32_TOL):
                    errs.append((i, e))
        except Exception as ex:                                           # noqa: BLE001
            errs.append((i, repr(ex)))
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs[:3]


def test_concurrent_device_pointer_calls_on_two_streams(aule):
    """Device-pointer entries may run concurrently: more launches in flight than the scheduler-counter ring has slots,
    from two threads on two streams -- no work item may be skipped or duplicated."""
    import torch
    from aule import cuda_flash
    torch.manual_seed(3)
    q = torch.randn(1, 8, 512, 128, device="cuda").bfloat16()
    k = torch.randn(1, 2, 512, 128, device="cuda").bfloat16()
    v = torch.randn(1, 2, 512, 128, device="cuda").bfloat16()
    ref, _ = cuda_flash.forward_with_lse(q, k, v, causal=True)
    torch.cuda.synchronize()
    bad = []

    def work():
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            outs = [cuda_flash.forward_with_lse(q, k, v, causal=True)[0] for _ in range(400)]
            s.synchronize()
        bad.extend(i for i, o in enumerate(outs) if not torch.equal(o, ref))
    th = [threading.Thread(target=work) for _ in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not bad, bad[:5]


def test_spanning_call_native(aule):
    """aule_attention_forward_spanning_dptr: tensors on GPU 0, computed by every visible GPU (one GPU: degenerate but the
    same code path); bit-identical to the plain call, timings filled."""
    import torch
    from aule import cuda_flash, ffi
    lib = ffi.ensure_init()
    n = min(lib.aule_device_count(), torch.cuda.device_count())
    B, Hq, Hkv, S, D = 2, 8, 4, 640, 128
    torch.manual_seed(0)
    q = torch.randn(B, Hq, S, D, device="cuda:0").bfloat16()
    k = torch.randn(B, Hkv, S, D, device="cuda:0").bfloat16()
    v = torch.randn(B, Hkv, S, D, device="cuda:0").bfloat16()
    ref, ref_lse = cuda_flash.forward_with_lse(q, k, v, causal=True)
    for ndev in sorted({1, n}):
        for chunks in (1, 3):
            o = torch.zeros_like(q); lse = torch.zeros(B, Hq, S, device="cuda:0")
            devs = (ctypes.c_int32 * ndev)(*range(ndev))
            tm = (ctypes.c_float * 4)()
            rc = lib.aule_attention_forward_spanning_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
                                                          B, Hq, Hkv, S, S, D, ffi.DTYPE_BF16, 0.0, 1, -1, 0,
                                                          torch.cuda.current_stream().cuda_stream, devs, ndev, chunks, tm)
            assert rc == 0, ffi.last_error()
            torch.cuda.synchronize()
            assert torch.equal(o, ref) and torch.equal(lse, ref_lse), (ndev, chunks)
            assert tm[3] > 0 and tm[1] > 0
    devs = (ctypes.c_int32 * 1)(7 if n < 8 else 99)
    rc = lib.aule_attention_forward_spanning_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), ref.data_ptr(), 0, B, Hq, Hkv, S, S, D,
                                                  ffi.DTYPE_BF16, 0.0, 1, -1, 0, 0, devs, 1, 1, None)
    assert rc == -4 and "device" in ffi.last_error()


def test_sdpa_shim_only_routes_tensor_core_shapes(aule):
    """shapes that would land on the CUDA-core kernels (fp32, head_dim % 8 != 0) stay with the original SDPA."""
    import torch
    import torch.nn.functional as F
    from aule import ffi
    lib = ffi.load_library()
    q = torch.randn(1, 4, 128, 40, device="cuda", dtype=torch.bfloat16)
    n0 = lib.aule_launch_count()
    o = aule.scaled_dot_product_attention(q, q, q, is_causal=True)       # D = 40: tensor-core kernel (padded to 64)
    assert lib.aule_launch_count() == n0 + 1 and lib.aule_last_kernel().decode() == "aule_fwd_sm100_bf16_d64"
    ref = F.scaled_dot_product_attention(q.float(), q.float(), q.float(), is_causal=True)
    assert (o.float() - ref).abs().max().item() <= 2e-2
    qg = q.clone().requires_grad_()                                        # padded-D training is routed too (tensor-core backward)
    aule.scaled_dot_product_attention(qg, qg, qg, is_causal=True).sum().backward()
    assert lib.aule_last_kernel().decode() == "aule_bwd_dq_sm100_bf16_d64" and torch.isfinite(qg.grad).all()
    n0 = lib.aule_launch_count()
    aule.scaled_dot_product_attention(q.float(), q.float(), q.float(), is_causal=True)     # fp32 -> original SDPA
    q2 = torch.randn(1, 4, 128, 36, device="cuda", dtype=torch.bfloat16)
    aule.scaled_dot_product_attention(q2, q2, q2, is_causal=True)                          # D % 8 != 0 -> original SDPA
    assert lib.aule_launch_count() == n0


# ------------------------------------------------------------------ boundary proof
def test_reference_vulkan_py_drives_libaule(aule, tmp_path):
    """INTEGRATION.md level 1 on hardware: the reference's OWN python/aule package (unmodified, installed by pip under
    baseline/_ref) finds this repository's libaule.so in its package lib/ directory (vulkan.py:31-69), binds every
    prototype (vulkan.py:224-406), and its public API -- aule.vulkan.Aule().attention, attention_gpu with handles, the
    training ABI, and aule.flash_attention on NumPy arrays (which then picks its 'vulkan' backend, __init__.py:188-193)
    -- runs on the B200 through it.  Shapes of python/tests/test_vulkan.py / test_cpu.py; tolerance 1e-4 -> 1e-3 (fp32)."""
    ref_pkg = os.path.join(ROOT, "baseline", "_ref", "aule")
    if not os.path.isdir(ref_pkg):
        pytest.skip("baseline/_ref (pip install of the reference package) not present on this box")
    import shutil
    stage = tmp_path / "refpkg"
    shutil.copytree(ref_pkg, stage / "aule", ignore=shutil.ignore_patterns("__pycache__"))
    (stage / "aule" / "lib").mkdir()
    os.symlink(os.path.join(ROOT, "aule-attention_b200", "python", "aule", "lib", "libaule.so"), stage / "aule" / "lib" / "libaule.so")
    script = r'''
import json, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(1, sys.argv[2])
import aule                                   # the REFERENCE package
import aule.vulkan as vk
from oracle import attention_oracle as orc
assert aule.__file__.startswith(sys.argv[1]), aule.__file__
res = {"backends": aule.get_available_backends()}
ctx = vk.Aule()
res["device"], res["vendor"] = ctx.device_name, ctx.vendor
worst = 0.0
for (B, H, S, D) in [(1, 4, 32, 64), (4, 8, 64, 64), (1, 8, 64, 32), (2, 2, 100, 64)]:
    np.random.seed(42)
    q, k, v = (np.random.randn(B, H, S, D).astype(np.float32) for _ in range(3))
    for causal in (True, False):
        out = ctx.attention(q, k, v, causal=causal)
        exp, _ = orc.attention_ref(q, k, v, causal=causal)
        worst = max(worst, float(np.abs(out - exp).max()))
        out2 = aule.flash_attention(q, k, v, causal=causal)        # public entry, NumPy inputs -> its vulkan backend -> libaule.so
        worst = max(worst, float(np.abs(out2 - exp).max()))
res["fwd_worst_abs"] = worst
np.random.seed(1)
q = np.random.randn(1, 12, 48, 64).astype(np.float32); k = np.random.randn(1, 2, 80, 64).astype(np.float32); v = np.random.randn(1, 2, 80, 64).astype(np.float32)
tq, tk, tv, to = ctx.tensor(q.shape), ctx.tensor(k.shape), ctx.tensor(v.shape), ctx.tensor(q.shape)
tq.upload(q); tk.upload(k); tv.upload(v)
ctx.attention_gpu(tq, tk, tv, to, causal=False)                    # handles: GQA 12/2, Sq != Sk (lib.zig:496-529)
exp, _ = orc.attention_ref(q, k, v, causal=False)
res["gpu_handles_worst_abs"] = float(np.abs(to.download() - exp).max())
if ctx.supports_backward:
    np.random.seed(2)
    q, k, v, do = (np.random.randn(1, 2, 64, 64).astype(np.float32) for _ in range(4))
    o, lse = ctx.attention_forward_with_lse(q, k, v, causal=True)
    dq, dk, dv = ctx.attention_backward(q, k, v, o, do, lse, causal=True)
    eq, ek, ev, eo, el = orc.attention_bwd_ref(q, k, v, do, causal=True)
    res["bwd_worst_abs"] = float(max(np.abs(dq - eq).max(), np.abs(dk - ek).max(), np.abs(dv - ev).max()))
    res["lse_worst_abs"] = float(np.abs(np.asarray(lse).reshape(el.shape) - el).max())
print("RESULT " + json.dumps(res))
'''
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    p = subprocess.run([sys.executable, "-c", script, str(stage), ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    res = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    assert "vulkan" in res["backends"], res                              # the reference believes it is talking to its own library
    assert res["vendor"] == "nvidia" and "B200" in res["device"], res
    assert res["fwd_worst_abs"] <= 1e-3 and res["gpu_handles_worst_abs"] <= 1e-3, res
    assert res.get("bwd_worst_abs", 0.0) <= 1e-3 and res.get("lse_worst_abs", 0.0) <= 1e-3, res
