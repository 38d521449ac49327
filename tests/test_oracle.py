"""Pins the oracle (oracle/) against the reference's own golden vectors and
known-answer tests. CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, ref_inputs
from oracle import attention_oracle as orc
from oracle import c_ref

GOLD_CASES = [
    ("cpu_causal_1x4x32x64", (1, 4, 32, 64), True),
    ("cpu_noncausal_1x4x32x64", (1, 4, 32, 64), False),
    ("cpu_batch_4x8x64x64", (4, 8, 64, 64), True),
    ("cpu_d32_1x4x32x32", (1, 4, 32, 32), True),
    ("cpu_d128_1x4x32x128", (1, 4, 32, 128), True),
    ("triton_shape_1x8x64x64_causal", (1, 8, 64, 64), True),
    ("triton_shape_1x8x64x64_noncausal", (1, 8, 64, 64), False),
    ("triton_shape_1x8x64x128_causal", (1, 8, 64, 128), True),
    ("configA_1x8x256x64_causal", (1, 8, 256, 64), True),
]


@pytest.mark.parametrize("name,shape,causal", GOLD_CASES)
def test_numpy_restatement_matches_reference_golden(golden, name, shape, causal):
    """cpu_attention restates __init__.py:247-271; tolerance of the reference's
    own test (python/tests/test_cpu.py:29: rtol=1e-4, atol=1e-4)."""
    q, k, v = ref_inputs(*shape)
    out = orc.cpu_attention(q, k, v, causal=causal)
    assert out.dtype == np.float32
    np.testing.assert_allclose(out, golden[name], rtol=1e-4, atol=1e-4)
    # same calls on the same NumPy build -> expected to be much tighter than that
    assert orc.max_abs_diff(out, golden[name]) < 1e-6


@pytest.mark.parametrize("name,shape,causal", GOLD_CASES)
def test_extended_oracle_matches_reference_golden(golden, name, shape, causal):
    q, k, v = ref_inputs(*shape)
    o64, lse = orc.attention_ref(q, k, v, causal=causal, acc=np.float64)
    np.testing.assert_allclose(o64, golden[name], rtol=1e-4, atol=1e-5)
    assert lse.shape == shape[:3] and np.isfinite(lse).all()


@pytest.mark.parametrize("name,shape,causal", GOLD_CASES[:8])
def test_c_restatement_matches_reference_golden(golden, name, shape, causal):
    """attention_ref.c (== attention_ref.zig loops) vs the reference NumPy path;
    tolerance shape of tests/test_attention.zig:60-77."""
    q, k, v = ref_inputs(*shape)
    out = c_ref.forward(q, k, v, causal)
    assert orc.max_abs_diff(out, golden[name]) < 1e-4 or orc.max_rel_diff(out, golden[name]) < 1e-3
    assert abs(c_ref.max_abs_diff(out, golden[name]) - orc.max_abs_diff(out, golden[name])) < 1e-7


def test_golden_drift_alarm(golden):
    """Survey-time fingerprint of config A on the reference path."""
    o = golden["configA_1x8x256x64_causal"]
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_numpy_path.json")))
    assert abs(float(o.sum(dtype=np.float64)) - (-302.5487)) < 1e-3
    assert abs(float(np.abs(o).max()) - 3.2233982) < 1e-6
    q, k, v = ref_inputs(1, 8, 256, 64)
    assert o[0, 0, 0, 0] == v[0, 0, 0, 0]          # row 0 sees only key 0
    assert meta["cases"]["configA_1x8x256x64_causal"]["sha256_16"] == "11821f11a0960c3f"


# ---- known-answer vectors held by the reference's own tests -------------------
def test_known_answer_attention_ref_zig():
    """attention_ref.zig:250-298: Q=K=0.5, V=[[1,2,3,4],[5,6,7,8]] -> rows [3,4,5,6]."""
    q = np.full((1, 1, 2, 4), 0.5, np.float32)
    k = q.copy()
    v = np.array([[1, 2, 3, 4], [5, 6, 7, 8]], np.float32).reshape(1, 1, 2, 4)
    exp = np.array([3, 4, 5, 6], np.float32)
    for out in (c_ref.forward(q, k, v, False), orc.cpu_attention(q, k, v, causal=False),
                orc.attention_ref(q, k, v, causal=False)[0]):
        assert np.abs(out[0, 0] - exp).max() < 1e-3


def test_known_answer_uniform_weights():
    """tests/test_attention.zig:158-219: Q=K=0.5, S=4, D=8, V[i,d]=i*8+d -> column mean."""
    q = np.full((1, 1, 4, 8), 0.5, np.float32)
    v = (np.arange(4)[:, None] * 8 + np.arange(8)[None, :]).astype(np.float32).reshape(1, 1, 4, 8)
    exp = v[0, 0].mean(axis=0)
    for out in (c_ref.forward(q, q, v, False), orc.attention_ref(q, q, v, causal=False)[0]):
        assert np.abs(out[0, 0] - exp[None, :]).max() < 0.01


def test_known_answer_identity_kv():
    """tests/test_attention.zig:221-270: Q[i,i]=K[i,i]=10, V[n]=0.1n -> |O-V|<0.1."""
    S = D = 8
    q = np.zeros((1, 1, S, D), np.float32)
    q[0, 0, np.arange(S), np.arange(S)] = 10.0
    v = (0.1 * np.arange(S * D, dtype=np.float32)).reshape(1, 1, S, D)
    for out in (c_ref.forward(q, q, v, False), orc.attention_ref(q, q, v, causal=False)[0]):
        assert np.abs(out - v).max() < 0.1


def test_stability_no_nan():
    """tests/test_attention.zig:272-325: inputs U(-5,5), no NaN/Inf."""
    rng = np.random.RandomState(123)
    q, k, v = (rng.uniform(-5, 5, (1, 2, 32, 32)).astype(np.float32) for _ in range(3))
    for causal in (False, True):
        assert np.isfinite(c_ref.forward(q, k, v, causal)).all()
        assert np.isfinite(orc.attention_ref(q, k, v, causal=causal)[0]).all()


def test_batch_independence():
    """tests/test_attention.zig:327-384."""
    rng = np.random.RandomState(456)
    q, k, v = (rng.uniform(-0.5, 0.5, (2, 4, 32, 32)).astype(np.float32) for _ in range(3))
    o1 = c_ref.forward(q[:1], k[:1], v[:1], False)
    o2 = c_ref.forward(q, k, v, False)
    assert np.abs(o1[0] - o2[0]).max() < 1e-5


# ---- extensions: GQA / cross / scale / LSE / rows / backward -------------------
def test_gqa_pin(golden):
    """tests/test_gqa_unit.py:20-55 (MQA 4/1 == repeat_interleave), tol 1e-3."""
    np.random.seed(42)
    q = np.random.randn(1, 4, 16, 64).astype(np.float32)
    k = np.random.randn(1, 1, 16, 64).astype(np.float32)
    v = np.random.randn(1, 1, 16, 64).astype(np.float32)
    o, _ = orc.attention_ref(q, k, v, causal=True)
    np.testing.assert_allclose(o, golden["gqa_mqa_4to1_1x4x16x64"], rtol=1e-3, atol=1e-5)


def test_cross_attention_pin(golden):
    """tests/test_cross_attn.py:12-60 (Sq=16, Sk=32, non-causal)."""
    np.random.seed(42)
    q = np.random.randn(1, 4, 16, 64).astype(np.float32)
    k = np.random.randn(1, 4, 32, 64).astype(np.float32)
    v = np.random.randn(1, 4, 32, 64).astype(np.float32)
    o, _ = orc.attention_ref(q, k, v, causal=False)
    np.testing.assert_allclose(o, golden["cross_1x4_q16_k32_d64"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(orc.cpu_attention(q, k, v, causal=False),
                               golden["cross_1x4_q16_k32_d64"], rtol=1e-6, atol=1e-6)


def test_causal_is_top_left_aligned():
    """row 0 of a Sq=4, Sk=256 causal call equals v[...,0,:]."""
    q, k, v = ref_inputs(1, 2, 4, 32, Sk=256)
    o, _ = orc.attention_ref(q, k, v, causal=True)
    np.testing.assert_allclose(o[:, :, 0], v[:, :, 0], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(orc.cpu_attention(q, k, v, causal=True)[:, :, 0], v[:, :, 0], atol=1e-6)


def test_scale_and_lse():
    q, k, v = ref_inputs(2, 3, 20, 16)
    o, lse = orc.attention_ref(q, k, v, causal=True, scale=0.3)
    s = np.einsum('bhqd,bhkd->bhqk', q.astype(np.float64), k.astype(np.float64)) * 0.3
    s = np.where(np.triu(np.ones((20, 20), bool), 1), -np.inf, s)
    np.testing.assert_allclose(lse, np.log(np.exp(s).sum(-1)), rtol=1e-10)
    p = np.exp(s - lse[..., None])
    np.testing.assert_allclose(o, p @ v.astype(np.float64), rtol=1e-10, atol=1e-12)


def test_rows_matches_full():
    q, k, v = ref_inputs(2, 8, 96, 32, Hkv=2)
    o, lse = orc.attention_ref(q, k, v, causal=True)
    orow, lrow = orc.attention_rows(q, k, v, 1, 5, 40, 17, causal=True)
    np.testing.assert_allclose(orow, o[1, 5, 40:57], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(lrow, lse[1, 5, 40:57], rtol=1e-12)


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("hkv", [4, 2, 1])
def test_backward_against_finite_differences(causal, hkv):
    """dQ/dK/dV formulas (triton_flash.py:321-347) vs central differences in fp64."""
    rng = np.random.RandomState(7)
    B, Hq, S, D = 1, 4, 6, 4
    q = rng.randn(B, Hq, S, D)
    k = rng.randn(B, hkv, S + 2, D)
    v = rng.randn(B, hkv, S + 2, D)
    do = rng.randn(B, Hq, S, D)
    dq, dk, dv, _, _ = orc.attention_bwd_ref(q, k, v, do, causal=causal, scale=0.7)

    def loss(q_, k_, v_):
        return float((orc.attention_ref(q_, k_, v_, causal=causal, scale=0.7)[0] * do).sum())

    eps = 1e-6
    for arr, grad in ((q, dq), (k, dk), (v, dv)):
        for idx in [(0, 0, 0, 0), (0, arr.shape[1] - 1, 3, 2), (0, 0, arr.shape[2] - 1, 1)]:
            a0 = arr[idx]
            arr[idx] = a0 + eps
            lp = loss(q, k, v)
            arr[idx] = a0 - eps
            lm = loss(q, k, v)
            arr[idx] = a0
            assert abs((lp - lm) / (2 * eps) - grad[idx]) < 1e-6


def test_backward_against_torch_autograd():
    """python/tests/test_triton.py:66-94 oracle: torch SDPA autograd (CPU, fp64)."""
    torch = pytest.importorskip("torch")
    import torch.nn.functional as F
    tq, tk, tv, tdo = (torch.randn(2, 4, 16, 8, dtype=torch.float64, generator=torch.Generator().manual_seed(42 + i))
                       for i in range(4))
    tq.requires_grad_(); tk.requires_grad_(); tv.requires_grad_()
    F.scaled_dot_product_attention(tq, tk, tv, is_causal=True).backward(tdo)
    dq, dk, dv, _, _ = orc.attention_bwd_ref(tq.detach().numpy(), tk.detach().numpy(), tv.detach().numpy(),
                                             tdo.numpy(), causal=True)
    np.testing.assert_allclose(dq, tq.grad.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(dk, tk.grad.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(dv, tv.grad.numpy(), rtol=1e-9, atol=1e-11)


def test_bf16_helpers():
    torch = pytest.importorskip("torch")
    x = np.random.RandomState(0).randn(1000).astype(np.float32)
    ref = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    assert np.array_equal(orc.bf16_round(x), ref)
    assert np.array_equal(orc.from_bf16_bits(orc.to_bf16_bits(x)), ref)


def test_rope_ref_matches_reference_formulation():
    """oracle rope_ref vs the reference's apply_rope_separate formula (triton_flash.py:680-703) in torch."""
    torch = pytest.importorskip("torch")
    S, D = 12, 8
    half = D // 2
    freqs = 1.0 / (10000.0 ** (torch.arange(0, half, dtype=torch.float64) / half))
    ang = torch.arange(S, dtype=torch.float64)[:, None] * freqs[None, :]
    cos, sin = torch.cos(ang), torch.sin(ang)
    x = torch.randn(2, 3, S, D, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    x1, x2 = x[..., :half], x[..., half:]
    exp = x * torch.cat([cos, cos], -1)[None, None] + torch.cat([-x2, x1], -1) * torch.cat([sin, sin], -1)[None, None]
    np.testing.assert_allclose(orc.rope_ref(x.numpy(), cos.numpy(), sin.numpy()), exp.numpy(), rtol=1e-12, atol=1e-14)
    # rotation is orthogonal: transpose (negated sin) inverts it
    back = orc.rope_ref(exp.numpy(), cos.numpy(), -sin.numpy())
    np.testing.assert_allclose(back, x.numpy(), rtol=1e-12, atol=1e-13)


# ---------------------------------------------------------------------------------------------------
# Pinning against the reference's own Triton kernels (run unmodified under TRITON_INTERPRET=1 in the build
# container by tests/golden/gen_golden_triton.py): GQA/MQA, scale, LSE, cross attention, window, backward, paged.
def _triton_golden():
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(here, "reference_triton_path.json")) as f:
        meta = json.load(f)
    return np.load(os.path.join(here, "reference_triton_path.npz")), meta


_TG, _TM = _triton_golden()


@pytest.mark.parametrize("name", sorted(_TM["fwd"]))
def test_oracle_matches_reference_triton_kernels(name):
    m = _TM["fwd"][name]
    q, k, v, out = (_TG[f"{name}.{x}"] for x in ("q", "k", "v", "out"))
    # the generic Triton kernel keeps keys with i - j <= W (triton_flash.py:190-193); this build's ABI uses the
    # shader / AMD-kernel convention i - j < W (attention_f32.comp:176-178), so Triton's W is W + 1 here
    w = m["window"] + 1 if m["window"] > 0 else -1
    exp, lse = orc.attention_ref(q, k, v, causal=m["causal"], scale=m["scale"], window=w)
    np.testing.assert_allclose(out, exp, rtol=0, atol=5e-6)
    assert abs(float(out.astype(np.float64).sum()) - m["out_sum"]) < 1e-3
    if m["backward"]:
        dq, dk, dv, _, _ = orc.attention_bwd_ref(q, k, v, _TG[name + ".do"], causal=m["causal"], scale=m["scale"])
        np.testing.assert_allclose(_TG[name + ".lse"], lse, rtol=0, atol=5e-6)     # LSE = m + ln l, triton_flash.py:231-235
        for got, e in ((_TG[name + ".dq"], dq), (_TG[name + ".dk"], dk), (_TG[name + ".dv"], dv)):
            np.testing.assert_allclose(got, e, rtol=0, atol=2e-5)                   # fp32 atomics in the reference


@pytest.mark.parametrize("name", sorted(_TM["paged"]))
def test_paged_oracle_matches_reference_triton_kernel(name):
    m = _TM["paged"][name]
    a = {x: _TG[f"{name}.{x}"] for x in ("q", "k_cache", "v_cache", "block_tables", "context_lens", "out")}
    exp = orc.paged_decode_ref(a["q"], a["k_cache"], a["v_cache"], a["block_tables"], a["context_lens"], window=m["window"])
    np.testing.assert_allclose(a["out"], exp, rtol=0, atol=2e-6)


def test_paged_oracle_equals_dense_attention_on_gathered_cache():
    """Gathering the pages and running the dense oracle's last causal row gives the same answer (GQA, ragged)."""
    rng = np.random.default_rng(3)
    B, Hq, Hkv, D, bs, nb, mb = 2, 6, 2, 32, 16, 20, 5
    q = rng.standard_normal((B, Hq, D)).astype(np.float32)
    kc, vc = (rng.standard_normal((nb, bs, Hkv, D)).astype(np.float32) for _ in range(2))
    bt = rng.permutation(nb)[:B * mb].reshape(B, mb).astype(np.int32)
    lens = np.array([71, 5], dtype=np.int32)
    got = orc.paged_decode_ref(q, kc, vc, bt, lens)
    for b in range(B):
        n = int(lens[b])
        pos = np.arange(n)
        k = kc[bt[b][pos // bs], pos % bs].transpose(1, 0, 2)[None]     # [1,Hkv,n,D]
        v = vc[bt[b][pos // bs], pos % bs].transpose(1, 0, 2)[None]
        exp, _ = orc.attention_ref(q[b][None, :, None, :], k, v, causal=False)
        np.testing.assert_allclose(got[b], exp[0, :, 0, :], rtol=1e-12, atol=1e-12)
    assert np.all(orc.paged_decode_ref(q, kc, vc, bt, np.array([0, 0])) == 0)
