"""GPU parity for (i) the paged-KV decode kernel (reference: python/aule/triton_flash_amd.py:544-740)
and (ii) the main forward/backward path against golden vectors produced by the reference's own Triton kernels
(tests/golden/gen_golden_triton.py, TRITON_INTERPRET=1 in the build container).  Everything goes through the C ABI
(aule_attention_paged_decode_dptr / aule_attention_forward_dptr / aule_attention_backward_dptr).
Tolerances: bf16/fp16 kernels <= 1e-2 of the output scale, fp32 kernels <= 1e-3 (north_star)."""
import json
import os

import numpy as np
import pytest

from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_triton_path.json")) as _f:
    META = json.load(_f)
GOLD = np.load(os.path.join(HERE, "golden", "reference_triton_path.npz"))
TOL16, TOL32 = 1e-2, 1e-3


@pytest.fixture(scope="module")
def aule():
    import aule
    assert aule.get_available_backends() == ["cuda"], aule.get_backend_errors()
    return aule


def _last_kernel():
    from aule import ffi
    return ffi.load_library().aule_last_kernel().decode()


# ------------------------------------------------------------------ paged decode
@pytest.mark.parametrize("name", sorted(META["paged"]))
@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_paged_decode_matches_reference_triton_golden(aule, name, dtype):
    """The reference kernel's own outputs (fp32 inputs); ours computes on the 16-bit rounded cache."""
    import torch
    td = {"bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    m = META["paged"][name]
    a = {x: torch.from_numpy(GOLD[f"{name}.{x}"]).cuda() for x in ("q", "k_cache", "v_cache", "block_tables", "context_lens")}
    out = aule.flash_attention_paged(a["q"].to(td), a["k_cache"].to(td), a["v_cache"].to(td), a["block_tables"],
                                     a["context_lens"], window_size=m["window"])
    assert _last_kernel().startswith("aule_paged_")
    assert out.shape == tuple(GOLD[f"{name}.out"].shape) and out.dtype == td
    tol = 2e-2 if dtype == "bf16" else 4e-3          # inputs rounded to 16 bits vs the fp32 golden
    assert orc.rel_err_to_scale(out.float().cpu().numpy(), GOLD[f"{name}.out"]) <= tol


PAGED_SHAPES = [
    # B, Hq, Hkv, D, block_size, max_blocks, context_lens, window
    (1, 8, 8, 128, 16, 8, [128], -1),                       # exactly full pages
    (3, 32, 8, 128, 16, 40, [640, 1, 333], -1),             # Llama-3-8B decode shape (GQA 4:1), ragged, single token
    (2, 16, 1, 64, 32, 16, [500, 17], -1),                  # MQA with 16 q heads per kv head (all 16 MMA rows live)
    (2, 12, 4, 64, 64, 6, [384, 100], -1),                  # group of 3, 64-token pages
    (4, 8, 2, 128, 128, 20, [2560, 0, 1290, 129], -1),      # empty context -> zeros; split-KV path (nsplit > 1)
    (2, 8, 2, 128, 16, 300, [4800, 4111], -1),              # long context: many splits + combine kernel
    (2, 8, 2, 128, 16, 64, [1000, 37], 100),                # sliding window crossing page and tile boundaries
    (1, 4, 4, 64, 256, 3, [700], 513),                      # 256-token pages, window starts mid-tile
]


@pytest.mark.parametrize("B,Hq,Hkv,D,bs,mb,lens,window", PAGED_SHAPES)
@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_paged_decode_vs_oracle(aule, B, Hq, Hkv, D, bs, mb, lens, window, dtype):
    import torch
    td = {"bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    g = torch.Generator().manual_seed(11)
    nb = B * mb + 3
    q = torch.randn(B, Hq, D, generator=g).to(td)
    kc = torch.randn(nb, bs, Hkv, D, generator=g).to(td)
    vc = torch.randn(nb, bs, Hkv, D, generator=g).to(td)
    bt = torch.randperm(nb, generator=g)[:B * mb].reshape(B, mb).to(torch.int32)
    cl = torch.tensor(lens, dtype=torch.int32)
    out = aule.flash_attention_paged(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), window_size=window,
                                     max_context_len=max(lens))
    torch.cuda.synchronize()
    exp = orc.paged_decode_ref(q.float().numpy(), kc.float().numpy(), vc.float().numpy(), bt.numpy(), cl.numpy(), window=window)
    got = out.float().cpu().numpy()
    assert np.isfinite(got).all()
    assert orc.rel_err_to_scale(got, exp) <= TOL16
    for b, n in enumerate(lens):
        if n == 0:
            assert not got[b].any()
    # 4-D query [B, Hq, 1, D] is accepted and squeezed (triton_flash_amd.py:689-691)
    out4 = aule.flash_attention_paged(q.cuda().unsqueeze(2), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), window_size=window)
    assert out4.shape == (B, Hq, D)
    assert orc.rel_err_to_scale(out4.float().cpu().numpy(), exp) <= TOL16


def test_paged_decode_ignores_garbage_outside_the_context(aule):
    """Slots past context_len (and pages the table does not reference) may hold anything, NaN/Inf included."""
    import torch
    B, Hq, Hkv, D, bs, mb = 2, 8, 2, 128, 16, 8
    g = torch.Generator().manual_seed(5)
    nb = B * mb
    q = torch.randn(B, Hq, D, generator=g).bfloat16().cuda()
    kc = torch.randn(nb, bs, Hkv, D, generator=g).bfloat16().cuda()
    vc = torch.randn(nb, bs, Hkv, D, generator=g).bfloat16().cuda()
    bt = torch.arange(nb, dtype=torch.int32).reshape(B, mb).cuda()
    cl = torch.tensor([37, 100], dtype=torch.int32).cuda()
    ref = aule.flash_attention_paged(q, kc, vc, bt, cl)
    kc2, vc2 = kc.clone().view(B, mb * bs, Hkv, D), vc.clone().view(B, mb * bs, Hkv, D)
    for b, n in enumerate((37, 100)):
        kc2[b, n:] = float("nan")
        vc2[b, n:] = float("inf")
    got = aule.flash_attention_paged(q, kc2.view(nb, bs, Hkv, D), vc2.view(nb, bs, Hkv, D), bt, cl)
    assert torch.equal(ref, got)


def test_paged_decode_equals_dense_forward_last_row(aule):
    """Size-independent property: decode over a paged copy of K/V == the last causal row of the dense kernel."""
    import torch
    from aule import cuda_flash
    B, Hq, Hkv, S, D, bs = 2, 32, 8, 2048, 128, 16
    g = torch.Generator(device="cuda").manual_seed(9)
    q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
    dense, _ = cuda_flash.forward_with_lse(q, k, v, causal=True)
    nb = B * S // bs
    perm = torch.randperm(nb, device="cuda", generator=g)
    kc = torch.empty(nb, bs, Hkv, D, device="cuda", dtype=torch.bfloat16)
    vc = torch.empty_like(kc)
    kc[perm] = k.permute(0, 2, 1, 3).reshape(nb, bs, Hkv, D)           # page i of the logical order lives at perm[i]
    vc[perm] = v.permute(0, 2, 1, 3).reshape(nb, bs, Hkv, D)
    bt = perm.reshape(B, S // bs).to(torch.int32)
    cl = torch.full((B,), S, dtype=torch.int32, device="cuda")
    dec = aule.flash_attention_paged(q[:, :, -1, :], kc, vc, bt, cl)
    err = (dec.float() - dense[:, :, -1, :].float()).abs().max().item() / dense[:, :, -1, :].float().abs().max().item()
    assert err <= TOL16


def test_paged_decode_errors_are_loud(aule):
    import torch
    from aule import ffi
    q = torch.randn(1, 4, 64, device="cuda", dtype=torch.bfloat16)
    bt = torch.zeros(1, 2, dtype=torch.int32, device="cuda")
    cl = torch.ones(1, dtype=torch.int32, device="cuda")
    with pytest.raises(ffi.AuleError, match="multiple of 16"):
        c = torch.randn(4, 8, 4, 64, device="cuda", dtype=torch.bfloat16)
        aule.flash_attention_paged(q, c, c, bt, cl)
    with pytest.raises(ffi.AuleError, match="bfloat16 or float16"):
        c = torch.randn(4, 16, 4, 64, device="cuda")
        aule.flash_attention_paged(q, c, c, bt, cl)
    with pytest.raises(AssertionError, match="divisible"):
        c = torch.randn(4, 16, 3, 64, device="cuda", dtype=torch.bfloat16)
        aule.flash_attention_paged(q, c, c, bt, cl)
    with pytest.raises(ffi.AuleError, match="head_dim 64 or 128"):
        c = torch.randn(4, 16, 4, 32, device="cuda", dtype=torch.bfloat16)
        aule.flash_attention_paged(torch.randn(1, 4, 32, device="cuda", dtype=torch.bfloat16), c, c, bt, cl)


# ------------------------------------------------------------------ main path vs the reference's Triton kernels
@pytest.mark.parametrize("name", sorted(META["fwd"]))
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_forward_backward_match_reference_triton_golden(aule, name, dtype):
    """aule.flash_attention on the golden inputs vs what the reference's Triton kernels returned for them
    (forward, LSE and dQ/dK/dV).  f32 runs the CUDA-core kernels, bf16 the tcgen05 kernels (D in {64,128})."""
    import torch
    from aule import cuda_flash
    m = META["fwd"][name]
    td = {"f32": torch.float32, "bf16": torch.bfloat16}[dtype]
    tol = TOL32 if dtype == "f32" else 2e-2         # bf16: inputs are rounded to bf16 first, golden is fp32
    q, k, v = (torch.from_numpy(GOLD[f"{name}.{x}"]).cuda().to(td).requires_grad_(m["backward"]) for x in ("q", "k", "v"))
    w = m["window"] + 1 if m["window"] > 0 else -1   # Triton keeps i-j <= W, this ABI i-j < W (see tests/test_oracle.py)
    if w > 0 and not m["backward"]:
        out, lse = cuda_flash.forward_with_lse(q, k, v, causal=m["causal"], scale=m["scale"], window_size=w)
    else:
        out = aule.flash_attention(q, k, v, causal=m["causal"], scale=m["scale"])
    assert out.dtype == td and out.shape == q.shape
    assert orc.rel_err_to_scale(out.detach().float().cpu().numpy(), GOLD[f"{name}.out"]) <= tol
    if m["backward"]:
        out.backward(torch.from_numpy(GOLD[f"{name}.do"]).cuda().to(td))
        for t, key in ((q, "dq"), (k, "dk"), (v, "dv")):
            assert t.grad.dtype == td                                         # triton_flash.py:526
            assert orc.rel_err_to_scale(t.grad.float().cpu().numpy(), GOLD[f"{name}.{key}"]) <= tol
        _, lse = cuda_flash.forward_with_lse(q.detach(), k.detach(), v.detach(), causal=m["causal"], scale=m["scale"])
        np.testing.assert_allclose(lse.cpu().numpy(), GOLD[f"{name}.lse"], rtol=0, atol=2e-2 if dtype == "bf16" else 1e-4)
