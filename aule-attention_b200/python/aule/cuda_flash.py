"""
torch CUDA tensors -> libaule.so device-pointer ABI (no copies, caller's stream).

Host-side mirror of /root/reference/python/aule/triton_flash.py:386-558
(`FlashAttentionTritonFunc`, `flash_attention_triton`): same dtype policy
(bf16/fp16 kept, everything else computed in fp32, :405-411; result cast back to
the input dtype, :476), same saved tensors (q, k, v, out, L, :466), same gradient
dtype rule (:526).  PyTorch is only the pointer source here: tensors cross the
boundary as `tensor.data_ptr()` integers plus the current CUstream.
"""
import math

import torch

from . import ffi

_TORCH_TO_AULE = {torch.float32: ffi.DTYPE_F32, torch.bfloat16: ffi.DTYPE_BF16, torch.float16: ffi.DTYPE_F16}


def _check(rc, what):
    if rc != 0:
        raise ffi.AuleError(f"{what}: {ffi.last_error()}")


class FlashAttentionCudaFunc(torch.autograd.Function):
    """Autograd wrapper (mirror of triton_flash.py:386-526)."""

    @staticmethod
    def forward(ctx, q, k, v, causal, scale, window_size, allow_tf32=False):
        lib = ffi.ensure_init()
        B, Hq, Sq, D = q.shape
        _, Hkv, Sk, _ = k.shape
        if scale is None:
            scale = 1.0 / math.sqrt(D)                                   # triton_flash.py:394-395
        orig_dtype = q.dtype
        cdt = orig_dtype if orig_dtype in (torch.bfloat16, torch.float16) else torch.float32   # :405-411
        q, k, v = (t.to(cdt).contiguous() for t in (q, k, v))            # :399-401
        out = torch.empty_like(q)                                        # :434
        lse = torch.empty((B, Hq, Sq), device=q.device, dtype=torch.float32)   # :437
        dev = q.device.index if q.device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(dev).cuda_stream
        code = ffi.DTYPE_F32_TF32 if (allow_tf32 and cdt == torch.float32) else _TORCH_TO_AULE[cdt]
        rc = lib.aule_attention_forward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                             B, Hq, Hkv, Sq, Sk, D, code, float(scale),
                                             1 if causal else 0, int(window_size), dev, stream)
        _check(rc, "Attention failed")
        ctx.save_for_backward(q, k, v, out, lse)                         # :466
        ctx.causal, ctx.scale, ctx.window_size, ctx.orig_dtype, ctx.cdt = causal, float(scale), window_size, orig_dtype, cdt
        return out.to(orig_dtype)                                        # :476

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        lib = ffi.ensure_init()
        B, Hq, Sq, D = q.shape
        _, Hkv, Sk, _ = k.shape
        dout = dout.to(ctx.cdt).contiguous()                             # :492
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dev = q.device.index
        stream = torch.cuda.current_stream(dev).cuda_stream
        window = int(ctx.window_size) if ctx.window_size is not None and ctx.window_size > 0 else -1
        if window > 0:
            # the window is part of the mask in the backward too (the reference's backward forgets it, triton_flash.py:313-319)
            rc = lib.aule_attention_backward_window_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(),
                                                         lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                                         B, Hq, Hkv, Sq, Sk, D, _TORCH_TO_AULE[ctx.cdt], ctx.scale,
                                                         1 if ctx.causal else 0, window, dev, stream)
        else:
            rc = lib.aule_attention_backward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(),
                                                  lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                                  B, Hq, Hkv, Sq, Sk, D, _TORCH_TO_AULE[ctx.cdt], ctx.scale,
                                                  1 if ctx.causal else 0, dev, stream)
        _check(rc, "Backward pass failed")
        od = ctx.orig_dtype
        return dq.to(od), dk.to(od), dv.to(od), None, None, None, None   # :526


_tf32_override = None      # aule.set_fp32_tf32()


def tf32_allowed(allow_tf32=None):
    """fp32 inputs: exact fp32 arithmetic unless tf32 is allowed (argument, aule.set_fp32_tf32, or AULE_TF32=1 in the environment).
    The reference's Triton path always runs fp32 inputs as tf32 (tl.dot default); its Vulkan shaders are exact."""
    import os
    if allow_tf32 is not None:
        return bool(allow_tf32)
    if _tf32_override is not None:
        return _tf32_override
    return os.environ.get("AULE_TF32", "0") == "1"


def flash_attention_cuda(q, k, v, causal=True, scale=None, window_size=-1, allow_tf32=None):
    """Mirror of flash_attention_triton (triton_flash.py:529-558)."""
    return FlashAttentionCudaFunc.apply(q, k, v, causal, scale, window_size, tf32_allowed(allow_tf32))


def forward_with_lse(q, k, v, causal=True, scale=None, window_size=-1, allow_tf32=None):
    """Forward that also returns LSE [B,Hq,Sq] (fp32) -- no autograd."""
    lib = ffi.ensure_init()
    B, Hq, Sq, D = q.shape
    _, Hkv, Sk, _ = k.shape
    cdt = q.dtype if q.dtype in (torch.bfloat16, torch.float16) else torch.float32
    q, k, v = (t.to(cdt).contiguous() for t in (q, k, v))
    out = torch.empty_like(q)
    lse = torch.empty((B, Hq, Sq), device=q.device, dtype=torch.float32)
    dev = q.device.index if q.device.index is not None else torch.cuda.current_device()
    code = ffi.DTYPE_F32_TF32 if (cdt == torch.float32 and tf32_allowed(allow_tf32)) else _TORCH_TO_AULE[cdt]
    rc = lib.aule_attention_forward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                         B, Hq, Hkv, Sq, Sk, D, code,
                                         float(scale) if scale else 0.0, 1 if causal else 0, int(window_size), dev,
                                         torch.cuda.current_stream(dev).cuda_stream)
    _check(rc, "Attention failed")
    return out, lse


def flash_attention_host(q, k, v, causal=True, scale=None, window_size=-1, lse=None, device=0):
    """Host-resident torch tensors (pinned or pageable) through the pipelined
    host-buffer entry (aule_attention_forward_host). Returns a CPU tensor."""
    lib = ffi.ensure_init()
    B, Hq, Sq, D = q.shape
    _, Hkv, Sk, _ = k.shape
    cdt = q.dtype if q.dtype in (torch.bfloat16, torch.float16) else torch.float32
    q, k, v = (t.to(cdt).contiguous() for t in (q, k, v))
    out = torch.empty_like(q, pin_memory=q.is_pinned())
    lse_ptr = None
    if lse is not None:
        import ctypes
        lse_ptr = ctypes.cast(lse.data_ptr(), ctypes.POINTER(ctypes.c_float))
    rc = lib.aule_attention_forward_host(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse_ptr,
                                         B, Hq, Hkv, Sq, Sk, D, _TORCH_TO_AULE[cdt], float(scale) if scale else 0.0,
                                         1 if causal else 0, int(window_size), int(device))
    _check(rc, "Attention failed")
    return out


# =============================================================================
# RoPE (mirror of triton_flash.py:561-703: flash_attention_rope, precompute_rope_frequencies, apply_rope_separate)
# =============================================================================
def _rope_tables(cos, sin, D, S, device):
    """Validate and normalise cos/sin to contiguous fp32 [rows, D/2] on `device` (triton_flash.py:416-417 asserts
    cos.shape[-1] == head_dim // 2).  A [S, D] "full-dim" table or a table shorter than the sequence is an error,
    never a silent reinterpretation / out-of-bounds read."""
    half = D // 2
    if D % 2 != 0:
        raise ValueError(f"RoPE needs an even head_dim, got {D}")
    if cos is None or sin is None:
        raise ValueError("cos and sin are required for RoPE")
    if cos.shape[-1] != half or tuple(sin.shape) != tuple(cos.shape):
        raise ValueError(f"cos/sin must have shape [..., seq_len, head_dim//2 = {half}] and match each other, "
                         f"got cos {tuple(cos.shape)}, sin {tuple(sin.shape)}")
    lead = 1
    for n in cos.shape[:-2]:
        lead *= int(n)
    if lead != 1:
        raise ValueError(f"cos/sin must not carry batch/head dimensions larger than 1, got {tuple(cos.shape)}")
    cos = cos.reshape(-1, half).to(device=device, dtype=torch.float32).contiguous()
    sin = sin.reshape(-1, half).to(device=device, dtype=torch.float32).contiguous()
    if cos.shape[0] < S:
        raise ValueError(f"cos/sin tables have {cos.shape[0]} rows, the sequence needs {S}")
    return cos, sin


class _RopeFunc(torch.autograd.Function):
    """x -> RoPE(x), half-split (triton_flash.py:680-703) or interleaved (attention_f32.comp:98-111) convention.
    The backward pass is the transposed rotation (same kernel, -sin)."""

    @staticmethod
    def forward(ctx, x, cos, sin, interleaved):
        lib = ffi.ensure_init()
        B, H, S, D = x.shape
        cdt = x.dtype if x.dtype in _TORCH_TO_AULE else torch.float32
        xc = x.to(cdt).contiguous()
        cos, sin = _rope_tables(cos, sin, D, S, x.device)
        out = torch.empty_like(xc)
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        rc = lib.aule_rope_dptr(xc.data_ptr(), out.data_ptr(), cos.data_ptr(), sin.data_ptr(), B, H, S, D, cos.shape[0],
                                1 if interleaved else 0, _TORCH_TO_AULE[cdt], 0, dev, torch.cuda.current_stream(dev).cuda_stream)
        _check(rc, "RoPE failed")
        ctx.save_for_backward(cos, sin)
        ctx.dev, ctx.cdt, ctx.orig, ctx.interleaved = dev, cdt, x.dtype, bool(interleaved)
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        cos, sin = ctx.saved_tensors
        lib = ffi.ensure_init()
        B, H, S, D = g.shape
        gc = g.to(ctx.cdt).contiguous()
        out = torch.empty_like(gc)
        rc = lib.aule_rope_dptr(gc.data_ptr(), out.data_ptr(), cos.data_ptr(), sin.data_ptr(), B, H, S, D, cos.shape[0],
                                1 if ctx.interleaved else 0, _TORCH_TO_AULE[ctx.cdt], 1, ctx.dev,
                                torch.cuda.current_stream(ctx.dev).cuda_stream)
        _check(rc, "RoPE backward failed")
        return out.to(ctx.orig), None, None, None


def apply_rope(x, cos, sin, interleaved=False):
    """RoPE on the GPU through the C ABI (differentiable)."""
    return _RopeFunc.apply(x, cos, sin, interleaved)


def flash_attention_rope(q, k, v, cos, sin, causal=True, scale=None, window_size=-1, interleaved=False):
    """Mirror of triton_flash.py:561-603: RoPE on Q and K, then the fused attention.

    Inference (no gradient needed): one C call, aule_attention_forward_rope_dptr -- a single launch reads Q and K once
    and writes their rotated copies once (stream-ordered workspace), then the tcgen05 kernel runs on them.  K has to be
    rotated once per key, not once per (query block, key) pair, so the rotation is a prologue of the kernel's K/V stream,
    not a part of it.  With autograd: apply_rope (differentiable) + the attention Function.
    `interleaved=True` selects the Vulkan shader's pairing (attention_f32.comp:98-111)."""
    assert q.dim() == 4 and k.dim() == 4 and v.dim() == 4
    assert q.shape[-1] == k.shape[-1] == v.shape[-1]
    assert k.shape[1] == v.shape[1] and k.shape[2] == v.shape[2]
    assert q.shape[1] % k.shape[1] == 0
    assert cos is not None and sin is not None, "cos and sin are required for RoPE"
    B, Hq, Sq, D = q.shape
    _, Hkv, Sk, _ = k.shape
    needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (q, k, v))
    if needs_grad:
        return flash_attention_cuda(apply_rope(q, cos, sin, interleaved), apply_rope(k, cos, sin, interleaved), v,
                                    causal=causal, scale=scale, window_size=window_size)
    lib = ffi.ensure_init()
    cos, sin = _rope_tables(cos, sin, D, max(Sq, Sk), q.device)
    orig = q.dtype
    cdt = orig if orig in (torch.bfloat16, torch.float16) else torch.float32
    q, k, v = (t.to(cdt).contiguous() for t in (q, k, v))
    out = torch.empty_like(q)
    dev = q.device.index if q.device.index is not None else torch.cuda.current_device()
    rc = lib.aule_attention_forward_rope_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), 0, cos.data_ptr(),
                                              sin.data_ptr(), cos.shape[0], 1 if interleaved else 0, B, Hq, Hkv, Sq, Sk, D,
                                              _TORCH_TO_AULE[cdt], float(scale) if scale else 0.0, 1 if causal else 0,
                                              int(window_size), dev, torch.cuda.current_stream(dev).cuda_stream)
    _check(rc, "Attention failed")
    return out.to(orig)


def precompute_rope_frequencies(seq_len, head_dim, base=10000.0, device="cuda", dtype=torch.float32):
    """triton_flash.py:640-677: cos, sin of shape [seq_len, head_dim // 2]."""
    half_dim = head_dim // 2
    freqs = 1.0 / (base ** (torch.arange(0, half_dim, device=device, dtype=dtype) / half_dim))
    positions = torch.arange(seq_len, device=device, dtype=dtype)
    angles = positions[:, None] * freqs[None, :]
    return torch.cos(angles), torch.sin(angles)


def apply_rope_separate(q, k, cos, sin):
    """triton_flash.py:680-703: the plain torch formulation (reference for tests)."""
    def rotate_half(x):
        x1 = x[..., :x.shape[-1] // 2]
        x2 = x[..., x.shape[-1] // 2:]
        return torch.cat([-x2, x1], dim=-1)

    seq_len = q.shape[2]
    cos = cos[:seq_len].unsqueeze(0).unsqueeze(0)
    sin = sin[:seq_len].unsqueeze(0).unsqueeze(0)
    cos_full = torch.cat([cos, cos], dim=-1)
    sin_full = torch.cat([sin, sin], dim=-1)
    return q * cos_full + rotate_half(q) * sin_full, k * cos_full + rotate_half(k) * sin_full


def flash_attention_paged(q, k_cache, v_cache, block_tables, context_lens, scale=None, window_size=-1,
                          max_context_len=None):
    """Paged-KV decode: one query token per sequence against a vLLM-style block-table cache.

    Mirror of triton_flash_amd.py:662-740 (`flash_attention_paged_amd`), same arguments:
      q [batch, heads_q, head_dim] (or [batch, heads_q, 1, head_dim], :689-691)
      k_cache, v_cache [num_blocks, block_size, num_kv_heads, head_dim]
      block_tables [batch, max_blocks_per_seq] int32, context_lens [batch] int32
    Returns [batch, heads_q, head_dim].  `max_context_len` (extra, optional): a host-side upper bound of
    context_lens; the reference reads `context_lens.max().item()` back from the device on every call (:711) --
    here nothing is read back, the bound only sizes the split-KV grid (default: max_blocks_per_seq * block_size).
    """
    lib = ffi.ensure_init()
    if q.dim() == 4:
        assert q.shape[2] == 1, "PagedAttention only supports single query token"
        q = q.squeeze(2)
    batch, heads_q, head_dim = q.shape
    num_blocks_total, block_size, heads_kv, _ = k_cache.shape
    assert heads_q % heads_kv == 0, f"heads_q ({heads_q}) must be divisible by heads_kv ({heads_kv})"   # :696-697
    if k_cache.dtype not in (torch.bfloat16, torch.float16):
        raise ffi.AuleError("PagedAttention failed: the KV cache must be bfloat16 or float16")
    cdt = k_cache.dtype
    if scale is None:
        scale = 1.0 / math.sqrt(head_dim)                                # :699-700
    orig_dtype = q.dtype
    q = q.to(cdt).contiguous()                                           # :702-706
    k_cache, v_cache = k_cache.contiguous(), v_cache.to(cdt).contiguous()
    block_tables = block_tables.contiguous().to(torch.int32)
    context_lens = context_lens.contiguous().to(torch.int32)
    out = torch.empty(batch, heads_q, head_dim, device=q.device, dtype=cdt)   # :708
    dev = q.device.index if q.device.index is not None else torch.cuda.current_device()
    rc = lib.aule_attention_paged_decode_dptr(
        q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), block_tables.data_ptr(), context_lens.data_ptr(),
        out.data_ptr(), batch, heads_q, heads_kv, head_dim, num_blocks_total, block_size, block_tables.shape[1],
        int(max_context_len or 0), _TORCH_TO_AULE[cdt], float(scale), int(window_size), dev,
        torch.cuda.current_stream(dev).cuda_stream)
    _check(rc, "PagedAttention failed")
    return out.to(orig_dtype)
