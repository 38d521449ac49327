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
    def forward(ctx, q, k, v, causal, scale, window_size):
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
        rc = lib.aule_attention_forward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                             B, Hq, Hkv, Sq, Sk, D, _TORCH_TO_AULE[cdt], float(scale),
                                             1 if causal else 0, int(window_size), dev, stream)
        _check(rc, "Attention failed")
        ctx.save_for_backward(q, k, v, out, lse)                         # :466
        ctx.causal, ctx.scale, ctx.window_size, ctx.orig_dtype, ctx.cdt = causal, float(scale), window_size, orig_dtype, cdt
        return out.to(orig_dtype)                                        # :476

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        if ctx.window_size is not None and ctx.window_size > 0:
            raise ffi.AuleError("backward with a sliding window is not supported "
                                "(the reference's backward ignores the window, triton_flash.py:313-319)")
        lib = ffi.ensure_init()
        B, Hq, Sq, D = q.shape
        _, Hkv, Sk, _ = k.shape
        dout = dout.to(ctx.cdt).contiguous()                             # :492
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dev = q.device.index
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.aule_attention_backward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(),
                                              lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                              B, Hq, Hkv, Sq, Sk, D, _TORCH_TO_AULE[ctx.cdt], ctx.scale,
                                              1 if ctx.causal else 0, dev, stream)
        _check(rc, "Backward pass failed")
        od = ctx.orig_dtype
        return dq.to(od), dk.to(od), dv.to(od), None, None, None         # :526


def flash_attention_cuda(q, k, v, causal=True, scale=None, window_size=-1):
    """Mirror of flash_attention_triton (triton_flash.py:529-558)."""
    return FlashAttentionCudaFunc.apply(q, k, v, causal, scale, window_size)


def forward_with_lse(q, k, v, causal=True, scale=None, window_size=-1):
    """Forward that also returns LSE [B,Hq,Sq] (fp32) -- no autograd."""
    lib = ffi.ensure_init()
    B, Hq, Sq, D = q.shape
    _, Hkv, Sk, _ = k.shape
    cdt = q.dtype if q.dtype in (torch.bfloat16, torch.float16) else torch.float32
    q, k, v = (t.to(cdt).contiguous() for t in (q, k, v))
    out = torch.empty_like(q)
    lse = torch.empty((B, Hq, Sq), device=q.device, dtype=torch.float32)
    dev = q.device.index if q.device.index is not None else torch.cuda.current_device()
    rc = lib.aule_attention_forward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                         B, Hq, Hkv, Sq, Sk, D, _TORCH_TO_AULE[cdt],
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
