"""
aule (B200 build): FlashAttention-2 on NVIDIA B200 behind the aule-attention API.

Same public surface as the reference package (/root/reference/python/aule/__init__.py:564-592):

    from aule import flash_attention
    out = flash_attention(q, k, v, causal=True)          # [B, H, S, D] tensors

but ONE backend: hand-written sm_100a CUDA kernels in libaule.so, driven through a C ABI
(include/aule.h).  torch CUDA tensors cross the boundary as raw device pointers on the
current stream; NumPy arrays / host tensors go through the pipelined host-buffer entry.
There is no Triton, no Vulkan, no multi-backend dispatch and no CPU fallback: without a
B200 every compute call raises.
"""
import logging
import warnings

__version__ = "0.5.0"            # API level of the reference this build is a drop-in for (__init__.py:26)
__author__ = "Aule Technologies (B200 engine: aule-attention_b200)"

logger = logging.getLogger(__name__)

_backend_errors = {}
_cuda_available = False

from .ffi import Aule, GpuTensor, AuleError  # noqa: E402  (same names the reference exports, __init__.py:91)
from . import ffi as _ffi  # noqa: E402

try:
    _ffi.ensure_init()
    _cuda_available = True
    logger.debug("CUDA sm_100 backend loaded successfully")
except Exception as e:  # library missing or no B200: recorded, reported, never papered over
    _backend_errors['cuda'] = str(e)
    logger.debug(f"CUDA sm_100 backend failed to load: {e}")

_original_sdpa = None
_installed = False
_forced_backend = None    # kept for signature parity; only None / 'cuda' are accepted
_verbose = False


def _validate(query, key, value):
    """Shape checks of the reference entry point (__init__.py:140-160), same messages."""
    if query.ndim != 4:
        raise ValueError(f"query must be 4D [batch, heads, seq_len, head_dim], got shape {query.shape}")
    if key.ndim != 4:
        raise ValueError(f"key must be 4D [batch, heads, seq_len, head_dim], got shape {key.shape}")
    if value.ndim != 4:
        raise ValueError(f"value must be 4D [batch, heads, seq_len, head_dim], got shape {value.shape}")
    batch_q, heads_q, seq_q, head_dim_q = query.shape
    batch_k, heads_kv, seq_k, head_dim_k = key.shape
    batch_v, heads_v, seq_v, head_dim_v = value.shape
    if batch_q != batch_k or batch_q != batch_v:
        raise ValueError(f"Batch size mismatch: query={batch_q}, key={batch_k}, value={batch_v}")
    if head_dim_q != head_dim_k or head_dim_q != head_dim_v:
        raise ValueError(f"head_dim mismatch: query={head_dim_q}, key={head_dim_k}, value={head_dim_v}")
    if seq_k != seq_v:
        raise ValueError(f"Key/value seq_len mismatch: key={seq_k}, value={seq_v}")
    if heads_kv != heads_v:
        raise ValueError(f"Key/value heads mismatch: key={heads_kv}, value={heads_v}")
    if heads_q % heads_kv != 0:
        raise ValueError(f"heads_q ({heads_q}) must be divisible by heads_kv ({heads_kv}) for GQA")
    if head_dim_q > 128 or head_dim_q % 4 != 0:
        raise ValueError(f"head_dim must be a multiple of 4 and <= 128, got {head_dim_q}")


def flash_attention(query, key, value, rot_cos=None, rot_sin=None, causal=True, scale=None, window_size=-1):
    """
    FlashAttention-2 forward (differentiable for torch CUDA tensors).

    Signature and argument meaning of the reference (__init__.py:104-129).

    Args:
        query: [batch, heads_q, seq_len_q, head_dim] - torch.Tensor or numpy.ndarray
        key:   [batch, heads_kv, seq_len_k, head_dim]
        value: [batch, heads_kv, seq_len_k, head_dim]
        rot_cos / rot_sin: optional RoPE tables [.., seq_len, head_dim/2]; applied to Q and K before the attention with
            the pairing of the backend that consumes them in the reference -- the Vulkan shader's adjacent pairs
            (2i, 2i+1) (attention_f32.comp:98-111; the reference's Triton path drops them, __init__.py:204-207).
            Use flash_attention_rope for the Triton half-split convention.
        causal: top-left aligned causal mask (default True)
        scale: softmax scale (default 1/sqrt(head_dim))
        window_size: sliding window (-1 = full attention).  causal: key j visible iff 0 <= i-j < W; bidirectional:
            |i-j| <= W//2 (the shader's convention, attention_f32.comp:173-183).  The reference's generic Triton
            kernel keeps i-j <= W / |i-j| <= W (triton_flash.py:190-194): its W is W+1 (causal) / 2W (bidirectional) here.

    fp32 CUDA tensors run exact fp32 arithmetic (CUDA cores) unless tf32 is allowed -- aule.set_fp32_tf32(True) or AULE_TF32=1 --
    in which case head_dim <= 64 runs the forward on the tensor cores as tf32 (~1e-3 relative error), which is what the
    reference's Triton path always does with fp32 inputs (tl.dot, triton_flash.py:405-411).

    Returns: tensor with the shape (and dtype/device kind) of `query`.
    Raises: ValueError for invalid shapes; RuntimeError if no B200 backend is available.
    """
    _validate(query, key, value)
    if not _cuda_available:
        raise RuntimeError("aule (B200 build): CUDA sm_100 backend not available and there is no CPU fallback: "
                           + _backend_errors.get('cuda', 'unknown error'))
    if (rot_cos is None) != (rot_sin is None):
        raise ValueError("rot_cos and rot_sin must be given together")
    if _verbose:
        print(f"aule-attention: cuda-sm100 | shape={tuple(query.shape)} | causal={causal}")

    is_torch = False
    try:
        import torch
        is_torch = isinstance(query, torch.Tensor)
    except ImportError:
        pass

    if is_torch:
        from . import cuda_flash
        if rot_cos is not None:
            if not query.is_cuda:
                query, key, value = (t.cuda() for t in (query, key, value))
                return cuda_flash.flash_attention_rope(query, key, value, rot_cos, rot_sin, causal=causal, scale=scale,
                                                       window_size=window_size, interleaved=True).cpu()
            return cuda_flash.flash_attention_rope(query, key, value, rot_cos, rot_sin, causal=causal, scale=scale,
                                                   window_size=window_size, interleaved=True)
        if query.is_cuda:
            return cuda_flash.flash_attention_cuda(query, key, value, causal=causal, scale=scale, window_size=window_size)
        # host tensors: staged through HBM by the library (still GPU compute, not a CPU fallback)
        out = cuda_flash.flash_attention_host(query, key, value, causal=causal, scale=scale, window_size=window_size)
        return out.to(query.dtype)

    import numpy as np
    if rot_cos is not None:
        if scale is not None:
            raise ValueError("scale is not supported together with rot_cos/rot_sin on NumPy inputs (handle ABI, lib.zig:496-529)")
        return Aule().attention(query, key, value, rot_cos=rot_cos, rot_sin=rot_sin, causal=causal,
                                window_size=window_size).astype(query.dtype, copy=False)
    q = np.ascontiguousarray(query, dtype=np.float32)
    k = np.ascontiguousarray(key, dtype=np.float32)
    v = np.ascontiguousarray(value, dtype=np.float32)
    out = np.empty_like(q)
    lib = _ffi.ensure_init()
    B, Hq, Sq, D = q.shape
    _, Hkv, Sk, _ = k.shape
    rc = lib.aule_attention_forward_host(q.ctypes.data, k.ctypes.data, v.ctypes.data, out.ctypes.data, None,
                                         B, Hq, Hkv, Sq, Sk, D, _ffi.DTYPE_F32, float(scale) if scale else 0.0,
                                         1 if causal else 0, int(window_size), 0)
    if rc != 0:
        raise AuleError(f"Attention failed: {_ffi.last_error()}")
    return out.astype(query.dtype, copy=False)


attention = flash_attention      # alias, __init__.py:275


def set_fp32_tf32(allow):
    """(extension) True: fp32 CUDA tensors may run the forward as tf32 on the tensor cores (head_dim <= 64); False: always exact
    fp32; None: follow the AULE_TF32 environment variable (the default)."""
    from . import cuda_flash
    cuda_flash._tf32_override = None if allow is None else bool(allow)


def flash_attention_rope(q, k, v, cos, sin, causal=True, scale=None, window_size=-1, interleaved=False):
    """RoPE on Q and K, then attention -- reference: triton_flash.py:561-603 (half-split convention, the default);
    interleaved=True selects the Vulkan shader's adjacent-pair convention (attention_f32.comp:98-111)."""
    from .cuda_flash import flash_attention_rope as _impl
    _validate(q, k, v)
    if not _cuda_available:
        raise RuntimeError("aule (B200 build): CUDA sm_100 backend not available and there is no CPU fallback: "
                           + _backend_errors.get('cuda', 'unknown error'))
    return _impl(q, k, v, cos, sin, causal=causal, scale=scale, window_size=window_size, interleaved=interleaved)


def flash_attention_paged(q, k_cache, v_cache, block_tables, context_lens, scale=None, window_size=-1,
                          max_context_len=None):
    """PagedAttention decode (one query token per sequence, vLLM-style block tables) -- reference:
    flash_attention_paged_amd, triton_flash_amd.py:662-740 (exported at __init__.py:59,575)."""
    from .cuda_flash import flash_attention_paged as _impl
    if not _cuda_available:
        raise RuntimeError("aule (B200 build): CUDA sm_100 backend not available and there is no CPU fallback: "
                           + _backend_errors.get('cuda', 'unknown error'))
    return _impl(q, k_cache, v_cache, block_tables, context_lens, scale=scale, window_size=window_size,
                 max_context_len=max_context_len)


flash_attention_paged_amd = flash_attention_paged      # the name the reference exports (__init__.py:575)


def precompute_rope_frequencies(seq_len, head_dim, base=10000.0, device="cuda", dtype=None):
    from .cuda_flash import precompute_rope_frequencies as _impl
    import torch
    return _impl(seq_len, head_dim, base=base, device=device, dtype=dtype or torch.float32)


def apply_rope_separate(q, k, cos, sin):
    from .cuda_flash import apply_rope_separate as _impl
    return _impl(q, k, cos, sin)


# =============================================================================
# PyTorch SDPA compatibility layer (reference __init__.py:288-442)
# =============================================================================
def scaled_dot_product_attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None,
                                 enable_gqa=False):
    """Drop-in for torch.nn.functional.scaled_dot_product_attention. Falls back to the original
    SDPA for features outside the kernel (mask / dropout), same rule as __init__.py:321-347."""
    import torch
    # Only shapes the tensor-core kernel takes are routed here (16-bit, head_dim <= 128 and a multiple of 8): anything
    # that would land on the CUDA-core kernels (fp32, other head_dims) is SLOWER than the SDPA it would replace.
    # fp32: only when torch itself is allowed to use tf32 for fp32 matmuls, for shapes the tf32 kernel takes, and not for training
    # (the fp32 backward is the exact CUDA-core one)
    tf32_ok = (query.dtype == torch.float32 and torch.backends.cuda.matmul.allow_tf32 and query.dim() == 4
               and query.shape[-1] <= 64 and query.shape[-1] % 4 == 0
               and not (torch.is_grad_enabled() and (query.requires_grad or key.requires_grad or value.requires_grad)))
    needs_fallback = (attn_mask is not None or dropout_p > 0.0 or not _cuda_available or not query.is_cuda
                      or query.dim() != 4 or query.shape[-1] > 128 or (query.shape[-1] % 8 != 0 and not tf32_ok)
                      or (query.dtype not in (torch.bfloat16, torch.float16) and not tf32_ok)
                      or key.dtype != query.dtype or value.dtype != query.dtype
                      or (is_causal and query.shape[-2] != key.shape[-2])      # SDPA's causal is top-left too, but be strict
                      or (query.shape[1] != key.shape[1] and not enable_gqa))
    if needs_fallback:
        fn = _original_sdpa if _original_sdpa is not None else torch.nn.functional.scaled_dot_product_attention
        if fn is scaled_dot_product_attention:
            raise RuntimeError("no original SDPA available for fallback")
        return fn(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
                  enable_gqa=enable_gqa)
    if tf32_ok:
        from . import cuda_flash
        return cuda_flash.flash_attention_cuda(query, key, value, causal=is_causal, scale=scale, allow_tf32=True)
    return flash_attention(query, key, value, causal=is_causal, scale=scale)


def install(backend=None, verbose=False):
    """Install as torch's SDPA (reference __init__.py:353-421). `backend` accepts None or 'cuda'."""
    global _original_sdpa, _installed, _forced_backend, _verbose
    import torch
    import torch.nn.functional as F
    if backend is not None and backend != 'cuda':
        raise ValueError(f"Invalid backend '{backend}'. This build has a single backend: 'cuda' (or None)")
    if _installed:
        _forced_backend, _verbose = backend, verbose
        print(f"aule-attention: Updated (backend={backend or 'auto'}, verbose={verbose})")
        return
    _original_sdpa = F.scaled_dot_product_attention
    F.scaled_dot_product_attention = scaled_dot_product_attention
    torch.nn.functional.scaled_dot_product_attention = scaled_dot_product_attention
    _installed, _forced_backend, _verbose = True, backend, verbose
    print(f"aule-attention: Installed (cuda-sm100{', verbose' if verbose else ''})")


def uninstall():
    """Restore the original SDPA (reference __init__.py:424-442)."""
    global _installed
    if not _installed:
        print("aule-attention: Not installed")
        return
    import torch
    import torch.nn.functional as F
    if _original_sdpa is not None:
        F.scaled_dot_product_attention = _original_sdpa
        torch.nn.functional.scaled_dot_product_attention = _original_sdpa
    _installed = False
    print("aule-attention: Uninstalled, restored PyTorch SDPA")


# =============================================================================
# Backend introspection (reference __init__.py:445-561)
# =============================================================================
def get_available_backends():
    return ['cuda'] if _cuda_available else []


def get_backend_errors():
    return dict(_backend_errors)


def get_backend_info():
    info = {}
    if _cuda_available:
        try:
            dev = Aule().get_device_info()
            info['cuda'] = {'available': True, 'device': dev.get('device_name', 'Unknown'),
                            'sm_count': dev.get('sm_count'), 'devices': dev.get('devices'),
                            'description': 'sm_100a tcgen05/TMA FlashAttention-2 (libaule.so)'}
        except Exception:
            info['cuda'] = {'available': True, 'device': 'Unknown'}
    else:
        info['cuda'] = {'available': False, 'error': _backend_errors.get('cuda')}
    return info


def print_backend_info():
    print("=" * 60)
    print("AULE-ATTENTION (B200 build) v" + __version__)
    print("=" * 60)
    print(f"Available backends: {get_available_backends()}")
    for name, d in get_backend_info().items():
        print(f"[{name.upper()}] {d}")
    print("=" * 60)


__all__ = [
    "flash_attention", "attention", "scaled_dot_product_attention",
    "flash_attention_rope", "precompute_rope_frequencies", "apply_rope_separate",
    "flash_attention_paged", "flash_attention_paged_amd",
    "install", "uninstall",
    "get_available_backends", "get_backend_errors", "get_backend_info", "print_backend_info",
    "Aule", "GpuTensor", "AuleError",
    "__version__",
]
