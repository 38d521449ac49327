"""
ctypes bindings to libaule.so (the B200 CUDA engine).

Host-side mirror of the reference's FFI module, /root/reference/python/aule/vulkan.py:
the same `Aule` / `GpuTensor` / `AuleError` classes, method names, argument meaning and
error behaviour (vulkan.py:72-161 GpuTensor, :164-1167 Aule), bound to the same C symbols
(include/aule.h part 1), plus the device-pointer extension (part 2) that torch CUDA
tensors use.  There is NO CPU fallback: if the library or a B200 is missing, every
compute call raises AuleError.
"""
import ctypes
import os
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

_LIB = None
_INITIALIZED = False
_INIT_ERROR = None

DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
DTYPE_F32_TF32 = 3      # fp32 tensors, forward allowed to use tf32 tensor-core math (head_dim <= 64)


class AuleError(Exception):
    """Exception raised for aule library errors (vulkan.py:72-74)."""


def _find_library() -> Path:
    """Search order of vulkan.py:31-69 (package lib/ first), plus AULE_LIBRARY_PATH."""
    env = os.environ.get("AULE_LIBRARY_PATH")
    here = Path(__file__).resolve().parent
    candidates = ([Path(env)] if env else []) + [
        here / "lib" / "libaule.so",
        here.parent.parent / "build" / "libaule.so",
        Path("/usr/local/lib/libaule.so"),
        Path("/usr/lib/libaule.so"),
    ]
    for c in candidates:
        if c.exists():
            return c
    raise AuleError("Could not find aule library (libaule.so). Build it with "
                    "`make -C aule-attention_b200` or set AULE_LIBRARY_PATH.")


def _proto(lib):
    c = ctypes
    fp, u32, i32, u64, vp = c.POINTER(c.c_float), c.c_uint32, c.c_int32, c.c_uint64, c.c_void_p
    sig = {
        "aule_init": ([], i32), "aule_shutdown": ([], None), "aule_get_error": ([], c.c_char_p),
        "aule_get_backend_name": ([], c.c_char_p), "aule_supports_backward": ([], i32),
        "aule_get_vendor": ([], i32), "aule_get_gpu_vendor": ([], i32),
        "aule_get_device_name": ([c.c_char_p, u32], i32), "aule_is_amd_optimized": ([], i32),
        "aule_has_fp16": ([], i32), "aule_get_subgroup_size": ([], i32),
        "aule_set_shader_variant": ([c.c_uint8], i32), "aule_get_shader_variant": ([], i32),
        "aule_has_shader_variant": ([c.c_uint8], i32),
        "aule_attention_forward": ([fp, fp, fp, fp, u32, u32, u32, u32, i32], i32),
        "aule_attention_forward_with_lse": ([fp, fp, fp, fp, fp, u32, u32, u32, u32, i32], i32),
        "aule_attention_backward": ([fp] * 9 + [u32] * 4 + [i32], i32),
        "aule_tensor_create": ([u32] * 4, u64), "aule_tensor_create_u32": ([u32] * 4, u64),
        "aule_tensor_destroy": ([u64], None),
        "aule_tensor_upload": ([u64, fp, u32], i32), "aule_tensor_download": ([u64, fp, u32], i32),
        "aule_tensor_download_u32": ([u64, c.POINTER(u32), u32], i32),
        "aule_tensor_size": ([u64], u32), "aule_tensor_count": ([], u32), "aule_tensor_max": ([], u32),
        "aule_tensor_clear_all": ([], None),
        "aule_attention_forward_gpu": ([u64] * 6 + [i32, i32], i32),
        "aule_attention_forward_paged": ([u64] * 6 + [i32, i32], i32),
        "aule_spatial_sort": ([u64, u64, u64, u32], i32),
        "aule_attention_forward_gravity": ([u64] * 7 + [i32, u32, i32], i32),
        "aule_attention_forward_dptr": ([u64] * 5 + [u32] * 6 + [i32, c.c_float, i32, i32, i32, u64], i32),
        "aule_attention_backward_dptr": ([u64] * 9 + [u32] * 6 + [i32, c.c_float, i32, i32, u64], i32),
        "aule_attention_backward_window_dptr": ([u64] * 9 + [u32] * 6 + [i32, c.c_float, i32, i32, i32, u64], i32),
        "aule_attention_forward_host": ([vp, vp, vp, vp, fp] + [u32] * 6 + [i32, c.c_float, i32, i32, i32], i32),
        "aule_attention_backward_host": ([vp] * 5 + [fp] + [vp] * 3 + [u32] * 6 + [i32, c.c_float, i32, i32], i32),
        "aule_rope_dptr": ([u64] * 4 + [u32] * 5 + [i32, i32, i32, i32, u64], i32),
        "aule_attention_forward_rope_dptr": ([u64] * 7 + [u32, i32] + [u32] * 6 + [i32, c.c_float, i32, i32, i32, u64], i32),
        "aule_attention_forward_spanning_dptr": ([u64] * 5 + [u32] * 6 + [i32, c.c_float, i32, i32, i32, u64,
                                                  c.POINTER(i32), i32, i32, fp], i32),
        "aule_attention_paged_decode_dptr": ([u64] * 6 + [u32] * 8 + [i32, c.c_float, i32, i32, u64], i32),
        "aule_device_count": ([], i32), "aule_get_sm_count": ([i32], i32), "aule_synchronize": ([i32], i32),
        "aule_launch_count": ([], u64), "aule_last_kernel": ([], c.c_char_p), "aule_version": ([], c.c_char_p),
        "aule_set_kernel_path": ([i32], i32), "aule_set_trace_buffer": ([u64], i32), "aule_smoke_multiply": ([fp, fp, u32], i32),
    }
    for name, (args, res) in sig.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = res
    return sig


EXPORTED_SYMBOLS = None


def load_library(path: Optional[str] = None):
    """dlopen libaule.so once per process (vulkan.py:205-207) and set prototypes."""
    global _LIB, EXPORTED_SYMBOLS
    if _LIB is None:
        p = Path(path) if path else _find_library()
        lib = ctypes.CDLL(str(p))
        EXPORTED_SYMBOLS = sorted(_proto(lib))
        _LIB = lib
    return _LIB


def ensure_init():
    """aule_init() once globally (vulkan.py:214-220). Raises AuleError when there is no
    usable sm_100 device -- the engine has no CPU fallback by design."""
    global _INITIALIZED, _INIT_ERROR
    lib = load_library()
    if not _INITIALIZED:
        if lib.aule_init() != 0:
            _INIT_ERROR = lib.aule_get_error().decode()
            raise AuleError(f"Failed to initialize aule: {_INIT_ERROR}")
        _INITIALIZED = True
    return lib


def last_error() -> str:
    return load_library().aule_get_error().decode()


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class GpuTensor:
    """Device tensor handle (vulkan.py:77-161). 32-bit elements, true HBM storage."""

    def __init__(self, aule: "Aule", handle: int, shape: Tuple[int, ...], dtype=np.float32):
        self._aule, self._handle, self._shape, self._dtype = aule, handle, tuple(shape), np.dtype(dtype)
        self._size = int(np.prod(shape))

    dtype = property(lambda self: self._dtype)
    shape = property(lambda self: self._shape)
    size = property(lambda self: self._size)
    handle = property(lambda self: self._handle)

    def upload(self, data: np.ndarray) -> None:
        if tuple(data.shape) != self._shape:                          # vulkan.py:119-120
            raise ValueError(f"Shape mismatch: expected {self._shape}, got {data.shape}")
        data = np.ascontiguousarray(data, dtype=np.float32 if self._dtype != np.uint32 else np.uint32)
        rc = self._aule._lib.aule_tensor_upload(ctypes.c_uint64(self._handle), _fptr(data.view(np.float32)), self._size)
        if rc != 0:
            raise AuleError(f"Tensor upload failed: {last_error()}")

    def download(self) -> np.ndarray:
        out = np.empty(self._shape, dtype=np.float32)
        rc = self._aule._lib.aule_tensor_download(ctypes.c_uint64(self._handle), _fptr(out), self._size)
        if rc != 0:
            raise AuleError(f"Tensor download failed: {last_error()}")
        return out.view(self._dtype) if self._dtype == np.uint32 else out

    def destroy(self) -> None:
        if self._handle != 0:
            self._aule._lib.aule_tensor_destroy(ctypes.c_uint64(self._handle))
            self._handle = 0
            try:
                self._aule._tensors.remove(self)
            except ValueError:
                pass


class Aule:
    """Handle-based API of the C ABI (vulkan.py:164-1167)."""

    _VENDOR = {0: "other", 1: "amd", 2: "nvidia", 3: "intel", 4: "apple"}

    def __init__(self, library_path: Optional[str] = None):
        if library_path:
            load_library(library_path)
        self._lib = ensure_init()
        self._tensors = []
        self._initialized = True

    # ---- device info (vulkan.py:408-480)
    @property
    def device_name(self) -> str:
        buf = ctypes.create_string_buffer(256)
        n = self._lib.aule_get_device_name(buf, 256)
        return buf.value.decode() if n > 0 else "Unknown"

    @property
    def vendor(self) -> str:
        return self._VENDOR.get(self._lib.aule_get_vendor(), "unknown")

    is_amd_optimized = property(lambda self: self._lib.aule_is_amd_optimized() == 1)
    fp16_supported = property(lambda self: self._lib.aule_has_fp16() == 1)
    subgroup_size = property(lambda self: self._lib.aule_get_subgroup_size())
    supports_backward = property(lambda self: self._lib.aule_supports_backward() == 1)
    tensor_count = property(lambda self: self._lib.aule_tensor_count())
    tensor_max = property(lambda self: self._lib.aule_tensor_max())
    shader_variant = property(lambda self: self._lib.aule_get_shader_variant())
    shader_variant_name = property(lambda self: "sm100")
    available_shader_variants = property(lambda self: [0])

    def get_device_info(self) -> dict:
        return {"device_name": self.device_name, "vendor": self.vendor, "amd_optimized": False,
                "fp16_supported": True, "subgroup_size": 32, "backend": self._lib.aule_get_backend_name().decode(),
                "sm_count": self._lib.aule_get_sm_count(0), "devices": self._lib.aule_device_count()}

    def set_shader_variant(self, variant: int) -> None:
        rc = self._lib.aule_set_shader_variant(variant)
        if rc != 0:
            raise AuleError(f"Failed to set shader variant {variant}: {last_error()}")

    def has_shader_variant(self, variant: int) -> bool:
        return self._lib.aule_has_shader_variant(variant) == 1

    def clear_tensors(self) -> None:
        """vulkan.py:1120-1130: frees every handle-table slot of the library (all instances)."""
        for t in self._tensors:
            t._handle = 0                                            # the slots are gone: never destroy them twice
        self._tensors = []
        self._lib.aule_tensor_clear_all()

    # ---- tensors (vulkan.py:571-611)
    def tensor(self, shape, dtype=np.float32) -> GpuTensor:
        if len(shape) != 4:
            raise ValueError("Shape must be 4D: [batch, heads, seq, dim]")
        if shape[3] > 128:                                           # vulkan.py:589-590 (64 there)
            raise ValueError(f"head_dim must be <= 128, got {shape[3]}")
        create = self._lib.aule_tensor_create_u32 if np.dtype(dtype) == np.uint32 else self._lib.aule_tensor_create
        h = create(*[int(x) for x in shape])
        if h == 0:
            raise AuleError(f"Failed to create tensor: {last_error()}")
        t = GpuTensor(self, h, tuple(shape), dtype)
        self._tensors.append(t)     # close() / __exit__ release it (the reference forgets this append, vulkan.py:211)
        return t

    def attention_gpu(self, Q, K, V, output, rot_cos=None, rot_sin=None, causal=False, window_size=-1) -> None:
        """vulkan.py:613-659 -> aule_attention_forward_gpu."""
        rc = self._lib.aule_attention_forward_gpu(Q.handle, K.handle, V.handle, output.handle,
                                                  rot_cos.handle if rot_cos else 0, rot_sin.handle if rot_sin else 0,
                                                  1 if causal else 0, int(window_size))
        if rc != 0:
            raise AuleError(f"GPU attention failed: {last_error()}")

    def attention(self, query, key, value, rot_cos=None, rot_sin=None, causal=False, window_size=-1) -> np.ndarray:
        """vulkan.py:661-815: NumPy in, NumPy out (fp32), GQA and Sq != Sk allowed."""
        for name, t in (("query", query), ("key", key), ("value", value)):
            if t.ndim != 4:
                raise ValueError(f"{name} must be 4D [batch, heads, seq_len, head_dim], got shape {t.shape}")
        B, Hq, Sq, D = query.shape
        Bk, Hkv, Sk, Dk = key.shape
        if key.shape != value.shape:
            raise ValueError("Key and value shape mismatch")
        if B != Bk or D != Dk:
            raise ValueError("Batch size / head_dim mismatch between query and key")
        if Hq % Hkv != 0:
            raise ValueError(f"heads_q ({Hq}) must be divisible by heads_kv ({Hkv}) for GQA")
        if D > 128:
            raise ValueError(f"head_dim must be <= 128. Got {D}")
        q = np.ascontiguousarray(query, dtype=np.float32)
        k = np.ascontiguousarray(key, dtype=np.float32)
        v = np.ascontiguousarray(value, dtype=np.float32)
        if rot_cos is not None or rot_sin is not None:
            # vulkan.py:717-790: NumPy cos/sin [.., seq, head_dim/2] -> device tensors -> attention_gpu (interleaved pairs)
            if rot_cos is None or rot_sin is None:
                raise ValueError("rot_cos and rot_sin must be given together")
            half = D // 2
            cos = np.ascontiguousarray(rot_cos, dtype=np.float32).reshape(1, 1, -1, half)
            sin = np.ascontiguousarray(rot_sin, dtype=np.float32).reshape(1, 1, -1, half)
            with Aule() as ctx:
                tq, tk, tv, to = ctx.tensor(q.shape), ctx.tensor(k.shape), ctx.tensor(v.shape), ctx.tensor(q.shape)
                tc, ts = ctx.tensor(cos.shape), ctx.tensor(sin.shape)
                tq.upload(q); tk.upload(k); tv.upload(v); tc.upload(cos); ts.upload(sin)
                ctx.attention_gpu(tq, tk, tv, to, rot_cos=tc, rot_sin=ts, causal=causal, window_size=window_size)
                return to.download()
        out = np.empty_like(q)
        rc = self._lib.aule_attention_forward_host(q.ctypes.data, k.ctypes.data, v.ctypes.data, out.ctypes.data, None,
                                                   B, Hq, Hkv, Sq, Sk, D, DTYPE_F32, 0.0, 1 if causal else 0,
                                                   int(window_size), 0)
        if rc != 0:
            raise AuleError(f"Attention failed: {last_error()}")
        return out

    def attention_forward_with_lse(self, query, key, value, causal=False):
        """vulkan.py:824-889."""
        if query.shape != key.shape or query.shape != value.shape:
            raise ValueError("Q, K, V must have same shape")
        if len(query.shape) != 4:
            raise ValueError("Expected 4D tensors [batch, heads, seq, dim]")
        B, H, S, D = query.shape
        if D > 128:
            raise ValueError(f"head_dim must be <= 128. Got {D}")
        q, k, v = (np.ascontiguousarray(x, dtype=np.float32) for x in (query, key, value))
        out = np.empty_like(q)
        lse = np.empty((B, H, S), dtype=np.float32)
        rc = self._lib.aule_attention_forward_with_lse(_fptr(q), _fptr(k), _fptr(v), _fptr(out), _fptr(lse), B, H, S, D,
                                                       1 if causal else 0)
        if rc != 0:
            raise AuleError(f"Forward with LSE failed: {last_error()}")
        return out, lse

    def attention_backward(self, query, key, value, output, grad_output, lse, causal=False):
        """vulkan.py:891-962."""
        B, H, S, D = query.shape
        arrs = [np.ascontiguousarray(x, dtype=np.float32) for x in (query, key, value, output, grad_output, lse)]
        dq, dk, dv = np.empty_like(arrs[0]), np.empty_like(arrs[1]), np.empty_like(arrs[2])
        rc = self._lib.aule_attention_backward(*[_fptr(a) for a in arrs], _fptr(dq), _fptr(dk), _fptr(dv), B, H, S, D,
                                               1 if causal else 0)
        if rc != 0:
            raise AuleError(f"Backward pass failed: {last_error()}")
        return dq, dk, dv

    def spatial_sort(self, *a, **k):
        raise AuleError("spatial_sort is outside the B200 hot path (unsupported)")

    def attention_gravity(self, *a, **k):
        raise AuleError("attention_gravity is outside the B200 hot path (unsupported)")

    def close(self):
        for t in list(self._tensors):
            t.destroy()
        self._tensors = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


# ---- module-level conveniences (vulkan.py:1170-1290)
_default: Optional[Aule] = None


def _instance() -> Aule:
    global _default
    if _default is None:
        _default = Aule()
    return _default


def attention(query, key, value, causal=False, window_size=-1):
    return _instance().attention(query, key, value, causal=causal, window_size=window_size)


flash_attention = attention


def supports_backward() -> bool:
    try:
        return _instance().supports_backward
    except AuleError:
        return False


def attention_forward_with_lse(query, key, value, causal=False):
    return _instance().attention_forward_with_lse(query, key, value, causal=causal)


def attention_backward(query, key, value, output, grad_output, lse, causal=False):
    return _instance().attention_backward(query, key, value, output, grad_output, lse, causal=causal)
