"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/aule.h declares, keeps the reference's error conventions, and the Python mirror
validates like the reference and never falls back to a CPU implementation."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, _has_gpu

PKG = os.path.join(ROOT, "aule-attention_b200")
LIB = os.path.join(PKG, "python", "aule", "lib", "libaule.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-s", "-C", PKG])
    return ctypes.CDLL(LIB)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "aule.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"AULE_API\s+[\w\s\*]+?\b(aule_\w+)\s*\(", text)))


def test_header_declares_reference_abi():
    """Every export the reference's vulkan.py binds (vulkan.py:224-406) is declared."""
    syms = declared_symbols()
    needed = ["aule_init", "aule_shutdown", "aule_get_error", "aule_attention_forward", "aule_tensor_create",
              "aule_tensor_destroy", "aule_tensor_upload", "aule_tensor_download", "aule_tensor_download_u32",
              "aule_tensor_size", "aule_attention_forward_gpu", "aule_attention_forward_with_lse",
              "aule_attention_backward", "aule_supports_backward", "aule_get_vendor", "aule_spatial_sort",
              "aule_attention_forward_gravity", "aule_tensor_count", "aule_tensor_max", "aule_tensor_clear_all",
              "aule_get_device_name", "aule_get_gpu_vendor", "aule_is_amd_optimized", "aule_has_fp16",
              "aule_get_subgroup_size", "aule_set_shader_variant", "aule_get_shader_variant",
              "aule_has_shader_variant", "aule_tensor_create_u32", "aule_attention_forward_paged",
              "aule_get_backend_name"]
    assert len(needed) == 31                      # the 31 `export fn aule_*` of src/lib.zig
    for n in needed:
        assert n in syms, n
    for n in ("aule_attention_forward_dptr", "aule_attention_backward_dptr", "aule_attention_forward_host"):
        assert n in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"libaule.so does not export {name}"


def test_python_prototypes_cover_header(lib):
    import aule.ffi as ffi
    ffi.load_library(LIB)
    assert set(declared_symbols()) <= set(ffi.EXPORTED_SYMBOLS)


def test_error_convention_without_init(lib):
    """lib.zig:124-130 'No error'; not-initialised calls return -1 / 0 handle and set the string."""
    if _has_gpu():
        pytest.skip("needs a box without a CUDA driver")
    lib.aule_get_error.restype = ctypes.c_char_p
    lib.aule_tensor_create.restype = ctypes.c_uint64
    lib.aule_tensor_create.argtypes = [ctypes.c_uint32] * 4
    rc = lib.aule_init()
    assert rc == -1
    assert b"CUDA" in lib.aule_get_error()
    assert lib.aule_tensor_create(1, 1, 4, 4) == 0
    assert lib.aule_get_error() == b"Not initialized"
    fp = ctypes.POINTER(ctypes.c_float)
    a = np.zeros(16, np.float32)
    lib.aule_attention_forward.argtypes = [fp] * 4 + [ctypes.c_uint32] * 4 + [ctypes.c_int32]
    p = a.ctypes.data_as(fp)
    assert lib.aule_attention_forward(p, p, p, p, 1, 1, 4, 4, 1) == -1
    assert b"not initialized" in lib.aule_get_error().lower()
    assert lib.aule_get_vendor() == -1
    assert lib.aule_tensor_max() == 1024                   # lib.zig:16
    lib.aule_attention_forward_paged.argtypes = [ctypes.c_uint64] * 6 + [ctypes.c_int32] * 2
    assert lib.aule_attention_forward_paged(0, 0, 0, 0, 0, 0, 0, -1) == -1      # lib.zig:544 (no context)
    lib.aule_spatial_sort.argtypes = [ctypes.c_uint64] * 3 + [ctypes.c_uint32]
    assert lib.aule_spatial_sort(0, 0, 0, 0) == -10                              # out of scope: loud stub
    u64, u32, i32 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int32
    lib.aule_attention_paged_decode_dptr.argtypes = [u64] * 6 + [u32] * 8 + [i32, ctypes.c_float, i32, i32, u64]
    assert lib.aule_attention_paged_decode_dptr(0, 0, 0, 0, 0, 0, 1, 1, 1, 64, 1, 16, 1, 0, 1, 0.0, -1, 0, 0) == -1
    assert b"not initialized" in lib.aule_get_error().lower()


def test_python_validation_matches_reference_messages():
    import aule
    q = np.zeros((1, 4, 8, 16), np.float32)
    with pytest.raises(ValueError, match="query must be 4D"):
        aule.flash_attention(q[0], q, q)
    with pytest.raises(ValueError, match="Batch size mismatch"):
        aule.flash_attention(q, np.zeros((2, 4, 8, 16), np.float32), q)
    with pytest.raises(ValueError, match="head_dim mismatch"):
        aule.flash_attention(q, np.zeros((1, 4, 8, 32), np.float32), q)
    with pytest.raises(ValueError, match="Key/value seq_len mismatch"):
        aule.flash_attention(q, q, np.zeros((1, 4, 9, 16), np.float32))
    with pytest.raises(ValueError, match="must be divisible by heads_kv"):
        aule.flash_attention(q, np.zeros((1, 3, 8, 16), np.float32), np.zeros((1, 3, 8, 16), np.float32))


def test_no_cpu_fallback():
    """north_star: no CPU fallback -- without a device the product path must fail loudly."""
    if _has_gpu():
        pytest.skip("needs a box without a CUDA device")
    import aule
    assert aule.get_available_backends() == []
    assert "cuda" in aule.get_backend_errors()
    q = np.zeros((1, 2, 8, 16), np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        aule.flash_attention(q, q, q)
    with pytest.raises(aule.AuleError):
        aule.Aule()
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):          # paged decode and RoPE entries too
        aule.flash_attention_paged(torch.zeros(1, 2, 64), torch.zeros(2, 16, 2, 64), torch.zeros(2, 16, 2, 64),
                                   torch.zeros(1, 2, dtype=torch.int32), torch.ones(1, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        aule.flash_attention_rope(torch.zeros(1, 2, 8, 16), torch.zeros(1, 2, 8, 16), torch.zeros(1, 2, 8, 16),
                                  torch.ones(8, 8), torch.zeros(8, 8))
    assert aule.flash_attention_paged_amd is aule.flash_attention_paged   # the name the reference exports


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cu", ".cuh", ".zig")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "attention_oracle" not in text, f


def test_public_surface_matches_reference():
    import aule
    for name in ("flash_attention", "attention", "scaled_dot_product_attention", "install", "uninstall",
                 "get_available_backends", "get_backend_errors", "get_backend_info", "print_backend_info",
                 "Aule", "GpuTensor", "AuleError", "__version__"):
        assert hasattr(aule, name), name
    assert aule.attention is aule.flash_attention
    import inspect
    params = list(inspect.signature(aule.flash_attention).parameters)
    assert params == ["query", "key", "value", "rot_cos", "rot_sin", "causal", "scale", "window_size"]   # __init__.py:104
    from aule.vulkan import Aule, GpuTensor, AuleError  # noqa: F401  (reference import path)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/aule.h must compile as C99 (no C++-isms, no torch types) and a C program
    must link against libaule.so and reach aule_get_error()/aule_version() without a GPU."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "aule.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "t.c"
    src.write_text('#include "aule.h"\n#include <stdio.h>\n#include <string.h>\n'
                   'int main(void){ const char* v = aule_version(); const char* e = aule_get_error();\n'
                   '  if (!v || !e) return 2; printf("%s|%s\\n", v, e);\n'
                   '  return aule_attention_paged_decode_dptr(0,0,0,0,0,0,1,1,1,64,1,16,1,0,1,0.0f,-1,0,0) == -1 ? 0 : 3; }\n')
    libdir = os.path.join(ROOT, "aule-attention_b200", "python", "aule", "lib")
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-laule", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "No error" in out.stdout or "|" in out.stdout
