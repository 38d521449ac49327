"""Import-path compatibility: the reference exposes its FFI classes as `aule.vulkan`
(/root/reference/python/aule/vulkan.py).  In this build the same classes bind the
CUDA engine; see ffi.py."""
from .ffi import (Aule, GpuTensor, AuleError, attention, flash_attention, supports_backward,  # noqa: F401
                  attention_forward_with_lse, attention_backward)
