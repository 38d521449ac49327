// CUDA Driver API bound at run time (dlopen("libcuda.so.1")), the way the reference binds
// Vulkan (src/vulkan_context.zig:325-339 dlopens libvulkan.so.1) and HIP
// (src/backends/hip.zig:126-160).  libaule.so therefore has no link-time dependency on a
// driver: it loads on a CPU-only box (symbol tests) and aule_init() reports the failure.
#pragma once
#include <cuda.h>

#include <string>

namespace aule {

#define AULE_CU_FUNCS(X)                                                                         \
    X(cuInit, cuInit)                                                                            \
    X(cuDriverGetVersion, cuDriverGetVersion)                                                    \
    X(cuGetErrorString, cuGetErrorString)                                                        \
    X(cuGetErrorName, cuGetErrorName)                                                            \
    X(cuDeviceGetCount, cuDeviceGetCount)                                                        \
    X(cuDeviceGet, cuDeviceGet)                                                                  \
    X(cuDeviceGetName, cuDeviceGetName)                                                          \
    X(cuDeviceGetAttribute, cuDeviceGetAttribute)                                                \
    X(cuDevicePrimaryCtxRetain, cuDevicePrimaryCtxRetain)                                        \
    X(cuDevicePrimaryCtxRelease, cuDevicePrimaryCtxRelease_v2)                                   \
    X(cuCtxGetCurrent, cuCtxGetCurrent)                                                          \
    X(cuCtxSetCurrent, cuCtxSetCurrent)                                                          \
    X(cuCtxPushCurrent, cuCtxPushCurrent_v2)                                                     \
    X(cuCtxPopCurrent, cuCtxPopCurrent_v2)                                                       \
    X(cuCtxSynchronize, cuCtxSynchronize)                                                        \
    X(cuModuleLoadData, cuModuleLoadData)                                                        \
    X(cuModuleUnload, cuModuleUnload)                                                            \
    X(cuModuleGetFunction, cuModuleGetFunction)                                                  \
    X(cuFuncSetAttribute, cuFuncSetAttribute)                                                    \
    X(cuLaunchKernel, cuLaunchKernel)                                                            \
    X(cuMemAlloc, cuMemAlloc_v2)                                                                 \
    X(cuMemFree, cuMemFree_v2)                                                                   \
    X(cuMemAllocAsync, cuMemAllocAsync)                                                          \
    X(cuMemFreeAsync, cuMemFreeAsync)                                                            \
    X(cuDeviceGetDefaultMemPool, cuDeviceGetDefaultMemPool)                                      \
    X(cuMemPoolSetAttribute, cuMemPoolSetAttribute)                                              \
    X(cuMemcpyHtoD, cuMemcpyHtoD_v2)                                                             \
    X(cuMemcpyDtoH, cuMemcpyDtoH_v2)                                                             \
    X(cuMemcpyHtoDAsync, cuMemcpyHtoDAsync_v2)                                                   \
    X(cuMemcpyDtoHAsync, cuMemcpyDtoHAsync_v2)                                                   \
    X(cuMemsetD8Async, cuMemsetD8Async)                                                          \
    X(cuMemsetD32Async, cuMemsetD32Async)                                                        \
    X(cuStreamCreate, cuStreamCreate)                                                            \
    X(cuStreamDestroy, cuStreamDestroy_v2)                                                       \
    X(cuStreamSynchronize, cuStreamSynchronize)                                                  \
    X(cuStreamWaitEvent, cuStreamWaitEvent)                                                      \
    X(cuEventCreate, cuEventCreate)                                                              \
    X(cuEventRecord, cuEventRecord)                                                              \
    X(cuEventSynchronize, cuEventSynchronize)                                                    \
    X(cuEventDestroy, cuEventDestroy_v2)                                                         \
    X(cuEventElapsedTime, cuEventElapsedTime)                                                    \
    X(cuEventQuery, cuEventQuery)                                                                \
    X(cuMemHostAlloc, cuMemHostAlloc)                                                            \
    X(cuMemFreeHost, cuMemFreeHost)                                                              \
    X(cuPointerGetAttribute, cuPointerGetAttribute)                                              \
    X(cuMemcpyPeerAsync, cuMemcpyPeerAsync)                                                      \
    X(cuMemcpyDtoDAsync, cuMemcpyDtoDAsync_v2)                                                   \
    X(cuDeviceCanAccessPeer, cuDeviceCanAccessPeer)                                              \
    X(cuCtxEnablePeerAccess, cuCtxEnablePeerAccess)                                              \
    X(cuTensorMapEncodeTiled, cuTensorMapEncodeTiled)

struct CudaDriver {
#define AULE_DECL(name, sym) decltype(&::name) name = nullptr;
    AULE_CU_FUNCS(AULE_DECL)
#undef AULE_DECL
    void* handle = nullptr;
    // Returns "" on success, else a message naming what failed.
    std::string load();
    void unload();
    std::string error_string(CUresult r) const;
};

}  // namespace aule
