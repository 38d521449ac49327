#include "cuda_driver.h"

#include <dlfcn.h>

namespace aule {

std::string CudaDriver::load() {
    if (handle) return "";
    const char* names[] = {"libcuda.so.1", "libcuda.so"};
    for (const char* n : names) {
        handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (handle) break;
    }
    if (!handle) {
        const char* e = dlerror();
        return std::string("cannot open the CUDA driver (libcuda.so.1): ") + (e ? e : "unknown");
    }
#define AULE_LOAD(name, sym)                                                       \
    name = reinterpret_cast<decltype(name)>(dlsym(handle, #sym));                  \
    if (!name) {                                                                   \
        std::string msg = std::string("CUDA driver lacks symbol ") + #sym;         \
        unload();                                                                  \
        return msg;                                                                \
    }
    AULE_CU_FUNCS(AULE_LOAD)
#undef AULE_LOAD
    return "";
}

void CudaDriver::unload() {
    if (handle) dlclose(handle);
    *this = CudaDriver{};
}

std::string CudaDriver::error_string(CUresult r) const {
    const char* name = nullptr;
    const char* desc = nullptr;
    if (cuGetErrorName) cuGetErrorName(r, &name);
    if (cuGetErrorString) cuGetErrorString(r, &desc);
    return std::string(name ? name : "CUDA_ERROR_?") + " (" + (desc ? desc : "no description") + ")";
}

}  // namespace aule
