#include "engine.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../kernels/kernel_params.h"

// sm_100a cubin embedded by embed_cubin.S (.incbin), the analogue of lib.zig:29-50 @embedFile.
extern "C" const unsigned char aule_cubin_start[];
extern "C" const unsigned char aule_cubin_end[];

namespace aule {

using aule_kp::BwdCfg;
using aule_kp::BwdParams;
using aule_kp::FwdCfg;
using aule_kp::FwdParams;
using aule_kp::SimtParams;

static const char* kDtypeSuffix[3] = {"f32", "bf16", "f16"};

CtxGuard::CtxGuard(CudaDriver& drv, CUcontext ctx) : drv_(drv) {
    CUcontext cur = nullptr;
    drv_.cuCtxGetCurrent(&cur);
    if (cur != ctx) {
        drv_.cuCtxPushCurrent(ctx);
        pushed_ = true;
    }
}
CtxGuard::~CtxGuard() {
    if (pushed_) {
        CUcontext old;
        drv_.cuCtxPopCurrent(&old);
    }
}

bool Engine::log_enabled() {
    static const bool on = [] { const char* v = getenv("AULE_LOG"); return v && *v && *v != '0'; }();
    return on;
}

std::string Engine::check(CUresult r, const char* what) const {
    if (r == CUDA_SUCCESS) return "";
    return std::string(what) + ": " + drv_.error_string(r);
}

std::string Engine::init() {
    if (ready_) return "";
    std::string e = drv_.load();
    if (!e.empty()) return e;
    if (!(e = check(drv_.cuInit(0), "cuInit")).empty()) { drv_.unload(); return e; }
    int n = 0;
    if (!(e = check(drv_.cuDeviceGetCount(&n), "cuDeviceGetCount")).empty()) { drv_.unload(); return e; }
    if (n <= 0) { drv_.unload(); return "no CUDA device visible"; }
    std::string why;
    for (int i = 0; i < n; ++i) {
        std::string de = load_device(i);
        if (!de.empty()) why += (why.empty() ? "" : "; ") + de;
    }
    if (devices_.empty()) {
        drv_.unload();
        return "no usable sm_100 device: " + why;
    }
    // direct NVLink access between every pair of devices (spanning calls: cuMemcpyPeerAsync without a host bounce)
    for (Device& a : devices_)
        for (Device& b : devices_) {
            if (a.index == b.index) continue;
            int can = 0;
            if (drv_.cuDeviceCanAccessPeer(&can, a.dev, b.dev) != CUDA_SUCCESS || !can) continue;
            CtxGuard g(drv_, a.ctx);
            drv_.cuCtxEnablePeerAccess(b.ctx, 0);       // CUDA_ERROR_PEER_ACCESS_ALREADY_ENABLED is fine (torch may have done it)
        }
    ready_ = true;
    return "";
}

std::string Engine::load_device(int ordinal) {
    Device d;
    d.ordinal = ordinal;
    std::string e;
    if (!(e = check(drv_.cuDeviceGet(&d.dev, ordinal), "cuDeviceGet")).empty()) return e;
    drv_.cuDeviceGetName(d.name, sizeof(d.name), d.dev);
    drv_.cuDeviceGetAttribute(&d.cc_major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, d.dev);
    drv_.cuDeviceGetAttribute(&d.cc_minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, d.dev);
    drv_.cuDeviceGetAttribute(&d.sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, d.dev);
    if (d.cc_major != 10 || d.cc_minor != 0) {
        char buf[320];
        snprintf(buf, sizeof(buf), "device %d (%s) is sm_%d%d; this library carries sm_100a code only", ordinal, d.name,
                 d.cc_major, d.cc_minor);
        return buf;
    }
    if (!(e = check(drv_.cuDevicePrimaryCtxRetain(&d.ctx, d.dev), "cuDevicePrimaryCtxRetain")).empty()) return e;
    CtxGuard g(drv_, d.ctx);
    if (!(e = check(drv_.cuModuleLoadData(&d.mod, aule_cubin_start), "cuModuleLoadData(sm_100a cubin)")).empty()) {
        drv_.cuDevicePrimaryCtxRelease(d.dev);
        return e;
    }
    auto get = [&](CUfunction* f, const std::string& name) -> std::string {
        return check(drv_.cuModuleGetFunction(f, d.mod, name.c_str()), name.c_str());
    };
    for (int t = 0; t < 3 && e.empty(); ++t) {
        e = get(&d.fwd_simt[t], std::string("aule_fwd_simt_") + kDtypeSuffix[t]);
        if (e.empty()) e = get(&d.bwd_dq_simt[t], std::string("aule_bwd_dq_simt_") + kDtypeSuffix[t]);
        if (e.empty()) e = get(&d.bwd_dkv_simt[t], std::string("aule_bwd_dkv_simt_") + kDtypeSuffix[t]);
    }
    for (int t = 1; t < 3 && e.empty(); ++t) {
        e = get(&d.fwd_sm100[t][0], std::string("aule_fwd_sm100_") + kDtypeSuffix[t] + "_d64");
        if (e.empty()) e = get(&d.fwd_sm100[t][1], std::string("aule_fwd_sm100_") + kDtypeSuffix[t] + "_d128");
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.fwd_sm100[t][0], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)FwdCfg<64>::SMEM_BYTES), "cuFuncSetAttribute(smem d64)");
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.fwd_sm100[t][1], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)FwdCfg<128>::SMEM_BYTES), "cuFuncSetAttribute(smem d128)");
    }
    for (int t = 1; t < 3 && e.empty(); ++t) {
        // v3 dK/dV kernel: tuning builds only (optional symbols)
        for (int dd = 0; dd < 2; ++dd) {
            const std::string nm = std::string("aule_bwd_sm100_") + kDtypeSuffix[t] + (dd ? "_d128" : "_d64");
            if (drv_.cuModuleGetFunction(&d.bwd_sm100[t][dd], d.mod, nm.c_str()) != CUDA_SUCCESS) { d.bwd_sm100[t][dd] = nullptr; continue; }
            drv_.cuFuncSetAttribute(d.bwd_sm100[t][dd], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                    dd ? (int)BwdCfg<128>::SMEM_BYTES : (int)BwdCfg<64>::SMEM_BYTES);
        }
        e = get(&d.bwd_delta[t], std::string("aule_bwd_delta_") + kDtypeSuffix[t]);
        if (e.empty()) e = get(&d.bwd_dkvt_sm100[t][0], std::string("aule_bwd_dkvt_sm100_") + kDtypeSuffix[t] + "_d64");
        if (e.empty()) e = get(&d.bwd_dkvt_sm100[t][1], std::string("aule_bwd_dkvt_sm100_") + kDtypeSuffix[t] + "_d128");
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.bwd_dkvt_sm100[t][0], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)aule_kp::BwdTCfg<64>::SMEM_BYTES), "cuFuncSetAttribute(smem bwd dkvt d64)");
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.bwd_dkvt_sm100[t][1], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)aule_kp::BwdTCfg<128>::SMEM_BYTES), "cuFuncSetAttribute(smem bwd dkvt d128)");
        if (e.empty()) e = get(&d.bwd_dq_sm100[t][0], std::string("aule_bwd_dq_sm100_") + kDtypeSuffix[t] + "_d64");
        if (e.empty()) e = get(&d.bwd_dq_sm100[t][1], std::string("aule_bwd_dq_sm100_") + kDtypeSuffix[t] + "_d128");
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.bwd_dq_sm100[t][0], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)aule_kp::BwdDq2Cfg<64>::SMEM_BYTES), "cuFuncSetAttribute(smem bwd dq d64)");
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.bwd_dq_sm100[t][1], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)aule_kp::BwdDq2Cfg<128>::SMEM_BYTES), "cuFuncSetAttribute(smem bwd dq d128)");
        for (int dd = 0; dd < 2; ++dd) {     // v1 dQ kernel: tuning builds only (optional symbols)
            const std::string nm = std::string("aule_bwd_dq1_sm100_") + kDtypeSuffix[t] + (dd ? "_d128" : "_d64");
            if (drv_.cuModuleGetFunction(&d.bwd_dq1_sm100[t][dd], d.mod, nm.c_str()) != CUDA_SUCCESS) { d.bwd_dq1_sm100[t][dd] = nullptr; continue; }
            drv_.cuFuncSetAttribute(d.bwd_dq1_sm100[t][dd], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                    dd ? (int)aule_kp::BwdDqCfg<128>::SMEM_BYTES : (int)aule_kp::BwdDqCfg<64>::SMEM_BYTES);
        }
    }
    if (e.empty()) e = get(&d.fwd_tf32, "aule_fwd_sm100_tf32_d64");
    if (e.empty())
        e = check(drv_.cuFuncSetAttribute(d.fwd_tf32, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                          (int)aule_kp::FwdCfgT32::SMEM_BYTES), "cuFuncSetAttribute(smem fwd tf32)");
    for (int t = 1; t < 3 && e.empty(); ++t) {
        e = get(&d.bwd_fused_sm100[t], std::string("aule_bwd_fused_sm100_") + kDtypeSuffix[t] + "_d128");
        if (e.empty()) e = get(&d.bwd_dq_convert[t], std::string("aule_bwd_dq_convert_") + kDtypeSuffix[t]);
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.bwd_fused_sm100[t], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)aule_kp::BwdFCfg<128>::SMEM_BYTES), "cuFuncSetAttribute(smem bwd fused d128)");
    }
    for (int t = 1; t < 3 && e.empty(); ++t) {
        e = get(&d.bwd_fused2_sm100[t], std::string("aule_bwd_fused2_sm100_") + kDtypeSuffix[t] + "_d128");
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.bwd_fused2_sm100[t], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)aule_kp::BwdF2Cfg<128>::SMEM_BYTES), "cuFuncSetAttribute(smem bwd fused2 d128)");
    }
    // tuning builds only: optional symbols
    for (int dd = 0; dd < 2; ++dd) {
        for (int v = 0; v < 16; ++v) {
            const std::string nm = std::string("aule_fwd_sm100_bf16_d") + (dd ? "128" : "64") + "_e" + std::to_string(v);
            if (drv_.cuModuleGetFunction(&d.fwd_sm100_var[dd][v], d.mod, nm.c_str()) != CUDA_SUCCESS) { d.fwd_sm100_var[dd][v] = nullptr; continue; }
            drv_.cuFuncSetAttribute(d.fwd_sm100_var[dd][v], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                    dd ? (int)FwdCfg<128>::SMEM_BYTES : (int)FwdCfg<64>::SMEM_BYTES);
        }
        const std::string nm4 = std::string("aule_fwd4_sm100_bf16_d") + (dd ? "128" : "64");
        if (drv_.cuModuleGetFunction(&d.fwd4_sm100[dd], d.mod, nm4.c_str()) != CUDA_SUCCESS) d.fwd4_sm100[dd] = nullptr;
        else drv_.cuFuncSetAttribute(d.fwd4_sm100[dd], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                     dd ? (int)aule_kp::FwdCfg4<128>::SMEM_BYTES : (int)aule_kp::FwdCfg4<64>::SMEM_BYTES);
    }
    for (int t = 0; t < 3 && e.empty(); ++t) e = get(&d.rope[t], std::string("aule_rope_") + kDtypeSuffix[t]);
    for (int t = 1; t < 3 && e.empty(); ++t) {
        e = get(&d.paged[t][0], std::string("aule_paged_sm100_") + kDtypeSuffix[t] + "_d64");
        if (e.empty()) e = get(&d.paged[t][1], std::string("aule_paged_sm100_") + kDtypeSuffix[t] + "_d128");
        if (e.empty()) e = get(&d.paged_combine[t], std::string("aule_paged_combine_") + kDtypeSuffix[t]);
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.paged[t][0], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)aule_kp::PagedCfg<64>::SMEM_BYTES), "cuFuncSetAttribute(smem paged d64)");
        if (e.empty())
            e = check(drv_.cuFuncSetAttribute(d.paged[t][1], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                              (int)aule_kp::PagedCfg<128>::SMEM_BYTES), "cuFuncSetAttribute(smem paged d128)");
    }
    if (e.empty()) e = get(&d.smoke, "aule_smoke_multiply");
    if (e.empty()) e = check(drv_.cuMemAlloc(&d.sched, Device::kSchedSlots * sizeof(uint32_t)), "cuMemAlloc(scheduler counters)");
    for (uint32_t i = 0; i < Device::kSchedSlots && e.empty(); ++i)
        e = check(drv_.cuEventCreate(&d.sched_ev[i], CU_EVENT_DISABLE_TIMING), "cuEventCreate");
    for (uint32_t i = 0; i < Device::kMaxChunks && e.empty(); ++i) {
        e = check(drv_.cuEventCreate(&d.ev_in[i], CU_EVENT_DISABLE_TIMING), "cuEventCreate");
        if (e.empty()) e = check(drv_.cuEventCreate(&d.ev_c[i], CU_EVENT_DISABLE_TIMING), "cuEventCreate");
    }
    for (int i = 0; i < 4 && e.empty(); ++i) e = check(drv_.cuEventCreate(&d.bounce_ev[i], CU_EVENT_DISABLE_TIMING), "cuEventCreate");
    if (e.empty()) {
        // The backward's Delta workspace comes from the stream-ordered pool (cuMemAllocAsync). With the default
        // release threshold (0) the pool hands its memory back to the OS at every synchronisation and the next
        // call pays a fresh physical allocation (~0.5-1 ms, measured on config E); keep it cached instead.
        CUmemoryPool pool = nullptr;
        if (drv_.cuDeviceGetDefaultMemPool(&pool, d.dev) == CUDA_SUCCESS && pool) {
            cuuint64_t keep = ~0ull;
            drv_.cuMemPoolSetAttribute(pool, CU_MEMPOOL_ATTR_RELEASE_THRESHOLD, &keep);
        }
    }
    if (e.empty()) e = check(drv_.cuStreamCreate(&d.s_in, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
    if (e.empty()) e = check(drv_.cuStreamCreate(&d.s_compute, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
    if (e.empty()) e = check(drv_.cuStreamCreate(&d.s_out, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
    if (e.empty()) e = check(drv_.cuStreamCreate(&d.s_aux, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
    if (e.empty()) e = check(drv_.cuEventCreate(&d.ev_fork, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
    if (e.empty()) e = check(drv_.cuEventCreate(&d.ev_join, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
    if (!e.empty()) {
        drv_.cuModuleUnload(d.mod);
        drv_.cuDevicePrimaryCtxRelease(d.dev);
        return e;
    }
    if ((int)devices_.size() >= kMaxDevices) return "too many devices";
    d.index = (int)devices_.size();
    devices_.push_back(d);
    return "";
}

void Engine::shutdown() {
    if (!ready_) return;
    for (Device& d : devices_) {
        CtxGuard g(drv_, d.ctx);
        drv_.cuCtxSynchronize();
        for (int i = 0; i < 9; ++i)
            if (d.stage[i]) drv_.cuMemFree(d.stage[i]);
        if (d.sched) drv_.cuMemFree(d.sched);
        for (CUevent ev : d.sched_ev) if (ev) drv_.cuEventDestroy(ev);
        for (CUevent ev : d.ev_in) if (ev) drv_.cuEventDestroy(ev);
        for (CUevent ev : d.ev_c) if (ev) drv_.cuEventDestroy(ev);
        for (CUevent ev : d.bounce_ev) if (ev) drv_.cuEventDestroy(ev);
        for (int i = 0; i < 4; ++i) if (d.bounce[i]) drv_.cuMemFreeHost(d.bounce[i]);
        if (d.s_in) drv_.cuStreamDestroy(d.s_in);
        if (d.s_compute) drv_.cuStreamDestroy(d.s_compute);
        if (d.s_out) drv_.cuStreamDestroy(d.s_out);
        if (d.s_aux) drv_.cuStreamDestroy(d.s_aux);
        if (d.ev_fork) drv_.cuEventDestroy(d.ev_fork);
        if (d.ev_join) drv_.cuEventDestroy(d.ev_join);
        if (d.mod) drv_.cuModuleUnload(d.mod);
    }
    for (Device& d : devices_) drv_.cuDevicePrimaryCtxRelease(d.dev);
    devices_.clear();
    ready_ = false;
    // The driver library stays loaded (dlclose of libcuda is not safe with live CUDA users
    // such as PyTorch in the same process).
}

Device* Engine::by_ordinal(int ordinal) {
    for (Device& d : devices_)
        if (d.ordinal == ordinal) return &d;
    return nullptr;
}

std::string Engine::validate(const AttnShape& s, int32_t dtype) {
    char buf[256];
    if (dtype < 0 || dtype > 3) return "unsupported dtype (0=f32, 1=bf16, 2=f16, 3=f32 tensors with tf32 tensor-core math allowed)";
    if (!s.B || !s.Hq || !s.Hkv || !s.Sq || !s.Sk || !s.D) return "empty tensor dimension";
    if (s.Hq % s.Hkv != 0) {   // attention_gpu.zig:383-388, __init__.py:159-160
        snprintf(buf, sizeof(buf), "heads_q (%u) must be divisible by heads_kv (%u) for GQA", s.Hq, s.Hkv);
        return buf;
    }
    if (s.D > 128 || (s.D % 4) != 0) {   // README.md:202-206 (head_dim <= 128)
        snprintf(buf, sizeof(buf), "head_dim must be a multiple of 4 and <= 128, got %u", s.D);
        return buf;
    }
    return "";
}

std::string Engine::launch(Device& d, CUfunction fn, const char* name, unsigned gx, unsigned gy, unsigned gz,
                           unsigned bx, unsigned smem, CUstream stream, void** params) {
    std::string e = check(drv_.cuLaunchKernel(fn, gx, gy, gz, bx, 1, 1, smem, stream, params, nullptr), name);
    if (e.empty()) {
        ++launches_;
        std::lock_guard<std::mutex> g(name_mu_);
        last_kernel_ = name;
    }
    (void)d;
    return e;
}

std::string Engine::make_tmap(CUtensorMap* m, int32_t dtype, CUdeviceptr base, uint64_t bh, uint64_t S, uint32_t D,
                              bool mn_major_f32, uint32_t box_rows) const {
    // [bh, S, D] row-major viewed as a 3-D tensor (D innermost); box = 128 bytes x 128 rows x 1 (64 16-bit or 32 fp32
    // elements) with the 128-byte swizzle the UMMA descriptors in attn_fwd_sm100.cu expect. Out-of-range rows / columns
    // read as 0 and are clipped on store, which is how ragged Sq/Sk tails and padded head dims are handled.
    // mn_major_f32: the V operand of the tf32 kernel -- 32-bit MN-major operands need the 128-byte swizzle with 32-byte atoms.
    const bool f32 = dtype == kF32 || dtype == kTF32;
    const cuuint64_t es = f32 ? 4 : 2;
    cuuint64_t dims[3] = {D, S, bh};
    cuuint64_t strides[2] = {(cuuint64_t)D * es, (cuuint64_t)S * D * es};
    cuuint32_t box[3] = {f32 ? 32u : 64u, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = drv_.cuTensorMapEncodeTiled(
        m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
        3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        mn_major_f32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return check(r, "cuTensorMapEncodeTiled");
}

std::string Engine::forward(int dev, CUstream stream, CUdeviceptr q, CUdeviceptr k, CUdeviceptr v, CUdeviceptr o,
                            CUdeviceptr lse, const AttnShape& s, int32_t dtype, float scale, bool causal,
                            int32_t window) {
    if (!ready_) return "Library not initialized. Call aule_init() first.";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    std::string e = validate(s, dtype);
    if (!e.empty()) return e;
    if (!q || !k || !v || !o) return "null device pointer";
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    if (!(scale > 0.f)) scale = 1.0f / sqrtf((float)s.D);   // triton_flash.py:394-395
    if (window == 0) window = -1;

    // Tensor-core path: 16-bit, any head_dim <= 128 that is a multiple of 8 (the TMA boxes zero-fill D up to 64 / 128, the
    // way the reference pads to BLOCK_K = next_power_of_2(D), triton_flash.py:446), causal / full / sliding-window
    // (one- or two-sided) masks.  Everything else (fp32, D % 8 != 0, unaligned pointers) runs the CUDA-core kernels;
    // AULE_LOG=1 reports that choice on stderr.
    // fp32 tensors whose caller allows tf32 (dtype code 3): the tensor-core kernel of attn_fwd_tf32_sm100.cu for head_dim <= 64
    // (P stays fp32 in tensor memory: no room for D = 128); otherwise exactly the fp32 path.
    const bool tf32 = dtype == kTF32 && s.D <= 64 && path_ != kForceCudaCore && ((q | k | v | o) & 15) == 0;
    if (dtype == kTF32 && !tf32) dtype = kF32;
    const bool tc = tf32 || ((dtype == kBF16 || dtype == kF16) && (s.D % 8 == 0) && path_ != kForceCudaCore && ((q | k | v | o) & 15) == 0);
    if (!tc && log_enabled())
        fprintf(stderr, "[aule] forward [%u,%u(%u),%u/%u,%u] %s: CUDA-core kernel (%s)\n", s.B, s.Hq, s.Hkv, s.Sq, s.Sk, s.D,
                kDtypeSuffix[dtype], path_ == kForceCudaCore ? "forced" : dtype == kF32 ? "fp32 inputs" : (s.D % 8) ? "head_dim % 8 != 0" : "unaligned pointers");
    if (tc) {
        const uint32_t DP = s.D <= 64 ? 64 : 128;             // kernel (padded) head_dim
        CUtensorMap tmQ, tmK, tmV, tmO;
        if (!(e = make_tmap(&tmQ, dtype, q, (uint64_t)s.B * s.Hq, s.Sq, s.D)).empty()) return e;
        if (!(e = make_tmap(&tmK, dtype, k, (uint64_t)s.B * s.Hkv, s.Sk, s.D)).empty()) return e;
        if (!(e = make_tmap(&tmV, dtype, v, (uint64_t)s.B * s.Hkv, s.Sk, s.D, tf32)).empty()) return e;
        if (fwd_v4_ && !(e = make_tmap(&tmO, dtype, o, (uint64_t)s.B * s.Hq, s.Sq, s.D)).empty()) return e;
        FwdParams p;
        p.lse = (float*)lse;
        p.B = s.B; p.Hq = s.Hq; p.Hkv = s.Hkv; p.Sq = s.Sq; p.Sk = s.Sk;
        // GQA/MQA groups with an even number of q-heads: a work item pairs two q-heads of one KV head
        // (equal trip counts, shared K/V tiles); otherwise 256 rows of one head.
        p.pair_heads = ((s.Hq / s.Hkv) % 2 == 0 && pair_heads_enabled_) ? 1 : 0;
        p.o = (void*)o;
        p.num_q_super = p.pair_heads ? (s.Sq + 127) / 128 : (s.Sq + 255) / 256;
        const uint64_t tiles = (uint64_t)p.num_q_super * (p.pair_heads ? s.Hq / 2 : s.Hq) * s.B;
        if (tiles > 0xffffffffull) return "problem too large (work-item count exceeds 2^32)";
        p.num_tiles = (uint32_t)tiles;
        {   // keep ~32 MiB of K/V (two tensors, 16-bit) live per scheduling run so that it stays in the 126 MB L2
            const uint64_t unit_bytes = 2ull * s.Sk * s.D * (tf32 ? 4ull : 2ull);
            const uint64_t upr = std::max<uint64_t>(1, (32ull << 20) / std::max<uint64_t>(unit_bytes, 1));
            p.units_per_run = (uint32_t)std::min<uint64_t>(upr, (uint64_t)s.B * s.Hkv);
            if (!l2_runs_enabled_) p.units_per_run = s.B * s.Hkv;
        }
        p.scale = scale;
        p.scale_log2 = scale * 1.4426950408889634f;
        p.causal = causal ? 1 : 0;
        p.window = causal ? window : -1;
        p.D_real = s.D;
        // visibility i - win_left <= j <= i + win_right: causal keeps i-j < W (attention_f32.comp:176-178), the
        // bidirectional window keeps |i-j| <= W/2 (:180-183)
        p.win_right = causal ? 0u : (window > 0 ? (uint32_t)window / 2 : aule_kp::kWinInf);
        p.win_left = window > 0 ? (causal ? (uint32_t)window - 1 : (uint32_t)window / 2) : aule_kp::kWinInf;
        // one zeroed work counter per launch, from a ring; the slot's previous user (any stream) must have finished
        std::lock_guard<std::mutex> launch_guard(launch_mu_[d.index]);
        const uint32_t slot = d.sched_next++ % Device::kSchedSlots;
        if (d.sched_ev_live[slot] && !(e = check(drv_.cuEventSynchronize(d.sched_ev[slot]), "cuEventSynchronize(scheduler slot)")).empty()) return e;
        const CUdeviceptr counter = d.sched + 4ull * slot;
        if (!(e = check(drv_.cuMemsetD32Async(counter, 0, 1, stream), "cuMemsetD32Async(scheduler counter)")).empty()) return e;
        p.sched_counter = (uint32_t*)counter;
        auto launch_fwd = [&](CUfunction f, const char* nm, unsigned g_, unsigned smem_, void** prm) -> std::string {
            std::string le = launch(d, f, nm, g_, 1, 1, 512, smem_, stream, prm);
            if (le.empty()) {
                le = check(drv_.cuEventRecord(d.sched_ev[slot], stream), "cuEventRecord(scheduler slot)");
                d.sched_ev_live[slot] = le.empty();
            }
            return le;
        };
        p.cross_item = cross_item_enabled_ ? 1 : 0;
        p.trace = (unsigned long long*)trace_;
        const bool d128 = DP == 128;
        const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)d.sm_count);
        char name[64];
        if (fwd_v4_) {                     // A/B hook (tuning builds): the v4 kernel, TMA-stored output
            if (dtype != kBF16 || !d.fwd4_sm100[d128 ? 1 : 0] || s.D != DP || (window > 0 && !causal))
                return "the v4 forward kernel is only present in tuning builds (bf16, D in {64,128}, no bidirectional window)";
            void* params4[] = {&tmQ, &tmK, &tmV, &tmO, &p};
            snprintf(name, sizeof(name), "aule_fwd4_sm100_bf16_d%u", DP);
            return launch_fwd(d.fwd4_sm100[d128 ? 1 : 0], name, grid,
                              d128 ? aule_kp::FwdCfg4<128>::SMEM_BYTES : aule_kp::FwdCfg4<64>::SMEM_BYTES, params4);
        }
        if (tf32) {
            // operand-truncation bias of the tf32 MMA, twice in Q K^T (see the kernel's epilogue)
            p.scale *= 1.00072f;
            p.scale_log2 *= 1.00072f;
            if (fwd_v4_ || (path_ >= kVariantBase && path_ < kVariantBase + 16)) return "forward tuning variants are bf16 only";
            void* params32[] = {&tmQ, &tmK, &tmV, &p};
            return launch_fwd(d.fwd_tf32, "aule_fwd_sm100_tf32_d64", grid, aule_kp::FwdCfgT32::SMEM_BYTES, params32);
        }
        const unsigned smem = d128 ? FwdCfg<128>::SMEM_BYTES : FwdCfg<64>::SMEM_BYTES;
        void* params[] = {&tmQ, &tmK, &tmV, &p};
        snprintf(name, sizeof(name), "aule_fwd_sm100_%s_d%u", kDtypeSuffix[dtype], DP);
        CUfunction fn = d.fwd_sm100[dtype][d128 ? 1 : 0];
        if (path_ >= kVariantBase && path_ < kVariantBase + 16) {
            if (dtype != kBF16 || !d.fwd_sm100_var[d128 ? 1 : 0][path_ - kVariantBase]) return "forward tuning variants are only present in tuning builds (bf16)";
            fn = d.fwd_sm100_var[d128 ? 1 : 0][path_ - kVariantBase];
            snprintf(name, sizeof(name), "aule_fwd_sm100_bf16_d%u_e%d", DP, path_ - kVariantBase);
        }
        return launch_fwd(fn, name, grid, smem, params);
    }
    SimtParams p;
    memset(&p, 0, sizeof(p));
    p.q = (const void*)q; p.k = (const void*)k; p.v = (const void*)v; p.o = (void*)o;
    p.lse = (float*)lse;
    p.B = s.B; p.Hq = s.Hq; p.Hkv = s.Hkv; p.Sq = s.Sq; p.Sk = s.Sk; p.D = s.D;
    p.scale = scale; p.causal = causal ? 1 : 0; p.window = window;
    if (s.Hq > 65535 || s.B > 65535) return "heads/batch exceed the CUDA grid limit (65535)";
    void* params[] = {&p};
    char name[64];
    snprintf(name, sizeof(name), "aule_fwd_simt_%s", kDtypeSuffix[dtype]);
    return launch(d, d.fwd_simt[dtype], name, (s.Sq + aule_kp::SIMT_ROWS - 1) / aule_kp::SIMT_ROWS, s.Hq, s.B,
                  aule_kp::SIMT_ROWS * aule_kp::SIMT_LANES, 0, stream, params);
}

std::string Engine::backward(int dev, CUstream stream, CUdeviceptr q, CUdeviceptr k, CUdeviceptr v, CUdeviceptr o,
                             CUdeviceptr d_o, CUdeviceptr lse, CUdeviceptr dq, CUdeviceptr dk, CUdeviceptr dv,
                             const AttnShape& s, int32_t dtype, float scale, bool causal, int32_t window) {
    if (!ready_) return "Library not initialized. Call aule_init() first.";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    std::string e = validate(s, dtype);
    if (!e.empty()) return e;
    if (!q || !k || !v || !o || !d_o || !lse || !dq || !dk || !dv) return "null device pointer";
    if (s.Hq > 65535 || s.B > 65535) return "heads/batch exceed the CUDA grid limit (65535)";
    if (dtype == kTF32) dtype = kF32;                     // the fp32 backward is exact (CUDA cores)
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    if (!(scale > 0.f)) scale = 1.0f / sqrtf((float)s.D);
    CUdeviceptr delta = 0;
    const size_t dbytes = (size_t)s.B * s.Hq * s.Sq * sizeof(float);
    if (!(e = check(drv_.cuMemAllocAsync(&delta, dbytes, stream), "cuMemAllocAsync(delta)")).empty()) return e;

    if (window == 0) window = -1;
    // Tensor-core backward: 16-bit, any head_dim <= 128 that is a multiple of 8 (the kernels run at 64 / 128: the TMA boxes
    // zero-fill the padding columns of Q, K, V, dO and clip them from dK, dV; the dQ kernel predicates its own loads and
    // stores), causal / full masks.  A sliding window (the reference's backward ignores it, triton_flash.py:313-319 -- its
    // gradients are then wrong), fp32 and head dims that are not a multiple of 8 run the deterministic CUDA-core kernels.
    const bool legacy_kernels = bwd_dq_v1_ || (bwd_serial_ & 2);       // tuning builds: v1 dQ / v3 dK/dV kernels know neither
    const bool tc = (dtype == kBF16 || dtype == kF16) && (s.D % 8 == 0) && path_ != kForceCudaCore &&
                    ((q | k | v | o | d_o | dq | dk | dv) & 15) == 0 && !(legacy_kernels && ((s.D != 64 && s.D != 128) || window > 0));
    if (!tc && log_enabled())
        fprintf(stderr, "[aule] backward [%u,%u(%u),%u/%u,%u] %s: CUDA-core kernels (%s)\n", s.B, s.Hq, s.Hkv, s.Sq, s.Sk, s.D,
                kDtypeSuffix[dtype], path_ == kForceCudaCore ? "forced" : dtype == kF32 ? "fp32 inputs" : "head_dim % 8 != 0 or unaligned pointers");
    if (tc) {
        // Tensor-core backward: Delta pre-pass, then the dK/dV kernel (key block outer) and the dQ kernel (query block
        // outer).  No atomics and no workspace beyond Delta: bit-reproducible.
        const bool d128 = s.D > 64;                        // kernel (padded) head_dim 128, else 64
        const uint32_t DP = d128 ? 128u : 64u;
        char name[64];
        {
            uint64_t rows = (uint64_t)s.B * s.Hq * s.Sq;
            uint32_t D = s.D;
            void* params[] = {&o, &d_o, &delta, &rows, &D};
            snprintf(name, sizeof(name), "aule_bwd_delta_%s", kDtypeSuffix[dtype]);
            // 256 threads = 8 warps; a warp covers (32 / lanes-per-row) rows x 4 rows in flight (delta_body)
            const uint64_t rows_per_cta = 8ull * (32u / (s.D <= 32 ? 4u : (s.D <= 64 ? 8u : 16u))) * 4ull;
            e = launch(d, d.bwd_delta[dtype], name, (unsigned)((rows + rows_per_cta - 1) / rows_per_cta), 1, 1, 256, 0, stream, params);
        }
        std::lock_guard<std::mutex> launch_guard(launch_mu_[d.index]);   // s_aux and the fork / join events are per device
        if (e.empty() && bwd_two_streams_) e = check(drv_.cuEventRecord(d.ev_fork, stream), "cuEventRecord");
        if (e.empty()) {
            CUtensorMap tmQ, tmK, tmV, tmdO, tmdK, tmdV;
            e = make_tmap(&tmQ, dtype, q, (uint64_t)s.B * s.Hq, s.Sq, s.D);
            if (e.empty()) e = make_tmap(&tmdO, dtype, d_o, (uint64_t)s.B * s.Hq, s.Sq, s.D);
            if (e.empty()) e = make_tmap(&tmK, dtype, k, (uint64_t)s.B * s.Hkv, s.Sk, s.D);
            if (e.empty()) e = make_tmap(&tmV, dtype, v, (uint64_t)s.B * s.Hkv, s.Sk, s.D);
            if (e.empty()) e = make_tmap(&tmdK, dtype, dk, (uint64_t)s.B * s.Hkv, s.Sk, s.D);
            if (e.empty()) e = make_tmap(&tmdV, dtype, dv, (uint64_t)s.B * s.Hkv, s.Sk, s.D);
            BwdParams bp;
            bp.dq_out = (void*)dq; bp.q = (const void*)q; bp.d_o = (const void*)d_o; bp.lse = (const float*)lse; bp.delta = (const float*)delta;
            bp.B = s.B; bp.Hq = s.Hq; bp.Hkv = s.Hkv; bp.Sq = s.Sq; bp.Sk = s.Sk; bp.D_real = s.D;
            // same band as the forward (attention_f32.comp:173-183): causal window keeps 0 <= i-j < W, bidirectional |i-j| <= W/2
            bp.win_right = causal ? 0u : (window > 0 ? (uint32_t)window / 2 : aule_kp::kWinInf);
            bp.win_left = window > 0 ? (causal ? (uint32_t)window - 1 : (uint32_t)window / 2) : aule_kp::kWinInf;
            bp.scale = scale; bp.scale_log2 = scale * 1.4426950408889634f; bp.causal = causal ? 1 : 0;
            bp.order = bwd_serial_ | (bwd_legacy_poll_ ? 4 : 0) | (bwd_single_s_ ? 8 : 0) | (bwd_consumer_fence_ ? 16 : 0) | (bwd_fencer_ ? 32 : 0);
            // dK/dV CTA order: all units at once.  Launching the KV-block CTAs of a few (batch, kv-head) units together
            // (Q/dO L2-resident, BwdParams::units_per_run = ceil(SMs / KV blocks)) measured SLOWER: 1.36 vs 1.21 ms on
            // config C/2 (gpurun s21) -- the heavy CTAs of later runs start late and the tail grows.
            bp.units_per_run = 0;
            bp.trace = (unsigned long long*)trace_;
            bp.dq_acc = nullptr;
            const uint64_t ctas = (uint64_t)((s.Sk + 127) / 128) * s.Hkv * s.B;
            const uint64_t ctas_dq = (uint64_t)((s.Sq + 127) / 128) * s.Hq * s.B;
            if (e.empty() && (ctas > 0x7fffffffull || ctas_dq > 0x7fffffffull)) e = "problem too large (backward grid exceeds 2^31 CTAs)";
            if (e.empty() && (bwd_fused_ || bwd_fused2_) && s.D == 128 && bwd_order_ == 0 && window < 0) {
                // Fused backward (attn_bwd_fused_sm100.cu): dK / dV as below, dQ reduced into an fp32 accumulator that is
                // zeroed here and converted (x scale) afterwards.
                const size_t n = (size_t)s.B * s.Hq * s.Sq * s.D;
                CUdeviceptr acc = 0;
                e = check(drv_.cuMemAllocAsync(&acc, n * sizeof(float), stream), "cuMemAllocAsync(dQ accumulator)");
                if (e.empty()) e = check(drv_.cuMemsetD32Async(acc, 0, n, stream), "cuMemsetD32Async(dQ accumulator)");
                bp.dq_acc = (float*)acc;
                // CTA runs: the fp32 dQ rows a run reduces into (units x group x Sq x D x 4 bytes) should stay L2-resident
                // while the run lasts, and the last run should be large enough that its heaviest CTAs do not make a tail.
                bp.units_per_run = bwd_units_per_run_ ? (uint32_t)bwd_units_per_run_ : 8u;
                if (e.empty() && bwd_fused2_) {
                    // 64-query half steps: Q / dO travel in [64 rows][64 columns] boxes
                    CUtensorMap tmQh, tmdOh;
                    e = make_tmap(&tmQh, dtype, q, (uint64_t)s.B * s.Hq, s.Sq, s.D, false, 64);
                    if (e.empty()) e = make_tmap(&tmdOh, dtype, d_o, (uint64_t)s.B * s.Hq, s.Sq, s.D, false, 64);
                    if (e.empty()) {
                        void* params[] = {&tmQh, &tmK, &tmV, &tmdOh, &tmdK, &tmdV, &bp};
                        snprintf(name, sizeof(name), "aule_bwd_fused2_sm100_%s_d128", kDtypeSuffix[dtype]);
                        e = launch(d, d.bwd_fused2_sm100[dtype], name, (unsigned)ctas, 1, 1, (unsigned)aule_kp::BwdF2Cfg<128>::THREADS,
                                   aule_kp::BwdF2Cfg<128>::SMEM_BYTES, stream, params);
                    }
                } else if (e.empty()) {
                    void* params[] = {&tmQ, &tmK, &tmV, &tmdO, &tmdK, &tmdV, &bp};
                    snprintf(name, sizeof(name), "aule_bwd_fused_sm100_%s_d128", kDtypeSuffix[dtype]);
                    e = launch(d, d.bwd_fused_sm100[dtype], name, (unsigned)ctas, 1, 1, (unsigned)aule_kp::BwdFCfg<128>::THREADS,
                               aule_kp::BwdFCfg<128>::SMEM_BYTES, stream, params);
                }
                if (e.empty()) {
                    uint64_t n8 = n / 8;
                    float sc = scale;
                    void* params[] = {&acc, &dq, &n8, &sc};
                    snprintf(name, sizeof(name), "aule_bwd_dq_convert_%s", kDtypeSuffix[dtype]);
                    const unsigned blocks = (unsigned)std::min<uint64_t>((n8 + 255) / 256, (uint64_t)d.sm_count * 16);
                    e = launch(d, d.bwd_dq_convert[dtype], name, blocks, 1, 1, 256, 0, stream, params);
                }
                if (acc) drv_.cuMemFreeAsync(acc, stream);
                drv_.cuMemFreeAsync(delta, stream);
                return e;
            }
            if (e.empty() && bwd_order_ != 2) {                 // (timing hook: 2 = dQ kernel only)
                void* params[] = {&tmQ, &tmK, &tmV, &tmdO, &tmdK, &tmdV, &bp};
                if (bwd_serial_ & 2) {      // A/B hook (path bit 13): the v3 dK/dV kernel (P, dS staged through shared memory)
                    if (!d.bwd_sm100[dtype][d128 ? 1 : 0]) { drv_.cuMemFreeAsync(delta, stream); return "the v3 dK/dV kernel is only present in tuning builds"; }
                    snprintf(name, sizeof(name), "aule_bwd_sm100_%s_d%u", kDtypeSuffix[dtype], DP);
                    e = launch(d, d.bwd_sm100[dtype][d128 ? 1 : 0], name, (unsigned)ctas, 1, 1, (unsigned)BwdCfg<128>::THREADS,
                               d128 ? BwdCfg<128>::SMEM_BYTES : BwdCfg<64>::SMEM_BYTES, stream, params);
                } else {                    // v4: transposed score tiles, P^T / dS^T stay in TMEM
                    snprintf(name, sizeof(name), "aule_bwd_dkvt_sm100_%s_d%u", kDtypeSuffix[dtype], DP);
                    e = launch(d, d.bwd_dkvt_sm100[dtype][d128 ? 1 : 0], name, (unsigned)ctas, 1, 1,
                               (unsigned)aule_kp::BwdTCfg<128>::THREADS,
                               d128 ? aule_kp::BwdTCfg<128>::SMEM_BYTES : aule_kp::BwdTCfg<64>::SMEM_BYTES, stream, params);
                }
            }
            if (e.empty() && bwd_order_ != 1) {                 // (timing hook: 1 = dK/dV kernel only)
                // The dQ kernel is independent of the dK/dV kernel (both only read q,k,v,dO,LSE,Delta): it runs on a second
                // stream so that its CTAs fill the SMs the dK/dV kernel's last wave leaves idle (and vice versa) instead of
                // waiting for that kernel's tail.  Fork after the Delta pre-pass, join before returning to the caller's stream.
                const bool fork = bwd_order_ == 0 && !(bwd_serial_ & 1) && bwd_two_streams_;
                CUstream sq = stream;
                if (fork) {
                    sq = d.s_aux;
                    e = check(drv_.cuStreamWaitEvent(sq, d.ev_fork, 0), "cuStreamWaitEvent");
                }
                if (e.empty() && bwd_dq_v1_) {          // A/B hook (path bit 23): the v1 dQ kernel
                    if (!d.bwd_dq1_sm100[dtype][d128 ? 1 : 0]) e = "the v1 dQ kernel is only present in tuning builds";
                    void* params[] = {&tmK, &tmV, &bp};
                    snprintf(name, sizeof(name), "aule_bwd_dq1_sm100_%s_d%u", kDtypeSuffix[dtype], DP);
                    if (e.empty())
                        e = launch(d, d.bwd_dq1_sm100[dtype][d128 ? 1 : 0], name, (unsigned)ctas_dq, 1, 1,
                                   (unsigned)aule_kp::BwdDqCfg<128>::THREADS,
                                   d128 ? aule_kp::BwdDqCfg<128>::SMEM_BYTES : aule_kp::BwdDqCfg<64>::SMEM_BYTES, sq, params);
                } else if (e.empty()) {
                    void* params[] = {&tmK, &tmV, &tmdO, &bp};
                    snprintf(name, sizeof(name), "aule_bwd_dq_sm100_%s_d%u", kDtypeSuffix[dtype], DP);
                    e = launch(d, d.bwd_dq_sm100[dtype][d128 ? 1 : 0], name, (unsigned)ctas_dq, 1, 1,
                               (unsigned)aule_kp::BwdDq2Cfg<128>::THREADS,
                               d128 ? aule_kp::BwdDq2Cfg<128>::SMEM_BYTES : aule_kp::BwdDq2Cfg<64>::SMEM_BYTES, sq, params);
                }
                if (fork) {
                    if (e.empty()) e = check(drv_.cuEventRecord(d.ev_join, sq), "cuEventRecord");
                    if (e.empty()) e = check(drv_.cuStreamWaitEvent(stream, d.ev_join, 0), "cuStreamWaitEvent");
                }
            }
        }
        drv_.cuMemFreeAsync(delta, stream);
        return e;
    }
    SimtParams p;
    memset(&p, 0, sizeof(p));
    p.q = (const void*)q; p.k = (const void*)k; p.v = (const void*)v; p.o = (void*)o;
    p.lse = (float*)lse; p.d_o = (const void*)d_o;
    p.dq = (void*)dq; p.dk = (void*)dk; p.dv = (void*)dv; p.delta = (float*)delta;
    p.B = s.B; p.Hq = s.Hq; p.Hkv = s.Hkv; p.Sq = s.Sq; p.Sk = s.Sk; p.D = s.D;
    p.scale = scale; p.causal = causal ? 1 : 0; p.window = window;
    void* params[] = {&p};
    char name[64];
    const unsigned threads = aule_kp::SIMT_ROWS * aule_kp::SIMT_LANES;
    snprintf(name, sizeof(name), "aule_bwd_dq_simt_%s", kDtypeSuffix[dtype]);
    e = launch(d, d.bwd_dq_simt[dtype], name, (s.Sq + aule_kp::SIMT_ROWS - 1) / aule_kp::SIMT_ROWS, s.Hq, s.B, threads,
               0, stream, params);
    if (e.empty()) {
        snprintf(name, sizeof(name), "aule_bwd_dkv_simt_%s", kDtypeSuffix[dtype]);
        e = launch(d, d.bwd_dkv_simt[dtype], name, (s.Sk + aule_kp::SIMT_ROWS - 1) / aule_kp::SIMT_ROWS, s.Hkv, s.B,
                   threads, 0, stream, params);
    }
    drv_.cuMemFreeAsync(delta, stream);
    return e;
}

std::string Engine::ensure_stage(Device& d, int slot, size_t bytes) {
    if (d.stage_cap[slot] >= bytes) return "";
    if (d.stage[slot]) {
        drv_.cuCtxSynchronize();
        drv_.cuMemFree(d.stage[slot]);
        d.stage[slot] = 0;
        d.stage_cap[slot] = 0;
    }
    std::string e = check(drv_.cuMemAlloc(&d.stage[slot], bytes), "cuMemAlloc(staging)");
    if (e.empty()) d.stage_cap[slot] = bytes;
    return e;
}

bool Engine::host_pinned(const void* p) const {
    // CU_MEMORYTYPE_HOST for page-locked (cuMemHostAlloc / cudaHostRegister / torch pin_memory) allocations; the query
    // fails with CUDA_ERROR_INVALID_VALUE for ordinary pageable memory.
    unsigned int mt = 0;
    return drv_.cuPointerGetAttribute(&mt, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)(uintptr_t)p) == CUDA_SUCCESS &&
           mt == CU_MEMORYTYPE_HOST;
}

std::string Engine::ensure_bounce(Device& d, int slot, size_t bytes) {
    if (d.bounce_cap[slot] >= bytes) return "";
    if (d.bounce[slot]) {
        drv_.cuCtxSynchronize();
        drv_.cuMemFreeHost(d.bounce[slot]);
        d.bounce[slot] = nullptr;
        d.bounce_cap[slot] = 0;
    }
    std::string e = check(drv_.cuMemHostAlloc(&d.bounce[slot], bytes, 0), "cuMemHostAlloc(bounce buffer)");
    if (e.empty()) d.bounce_cap[slot] = bytes;
    return e;
}

std::string Engine::forward_host(int dev, const void* q, const void* k, const void* v, void* o, float* lse,
                                 const AttnShape& s, int32_t dtype, float scale, bool causal, int32_t window,
                                 int* stage_code) {
    *stage_code = -1;
    if (!ready_) return "Library not initialized. Call aule_init() first.";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    *stage_code = -4;
    std::string e = validate(s, dtype);
    if (!e.empty()) return e;
    if (!q || !k || !v || !o) return "null host pointer";
    Device& d = *dp;
    std::lock_guard<std::mutex> host_guard(host_mu_[d.index]);   // staging buffers and the three streams are per device
    CtxGuard g(drv_, d.ctx);
    const size_t es = dtype_size(dtype);
    const uint32_t group = s.Hq / s.Hkv;
    const size_t q_unit = (size_t)group * s.Sq * s.D * es;      // bytes of Q/O per (batch, kv-head) unit
    const size_t kv_unit = (size_t)s.Sk * s.D * es;
    const size_t lse_unit = (size_t)group * s.Sq * sizeof(float);
    const uint32_t units = s.B * s.Hkv;
    *stage_code = -2;
    if (!(e = ensure_stage(d, 0, q_unit * units)).empty()) return e;
    if (!(e = ensure_stage(d, 1, kv_unit * units)).empty()) return e;
    if (!(e = ensure_stage(d, 2, kv_unit * units)).empty()) return e;
    if (!(e = ensure_stage(d, 3, q_unit * units)).empty()) return e;
    if (lse && !(e = ensure_stage(d, 4, lse_unit * units)).empty()) return e;

    // Chunk over units so that copy-in of chunk c+1, the kernel of chunk c and copy-out of chunk c-1 overlap (three streams,
    // persistent events between them).  The call is bound by the H2D copies (Q+K+V in, only O out; PCIe is full duplex), so
    // what is left after the last upload -- the last chunk's kernel and download -- is pure tail: chunk sizes shrink
    // geometrically (a quarter of what remains, down to one unit) so that the tail is one unit's worth, while the early
    // chunks stay large (uniform 8 chunks: 8.85 ms per config-C call against 7.84 ms for the bare copies; 16 / 32 uniform
    // chunks were slower, per-copy overheads).  AULE_HOST_CHUNKS=n forces n uniform chunks (tuning hook).
    std::vector<uint32_t> bounds{0};
    {
        uint32_t forced = 0;
        if (const char* ov = getenv("AULE_HOST_CHUNKS")) { const long v_ = atol(ov); if (v_ > 0) forced = (uint32_t)std::min<long>(v_, Device::kMaxChunks); }
        if (forced) {
            const uint32_t nch = std::min(units, forced), per_ = (units + nch - 1) / nch;
            for (uint32_t u = per_; u < units; u += per_) bounds.push_back(u);
        } else {
            uint32_t done = 0;
            while (done < units && bounds.size() < Device::kMaxChunks) {
                const uint32_t left = units - done;
                done += std::max<uint32_t>(1, (left + 3) / 4);
                if (done < units) bounds.push_back(done);
            }
        }
        bounds.push_back(units);
    }
    const uint32_t nchunks = (uint32_t)bounds.size() - 1;
    uint32_t per = 0;                                              // largest chunk (sizes the bounce buffers)
    for (uint32_t c = 0; c < nchunks; ++c) per = std::max(per, bounds[c + 1] - bounds[c]);

    // Pageable callers (every NumPy caller of the legacy ABI): the DMA engines cannot read pageable memory, and the
    // driver's own fallback is a synchronous staged copy.  Stage through the library's pinned bounce buffers instead,
    // double-buffered per chunk: the CPU fills / drains one buffer while the copy engine works on the other.
    const bool in_pinned = host_pinned(q) && host_pinned(k) && host_pinned(v);
    const bool out_pinned = host_pinned(o) && (!lse || host_pinned(lse));
    const size_t in_chunk = (size_t)per * (q_unit + 2 * kv_unit), out_chunk = (size_t)per * (q_unit + (lse ? lse_unit : 0));
    if (!in_pinned)
        for (int b = 0; b < 2; ++b)
            if (!(e = ensure_bounce(d, b, in_chunk)).empty()) return e;
    if (!out_pinned)
        for (int b = 2; b < 4; ++b)
            if (!(e = ensure_bounce(d, b, out_chunk)).empty()) return e;
    bool bounce_busy[4] = {false, false, false, false};

    struct Pending { bool live = false; uint32_t u0 = 0, nu = 0; int buf = 0; } pend;   // download waiting in a bounce buffer
    auto drain = [&](Pending& pd) -> std::string {            // bounce buffer -> caller's pageable output
        if (!pd.live) return "";
        std::string de = check(drv_.cuEventSynchronize(d.bounce_ev[pd.buf]), "download");
        if (!de.empty()) return de;
        const char* src = (const char*)d.bounce[pd.buf];
        memcpy((char*)o + pd.u0 * q_unit, src, pd.nu * q_unit);
        if (lse) memcpy((char*)lse + pd.u0 * lse_unit, src + pd.nu * q_unit, pd.nu * lse_unit);
        bounce_busy[pd.buf] = false;
        pd.live = false;
        return "";
    };

    int code = 0;
    for (uint32_t c = 0; c < nchunks && e.empty(); ++c) {
        const uint32_t u0 = bounds[c], nu = bounds[c + 1] - bounds[c];
        code = -3;
        if (in_pinned) {
            e = check(drv_.cuMemcpyHtoDAsync(d.stage[0] + u0 * q_unit, (const char*)q + u0 * q_unit, nu * q_unit, d.s_in), "upload Q");
            if (e.empty()) e = check(drv_.cuMemcpyHtoDAsync(d.stage[1] + u0 * kv_unit, (const char*)k + u0 * kv_unit, nu * kv_unit, d.s_in), "upload K");
            if (e.empty()) e = check(drv_.cuMemcpyHtoDAsync(d.stage[2] + u0 * kv_unit, (const char*)v + u0 * kv_unit, nu * kv_unit, d.s_in), "upload V");
        } else {
            const int b = (int)(c & 1);
            if (bounce_busy[b]) e = check(drv_.cuEventSynchronize(d.bounce_ev[b]), "upload");   // its previous DMA has finished
            if (e.empty()) {
                char* bb = (char*)d.bounce[b];
                memcpy(bb, (const char*)q + u0 * q_unit, nu * q_unit);
                memcpy(bb + nu * q_unit, (const char*)k + u0 * kv_unit, nu * kv_unit);
                memcpy(bb + nu * (q_unit + kv_unit), (const char*)v + u0 * kv_unit, nu * kv_unit);
                e = check(drv_.cuMemcpyHtoDAsync(d.stage[0] + u0 * q_unit, bb, nu * q_unit, d.s_in), "upload Q");
                if (e.empty()) e = check(drv_.cuMemcpyHtoDAsync(d.stage[1] + u0 * kv_unit, bb + nu * q_unit, nu * kv_unit, d.s_in), "upload K");
                if (e.empty()) e = check(drv_.cuMemcpyHtoDAsync(d.stage[2] + u0 * kv_unit, bb + nu * (q_unit + kv_unit), nu * kv_unit, d.s_in), "upload V");
                if (e.empty()) e = check(drv_.cuEventRecord(d.bounce_ev[b], d.s_in), "cuEventRecord");
                bounce_busy[b] = true;
            }
        }
        if (e.empty()) e = check(drv_.cuEventRecord(d.ev_in[c], d.s_in), "cuEventRecord");
        if (e.empty()) e = check(drv_.cuStreamWaitEvent(d.s_compute, d.ev_in[c], 0), "cuStreamWaitEvent");
        if (!e.empty()) break;
        code = -4;
        // A chunk of `nu` units is itself a [nu, group, S, D] / [nu, 1, S, D] attention problem.
        AttnShape cs{nu, group, 1, s.Sq, s.Sk, s.D};
        e = forward(dev, d.s_compute, d.stage[0] + u0 * q_unit, d.stage[1] + u0 * kv_unit, d.stage[2] + u0 * kv_unit,
                    d.stage[3] + u0 * q_unit, lse ? d.stage[4] + u0 * lse_unit : 0, cs, dtype, scale, causal, window);
        if (e.empty()) e = check(drv_.cuEventRecord(d.ev_c[c], d.s_compute), "cuEventRecord");
        if (e.empty()) e = check(drv_.cuStreamWaitEvent(d.s_out, d.ev_c[c], 0), "cuStreamWaitEvent");
        if (!e.empty()) break;
        code = -5;
        if (out_pinned) {
            e = check(drv_.cuMemcpyDtoHAsync((char*)o + u0 * q_unit, d.stage[3] + u0 * q_unit, nu * q_unit, d.s_out), "download O");
            if (e.empty() && lse)
                e = check(drv_.cuMemcpyDtoHAsync((char*)lse + u0 * lse_unit, d.stage[4] + u0 * lse_unit, nu * lse_unit, d.s_out), "download LSE");
        } else {
            const int b = 2 + (int)(c & 1);
            Pending mine;
            mine.live = true; mine.u0 = u0; mine.nu = nu; mine.buf = b;
            if (bounce_busy[b]) e = "internal: download bounce buffer still busy";
            char* bb = (char*)d.bounce[b];
            if (e.empty()) e = check(drv_.cuMemcpyDtoHAsync(bb, d.stage[3] + u0 * q_unit, nu * q_unit, d.s_out), "download O");
            if (e.empty() && lse) e = check(drv_.cuMemcpyDtoHAsync(bb + nu * q_unit, d.stage[4] + u0 * lse_unit, nu * lse_unit, d.s_out), "download LSE");
            if (e.empty()) e = check(drv_.cuEventRecord(d.bounce_ev[b], d.s_out), "cuEventRecord");
            bounce_busy[b] = true;
            if (e.empty()) e = drain(pend);                    // the previous chunk's output, while this chunk computes
            pend = mine;
        }
    }
    if (e.empty()) { code = -4; e = check(drv_.cuStreamSynchronize(d.s_compute), "attention kernel"); }
    if (e.empty()) { code = -5; e = check(drv_.cuStreamSynchronize(d.s_out), "download"); }
    if (e.empty()) e = drain(pend);
    if (!e.empty()) { drv_.cuStreamSynchronize(d.s_in); drv_.cuStreamSynchronize(d.s_compute); drv_.cuStreamSynchronize(d.s_out); }
    *stage_code = e.empty() ? 0 : code;
    return e;
}

std::string Engine::backward_host(int dev, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                  const float* lse, void* dq, void* dk, void* dv, const AttnShape& s, int32_t dtype,
                                  float scale, bool causal, int* stage_code) {
    *stage_code = -1;
    if (!ready_) return "Library not initialized. Call aule_init() first.";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    *stage_code = -4;
    std::string e = validate(s, dtype);
    if (!e.empty()) return e;
    Device& d = *dp;
    std::lock_guard<std::mutex> host_guard(host_mu_[d.index]);
    CtxGuard g(drv_, d.ctx);
    const size_t es = dtype_size(dtype);
    const size_t qb = (size_t)s.B * s.Hq * s.Sq * s.D * es, kb = (size_t)s.B * s.Hkv * s.Sk * s.D * es;
    const size_t lb = (size_t)s.B * s.Hq * s.Sq * sizeof(float);
    const size_t sizes[9] = {qb, kb, kb, qb, lb, qb, qb, kb, kb};   // q k v o lse do dq dk dv
    *stage_code = -2;
    for (int i = 0; i < 9; ++i)
        if (!(e = ensure_stage(d, i, sizes[i])).empty()) return e;
    *stage_code = -3;
    const void* src[6] = {q, k, v, o, lse, d_o};
    for (int i = 0; i < 6; ++i)
        if (!(e = check(drv_.cuMemcpyHtoDAsync(d.stage[i], src[i], sizes[i], d.s_compute), "upload")).empty()) return e;
    *stage_code = -4;
    e = backward(dev, d.s_compute, d.stage[0], d.stage[1], d.stage[2], d.stage[3], d.stage[5], d.stage[4], d.stage[6],
                 d.stage[7], d.stage[8], s, dtype, scale, causal, -1);
    if (!e.empty()) return e;
    void* dst[3] = {dq, dk, dv};
    for (int i = 0; i < 3; ++i)
        if (!(e = check(drv_.cuMemcpyDtoHAsync(dst[i], d.stage[6 + i], sizes[6 + i], d.s_compute), "download")).empty()) {
            *stage_code = -5;
            return e;
        }
    if (!(e = check(drv_.cuStreamSynchronize(d.s_compute), "backward kernels")).empty()) return e;
    *stage_code = 0;
    return "";
}

std::string Engine::smoke_multiply(int dev, const float* in, float* out, uint32_t n) {
    if (!ready_) return "Library not initialized. Call aule_init() first.";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    Device& d = *dp;
    std::lock_guard<std::mutex> host_guard(host_mu_[d.index]);
    CtxGuard g(drv_, d.ctx);
    std::string e;
    if (!(e = ensure_stage(d, 0, (size_t)n * 4)).empty()) return e;
    if (!(e = ensure_stage(d, 3, (size_t)n * 4)).empty()) return e;
    if (!(e = check(drv_.cuMemcpyHtoDAsync(d.stage[0], in, (size_t)n * 4, d.s_compute), "upload")).empty()) return e;
    CUdeviceptr a = d.stage[0], b = d.stage[3];
    void* params[] = {&a, &b, &n};
    if (!(e = launch(d, d.smoke, "aule_smoke_multiply", (n + 255) / 256, 1, 1, 256, 0, d.s_compute, params)).empty()) return e;
    if (!(e = check(drv_.cuMemcpyDtoHAsync(out, b, (size_t)n * 4, d.s_compute), "download")).empty()) return e;
    return check(drv_.cuStreamSynchronize(d.s_compute), "smoke kernel");
}

std::string Engine::rope(int dev, CUstream stream, CUdeviceptr xq, CUdeviceptr oq, uint64_t bhq, uint32_t Sq, CUdeviceptr xk,
                         CUdeviceptr ok, uint64_t bhk, uint32_t Sk, CUdeviceptr cos, CUdeviceptr sin, uint32_t table_rows,
                         uint32_t D, int32_t mode, int32_t dtype, float sign) {
    if (!ready_) return "Library not initialized. Call aule_init() first.";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    if (dtype < 0 || dtype > 3) return "unsupported dtype (0=f32, 1=bf16, 2=f16, 3=f32 tensors with tf32 tensor-core math allowed)";
    if (dtype == kTF32) dtype = kF32;
    if (!xq || !oq || !cos || !sin) return "null device pointer";
    if (bhk && (!xk || !ok)) return "null device pointer";
    if (D == 0 || (D & 1) || !Sq || !bhq) return "RoPE needs an even head_dim and non-empty tensors";
    if (mode != 0 && mode != 1) return "RoPE convention must be 0 (half-split) or 1 (interleaved pairs)";
    const uint32_t need = std::max(Sq, bhk ? Sk : 0u);
    if (table_rows < need) {          // triton_flash.py:416-417 asserts the table shape; a short table would be read out of bounds
        char buf[160];
        snprintf(buf, sizeof(buf), "cos/sin tables have %u rows, the sequences need %u", table_rows, need);
        return buf;
    }
    const size_t es = dtype_size(dtype);
    const bool vec = (D % 8) == 0;
    if (vec && (((xq | oq | xk | ok) & (es * 4 - 1)) || ((cos | sin) & 15))) return "RoPE tensors must be 16-byte aligned";
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    aule_kp::RopeParams p;
    p.xq = (const void*)xq; p.oq = (void*)oq; p.rows_q = bhq * Sq; p.Sq = Sq;
    p.xk = (const void*)xk; p.ok = (void*)ok; p.rows_k = bhk ? bhk * Sk : 0; p.Sk = bhk ? Sk : 1;
    p.cs = (const float*)cos; p.sn = (const float*)sin; p.D = D; p.mode = mode; p.sign = sign;
    void* params[] = {&p};
    char name[32];
    snprintf(name, sizeof(name), "aule_rope_%s", kDtypeSuffix[dtype]);
    const uint64_t work = (p.rows_q + p.rows_k) * (vec ? D / 8 : D / 2);
    return launch(d, d.rope[dtype], name, (unsigned)std::min<uint64_t>((work + 255) / 256, (uint64_t)d.sm_count * 32), 1, 1, 256, 0, stream, params);
}

std::string Engine::forward_rope(int dev, CUstream stream, CUdeviceptr q, CUdeviceptr k, CUdeviceptr v, CUdeviceptr o,
                                 CUdeviceptr lse, CUdeviceptr cos, CUdeviceptr sin, uint32_t table_rows, int32_t mode,
                                 const AttnShape& s, int32_t dtype, float scale, bool causal, int32_t window) {
    if (!ready_) return "Library not initialized. Call aule_init() first.";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    std::string e = validate(s, dtype);
    if (!e.empty()) return e;
    if (!q || !k || !v || !o) return "null device pointer";
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    // K has to be rotated once per key, not once per (query block, key) pair, so the rotation is a prologue: ONE launch reads
    // Q and K once and writes the rotated copies once (stream-ordered workspace), then the fused kernel runs on them.
    const size_t es = dtype_size(dtype);
    const size_t qb = (size_t)s.B * s.Hq * s.Sq * s.D * es, kb = (size_t)s.B * s.Hkv * s.Sk * s.D * es;
    CUdeviceptr ws = 0;
    if (!(e = check(drv_.cuMemAllocAsync(&ws, qb + kb, stream), "cuMemAllocAsync(RoPE workspace)")).empty()) return e;
    e = rope(dev, stream, q, ws, (uint64_t)s.B * s.Hq, s.Sq, k, ws + qb, (uint64_t)s.B * s.Hkv, s.Sk, cos, sin, table_rows, s.D,
             mode, dtype, 1.f);
    if (e.empty()) e = forward(dev, stream, ws, ws + qb, v, o, lse, s, dtype, scale, causal, window);
    drv_.cuMemFreeAsync(ws, stream);
    return e;
}

// A single call whose tensors live on ONE device, computed by several (SURVEY 8e "spanning call"): the (batch, kv-head)
// units are split contiguously over `ndev` devices; every device other than the source receives its Q/K/V slabs over
// NVLink (cuMemcpyPeerAsync), runs the fused kernel on them and returns its O (and LSE) slab; the source device computes
// its own share in place.  Each peer works in `chunks` sub-ranges so that the copy-in of chunk c+1 and the copy-out of
// chunk c-1 overlap the kernel of chunk c (three streams per device).  Asynchronous for the caller: `stream` of the source
// device waits for the gathered result.  timings_ms (optional, makes the call synchronous):
// [0] scatter, [1] kernel, [2] gather = max over devices of each phase's device-timed duration, [3] total on the source stream.
std::string Engine::forward_spanning(int src_dev, CUstream stream, CUdeviceptr q, CUdeviceptr k, CUdeviceptr v, CUdeviceptr o,
                                     CUdeviceptr lse, const AttnShape& s, int32_t dtype, float scale, bool causal,
                                     int32_t window, const int32_t* devices, int32_t ndev, int32_t chunks, float* timings_ms) {
    if (!ready_) return "Library not initialized. Call aule_init() first.";
    std::string e = validate(s, dtype);
    if (!e.empty()) return e;
    if (!q || !k || !v || !o) return "null device pointer";
    if (ndev < 1 || ndev > 64 || !devices) return "spanning call needs 1..64 devices";
    Device* sp = by_ordinal(src_dev);
    if (!sp) return "invalid source device index";
    std::vector<Device*> devs;
    for (int i = 0; i < ndev; ++i) {
        Device* dp = by_ordinal(devices[i]);
        if (!dp) return "invalid device index in the spanning set";
        for (Device* x : devs) if (x == dp) return "duplicate device in the spanning set";
        devs.push_back(dp);
    }
    if (std::find(devs.begin(), devs.end(), sp) == devs.end()) return "the source device must be part of the spanning set";
    chunks = std::max(1, std::min<int32_t>(chunks, (int32_t)Device::kMaxChunks));
    const size_t es = dtype_size(dtype);
    const uint32_t group = s.Hq / s.Hkv, units = s.B * s.Hkv;
    const size_t q_unit = (size_t)group * s.Sq * s.D * es, kv_unit = (size_t)s.Sk * s.D * es, lse_unit = (size_t)group * s.Sq * sizeof(float);
    // contiguous unit ranges, as even as possible
    std::vector<uint32_t> u0(ndev + 1, 0);
    for (int i = 0; i < ndev; ++i) u0[i + 1] = u0[i] + units / ndev + ((uint32_t)i < units % ndev ? 1 : 0);
    // lock every participating device's staging state in index order (no lock-order inversion between concurrent callers)
    std::vector<Device*> order(devs);
    std::sort(order.begin(), order.end(), [](Device* a, Device* b) { return a->index < b->index; });
    std::vector<std::unique_lock<std::mutex>> locks;
    for (Device* dp : order) locks.emplace_back(host_mu_[dp->index]);

    struct Ev { CUevent ev[5] = {}; CUcontext ctx = nullptr; };   // in0, in1, k0, k1(=out0), out1 -- timing-enabled, per device
    std::vector<Ev> evs(ndev);
    auto cleanup = [&]() {
        for (int i = 0; i < ndev; ++i) {
            CtxGuard g(drv_, devs[i]->ctx);
            for (CUevent x : evs[i].ev) if (x) drv_.cuEventDestroy(x);
        }
    };
    CUevent ev_ready = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    {   // inputs are ready when the caller's stream reaches this point
        CtxGuard g(drv_, sp->ctx);
        e = check(drv_.cuEventCreate(&ev_ready, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
        if (e.empty()) e = check(drv_.cuEventCreate(&ev_t0, CU_EVENT_DEFAULT), "cuEventCreate");
        if (e.empty()) e = check(drv_.cuEventCreate(&ev_t1, CU_EVENT_DEFAULT), "cuEventCreate");
        if (e.empty()) e = check(drv_.cuEventRecord(ev_t0, stream), "cuEventRecord");
        if (e.empty()) e = check(drv_.cuEventRecord(ev_ready, stream), "cuEventRecord");
    }
    for (int i = 0; i < ndev && e.empty(); ++i) {
        Device& d = *devs[i];
        const uint32_t nu = u0[i + 1] - u0[i];
        if (nu == 0) continue;
        CtxGuard g(drv_, d.ctx);
        for (int x = 0; x < 5 && e.empty(); ++x) e = check(drv_.cuEventCreate(&evs[i].ev[x], CU_EVENT_DEFAULT), "cuEventCreate");
        if (!e.empty()) break;
        AttnShape cs{nu, group, 1, s.Sq, s.Sk, s.D};
        if (&d == sp) {                       // the source computes its own share in place, on its compute stream
            e = check(drv_.cuStreamWaitEvent(d.s_compute, ev_ready, 0), "cuStreamWaitEvent");
            if (e.empty()) e = check(drv_.cuEventRecord(evs[i].ev[2], d.s_compute), "cuEventRecord");
            if (e.empty()) e = forward(d.ordinal, d.s_compute, q + u0[i] * q_unit, k + u0[i] * kv_unit, v + u0[i] * kv_unit,
                                       o + u0[i] * q_unit, lse ? lse + u0[i] * lse_unit : 0, cs, dtype, scale, causal, window);
            if (e.empty()) e = check(drv_.cuEventRecord(evs[i].ev[3], d.s_compute), "cuEventRecord");
            continue;
        }
        if (!(e = ensure_stage(d, 0, q_unit * nu)).empty()) break;
        if (!(e = ensure_stage(d, 1, kv_unit * nu)).empty()) break;
        if (!(e = ensure_stage(d, 2, kv_unit * nu)).empty()) break;
        if (!(e = ensure_stage(d, 3, q_unit * nu)).empty()) break;
        if (lse && !(e = ensure_stage(d, 4, lse_unit * nu)).empty()) break;
        e = check(drv_.cuStreamWaitEvent(d.s_in, ev_ready, 0), "cuStreamWaitEvent");
        if (e.empty()) e = check(drv_.cuEventRecord(evs[i].ev[0], d.s_in), "cuEventRecord");
        // A chunk is a kernel launch of its own: it must still fill the GPU.  Keep at least ~4 work items per SM in a chunk
        // (config D at 8 GPUs: 4 heads per peer, 128 items of up to 256 blocks each per head -- one head per launch ran
        // 1.65 ms against 0.81 ms for the four heads in one launch), however many chunks the caller asked for.
        const uint64_t items_per_unit = (uint64_t)((s.Sq + 255) / 256) * group;
        const uint32_t by_work = (uint32_t)std::max<uint64_t>(1, (uint64_t)nu * items_per_unit / (4ull * (uint64_t)d.sm_count));
        const uint32_t nch = std::min<uint32_t>(std::min<uint32_t>((uint32_t)chunks, by_work), nu), per = (nu + nch - 1) / nch;
        for (uint32_t c = 0; c < nch && e.empty(); ++c) {
            const uint32_t a = c * per;
            if (a >= nu) break;
            const uint32_t n = std::min(per, nu - a), ga = u0[i] + a;      // local / global first unit of the chunk
            e = check(drv_.cuMemcpyPeerAsync(d.stage[0] + a * q_unit, d.ctx, q + ga * q_unit, sp->ctx, n * q_unit, d.s_in), "scatter Q");
            if (e.empty()) e = check(drv_.cuMemcpyPeerAsync(d.stage[1] + a * kv_unit, d.ctx, k + ga * kv_unit, sp->ctx, n * kv_unit, d.s_in), "scatter K");
            if (e.empty()) e = check(drv_.cuMemcpyPeerAsync(d.stage[2] + a * kv_unit, d.ctx, v + ga * kv_unit, sp->ctx, n * kv_unit, d.s_in), "scatter V");
            if (e.empty()) e = check(drv_.cuEventRecord(d.ev_in[c], d.s_in), "cuEventRecord");
            if (e.empty() && c + 1 == nch) e = check(drv_.cuEventRecord(evs[i].ev[1], d.s_in), "cuEventRecord");
            if (e.empty()) e = check(drv_.cuStreamWaitEvent(d.s_compute, d.ev_in[c], 0), "cuStreamWaitEvent");
            if (e.empty() && c == 0) e = check(drv_.cuEventRecord(evs[i].ev[2], d.s_compute), "cuEventRecord");
            AttnShape ccs{n, group, 1, s.Sq, s.Sk, s.D};
            if (e.empty()) e = forward(d.ordinal, d.s_compute, d.stage[0] + a * q_unit, d.stage[1] + a * kv_unit, d.stage[2] + a * kv_unit,
                                       d.stage[3] + a * q_unit, lse ? d.stage[4] + a * lse_unit : 0, ccs, dtype, scale, causal, window);
            if (e.empty()) e = check(drv_.cuEventRecord(d.ev_c[c], d.s_compute), "cuEventRecord");
            if (e.empty() && c + 1 == nch) e = check(drv_.cuEventRecord(evs[i].ev[3], d.s_compute), "cuEventRecord");
            if (e.empty()) e = check(drv_.cuStreamWaitEvent(d.s_out, d.ev_c[c], 0), "cuStreamWaitEvent");
            if (e.empty()) e = check(drv_.cuMemcpyPeerAsync(o + ga * q_unit, sp->ctx, d.stage[3] + a * q_unit, d.ctx, n * q_unit, d.s_out), "gather O");
            if (e.empty() && lse) e = check(drv_.cuMemcpyPeerAsync(lse + ga * lse_unit, sp->ctx, d.stage[4] + a * lse_unit, d.ctx, n * lse_unit, d.s_out), "gather LSE");
        }
        if (e.empty()) e = check(drv_.cuEventRecord(evs[i].ev[4], d.s_out), "cuEventRecord");
    }
    // the caller's stream continues when every device's result has landed on the source
    if (e.empty()) {
        CtxGuard g(drv_, sp->ctx);
        for (int i = 0; i < ndev && e.empty(); ++i) {
            if (u0[i + 1] == u0[i]) continue;
            e = check(drv_.cuStreamWaitEvent(stream, devs[i] == sp ? evs[i].ev[3] : evs[i].ev[4], 0), "cuStreamWaitEvent");
        }
        if (e.empty()) e = check(drv_.cuEventRecord(ev_t1, stream), "cuEventRecord");
    }
    // The per-call events and the staging buffers are owned until the work has drained; timing needs it anyway.
    {
        CtxGuard g(drv_, sp->ctx);
        CUresult r = drv_.cuEventSynchronize(ev_t1);
        if (e.empty()) e = check(r, "spanning call");
        for (Device* dp : devs) { CtxGuard g2(drv_, dp->ctx); drv_.cuStreamSynchronize(dp->s_in); drv_.cuStreamSynchronize(dp->s_compute); drv_.cuStreamSynchronize(dp->s_out); }
    }
    if (e.empty() && timings_ms) {
        float sc = 0.f, kn = 0.f, ga = 0.f, tot = 0.f;
        for (int i = 0; i < ndev; ++i) {
            if (u0[i + 1] == u0[i]) continue;
            CtxGuard g(drv_, devs[i]->ctx);
            float ms = 0.f;
            if (devs[i] != sp) {
                if (drv_.cuEventElapsedTime(&ms, evs[i].ev[0], evs[i].ev[1]) == CUDA_SUCCESS) sc = std::max(sc, ms);
                if (drv_.cuEventElapsedTime(&ms, evs[i].ev[3], evs[i].ev[4]) == CUDA_SUCCESS) ga = std::max(ga, ms);
            }
            if (drv_.cuEventElapsedTime(&ms, evs[i].ev[2], evs[i].ev[3]) == CUDA_SUCCESS) kn = std::max(kn, ms);
        }
        {
            CtxGuard g(drv_, sp->ctx);
            drv_.cuEventElapsedTime(&tot, ev_t0, ev_t1);
        }
        timings_ms[0] = sc; timings_ms[1] = kn; timings_ms[2] = ga; timings_ms[3] = tot;
    }
    cleanup();
    {
        CtxGuard g(drv_, sp->ctx);
        if (ev_ready) drv_.cuEventDestroy(ev_ready);
        if (ev_t0) drv_.cuEventDestroy(ev_t0);
        if (ev_t1) drv_.cuEventDestroy(ev_t1);
    }
    return e;
}

std::string Engine::make_tmap_paged(CUtensorMap* m, int32_t dtype, CUdeviceptr base, uint32_t num_blocks,
                                    uint32_t block_size, uint32_t Hkv, uint32_t D) const {
    // [num_blocks, block_size, Hkv, D] 16-bit cache viewed as a 5-D tensor (64 | D/64 | Hkv | block_size | num_blocks):
    // splitting D into 64-element halves keeps the innermost box at 128 bytes (the 128-byte swizzle's limit) while a
    // token's D*2 bytes still arrive as one contiguous request.  Box = 16 tokens of one head of one page; a page index
    // outside [0, num_blocks) is out of bounds and reads as zeros.
    const cuuint64_t row = (cuuint64_t)D * 2;
    cuuint64_t dims[5] = {64, D / 64, Hkv, block_size, num_blocks};
    cuuint64_t strides[4] = {128, row, row * Hkv, row * Hkv * block_size};
    cuuint32_t box[5] = {64, D / 64, 1, 16, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = drv_.cuTensorMapEncodeTiled(
        m, dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)base, dims,
        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return check(r, "cuTensorMapEncodeTiled(paged cache)");
}

std::string Engine::paged_decode(int dev, CUstream stream, CUdeviceptr q, CUdeviceptr k_cache, CUdeviceptr v_cache,
                                 CUdeviceptr block_tables, CUdeviceptr context_lens, CUdeviceptr out, uint32_t B,
                                 uint32_t Hq, uint32_t Hkv, uint32_t D, uint32_t num_blocks, uint32_t block_size,
                                 uint32_t max_blocks, uint32_t max_context, int32_t dtype, float scale, int32_t window) {
    if (!ready_) return "Library not initialized. Call aule_init() first.";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    char buf[256];
    if (dtype != kBF16 && dtype != kF16) return "paged decode needs a 16-bit cache (1=bf16, 2=f16)";
    if (!B || !Hq || !Hkv || !num_blocks || !block_size || !max_blocks) return "empty tensor dimension";
    if (Hq % Hkv != 0) {   // triton_flash_amd.py:696-697
        snprintf(buf, sizeof(buf), "heads_q (%u) must be divisible by heads_kv (%u)", Hq, Hkv);
        return buf;
    }
    if (Hq / Hkv > 16) return "paged decode supports at most 16 query heads per kv head";
    if (D != 64 && D != 128) return "paged decode needs head_dim 64 or 128";
    if (block_size % 16 != 0 || block_size > 256) return "paged decode needs block_size to be a multiple of 16, at most 256";
    if (!q || !k_cache || !v_cache || !block_tables || !context_lens || !out) return "null device pointer";
    if ((k_cache | v_cache) & 15) return "k_cache / v_cache must be 16-byte aligned";
    if (q & 3) return "q must be 4-byte aligned";
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    if (!(scale > 0.f)) scale = 1.0f / sqrtf((float)D);   // triton_flash_amd.py:699-700
    if (window == 0) window = -1;

    CUtensorMap tmK, tmV;
    std::string e = make_tmap_paged(&tmK, dtype, k_cache, num_blocks, block_size, Hkv, D);
    if (e.empty()) e = make_tmap_paged(&tmV, dtype, v_cache, num_blocks, block_size, Hkv, D);
    if (!e.empty()) return e;

    // Split count: fill the resident CTA slots (2 per SM) in one wave, at least 4 tiles (64 tokens) per split.
    uint64_t span = max_context ? max_context : (uint64_t)max_blocks * block_size;
    if (window > 0) span = std::min<uint64_t>(span, (uint64_t)window + 15);
    const uint64_t tiles = (span + 15) / 16;
    const uint64_t base = (uint64_t)B * Hkv, slots = 2ull * d.sm_count;
    uint64_t nsplit = std::max<uint64_t>(1, slots / base);
    nsplit = std::min<uint64_t>(nsplit, std::max<uint64_t>(1, tiles / 4));
    nsplit = std::min<uint64_t>(nsplit, 128);
    if (const char* ov = getenv("AULE_PAGED_NSPLIT")) {   // tuning hook (tools/bench_paged.py sweeps it)
        const long v = atol(ov);
        if (v > 0) nsplit = std::min<uint64_t>((uint64_t)v, std::max<uint64_t>(1, tiles));
    }
    if (base * nsplit > 0x7fffffffull) return "problem too large (paged decode grid exceeds 2^31 CTAs)";

    aule_kp::PagedParams p;
    memset(&p, 0, sizeof(p));
    p.q = (const void*)q; p.out = (void*)out;
    p.block_tables = (const int32_t*)block_tables; p.context_lens = (const int32_t*)context_lens;
    p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.block_size = block_size; p.max_blocks = max_blocks;
    p.nsplit = (uint32_t)nsplit; p.scale_log2 = scale * 1.4426950408889634f; p.window = window;
    CUdeviceptr ws = 0;
    if (nsplit > 1) {
        const size_t rows = (size_t)B * Hq * nsplit;
        const size_t o_bytes = rows * D * sizeof(float);
        if (!(e = check(drv_.cuMemAllocAsync(&ws, o_bytes + rows * 2 * sizeof(float), stream), "cuMemAllocAsync(split workspace)")).empty()) return e;
        p.ws_o = (float*)ws;
        p.ws_ml = (float*)(ws + o_bytes);
    }
    const bool d128 = D == 128;
    char name[64];
    {
        void* params[] = {&tmK, &tmV, &p};
        snprintf(name, sizeof(name), "aule_paged_sm100_%s_d%u", kDtypeSuffix[dtype], D);
        e = launch(d, d.paged[dtype][d128 ? 1 : 0], name, (unsigned)(base * nsplit), 1, 1,
                   (unsigned)aule_kp::PagedCfg<128>::THREADS,
                   d128 ? aule_kp::PagedCfg<128>::SMEM_BYTES : aule_kp::PagedCfg<64>::SMEM_BYTES, stream, params);
    }
    if (e.empty() && nsplit > 1) {
        void* params[] = {&p, &D};
        snprintf(name, sizeof(name), "aule_paged_combine_%s", kDtypeSuffix[dtype]);
        e = launch(d, d.paged_combine[dtype], name, B * Hq, 1, 1, D, 0, stream, params);
    }
    if (ws) drv_.cuMemFreeAsync(ws, stream);
    return e;
}

std::string Engine::mem_alloc(int dev, size_t bytes, CUdeviceptr* out) {
    if (!ready_) return "Not initialized";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    return check(drv_.cuMemAlloc(out, bytes ? bytes : 4), "cuMemAlloc");
}
void Engine::mem_free(int dev, CUdeviceptr p) {
    Device* dp = ready_ ? by_ordinal(dev) : nullptr;
    if (!dp || !p) return;
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    drv_.cuMemFree(p);
}
std::string Engine::copy_h2d(int dev, CUdeviceptr dst, const void* src, size_t bytes) {
    Device* dp = ready_ ? by_ordinal(dev) : nullptr;
    if (!dp) return "invalid device index";
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    return check(drv_.cuMemcpyHtoD(dst, src, bytes), "cuMemcpyHtoD");
}
std::string Engine::copy_d2h(int dev, void* dst, CUdeviceptr src, size_t bytes) {
    Device* dp = ready_ ? by_ordinal(dev) : nullptr;
    if (!dp) return "invalid device index";
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    return check(drv_.cuMemcpyDtoH(dst, src, bytes), "cuMemcpyDtoH");
}
std::string Engine::synchronize(int dev) {
    if (!ready_) return "Not initialized";
    Device* dp = by_ordinal(dev);
    if (!dp) return "invalid device index";
    Device& d = *dp;
    CtxGuard g(drv_, d.ctx);
    return check(drv_.cuCtxSynchronize(), "cuCtxSynchronize");
}

}  // namespace aule
