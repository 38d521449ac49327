// Attention engine over the CUDA Driver API: per-device primary context, the embedded
// sm_100a module, TMA descriptor construction and kernel launches.
//
// Role-for-role replacement of the reference's device runtime and pipelines:
//   Engine::init / Device            <- src/vulkan_context.zig:52-251 (instance/device/queue),
//                                       src/backends/hip.zig:133-160 (module load)
//   Engine::forward / backward       <- src/attention_gpu.zig:360-453,:707-830 (AttentionEngine),
//                                       src/attention_pipeline.zig:312-391 (dispatch),
//                                       src/attention_backward_pipeline.zig:228-257,:490-519
//   DeviceBuffer / staging           <- src/buffer_manager.zig:39-78, src/gpu_tensor.zig:32-94
//   Engine::smoke_multiply           <- src/compute_pipeline.zig:203-254 + shaders/test.comp
// Unlike the reference (one blocking vkQueueSubmit + fence per call,
// attention_pipeline.zig:377-390) the device-pointer entry points are asynchronous on the
// caller's stream.
#pragma once
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "cuda_driver.h"

namespace aule {

// kTF32: fp32 tensors in memory, forward allowed to run on the tensor cores as tf32 (head_dim <= 64; otherwise, and in the
// backward, it behaves exactly like kF32)
enum DType : int32_t { kF32 = 0, kBF16 = 1, kF16 = 2, kTF32 = 3 };
inline size_t dtype_size(int32_t dt) { return (dt == kF32 || dt == kTF32) ? 4 : 2; }

struct AttnShape {
    uint32_t B, Hq, Hkv, Sq, Sk, D;
};

enum KernelPath : int32_t { kAuto = 0, kForceCudaCore = 1, kVariantBase = 16 /* 16+v: bf16 d128 tuning variant v */ };

struct Device {
    int ordinal = -1;
    CUdevice dev = 0;
    CUcontext ctx = nullptr;
    CUmodule mod = nullptr;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    char name[256] = {0};
    // kernels
    CUfunction fwd_simt[3] = {nullptr, nullptr, nullptr};      // indexed by DType
    CUfunction bwd_dq_simt[3] = {nullptr, nullptr, nullptr};
    CUfunction bwd_dkv_simt[3] = {nullptr, nullptr, nullptr};
    CUfunction fwd_sm100[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   // [dtype][D==128]
    // present only in tuning builds (make EXTRA_NVFLAGS=-DAULE_TUNING_VARIANTS): bf16 variants [D==128][v] and the v4 kernel
    CUfunction fwd_sm100_var[2][16] = {};
    CUfunction fwd4_sm100[2] = {nullptr, nullptr};
    CUfunction bwd_sm100[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   // [dtype][D==128]
    CUfunction bwd_dkvt_sm100[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   // transposed dK/dV kernel (v4)
    CUfunction bwd_delta[3] = {nullptr, nullptr, nullptr};
    CUfunction bwd_fused2_sm100[3] = {nullptr, nullptr, nullptr};
    CUfunction bwd_dq_sm100[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   // dQ kernel [dtype][D==128]
    CUfunction bwd_dq1_sm100[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};  // superseded v1 dQ kernel (tuning builds)
    CUfunction fwd_tf32 = nullptr;                                  // tf32 forward for fp32 tensors (head_dim <= 64)
    CUfunction bwd_fused_sm100[3] = {nullptr, nullptr, nullptr};   // fused backward (D = 128) [dtype]
    CUfunction bwd_dq_convert[3] = {nullptr, nullptr, nullptr};    // its fp32 dQ accumulator -> 16-bit
    CUfunction rope[3] = {nullptr, nullptr, nullptr};
    CUfunction paged[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};     // decode kernel [dtype][D==128]
    CUfunction paged_combine[3] = {nullptr, nullptr, nullptr};
    CUfunction smoke = nullptr;
    // streams for the host-staged entry points
    CUstream s_in = nullptr, s_compute = nullptr, s_out = nullptr;
    // backward: the dQ kernel runs on s_aux beside the dK/dV kernel on the caller's stream (fork / join events)
    CUstream s_aux = nullptr;
    CUevent ev_fork = nullptr, ev_join = nullptr;
    int index = -1;                  // position in Engine::devices_ (selects the per-device mutexes)
    // per-launch work counters of the forward kernel's dynamic scheduler: a ring of kSchedSlots x u32, each zeroed on the
    // caller's stream before its launch.  A slot is handed out again only after the launch that last used it has
    // completed (sched_ev, recorded on that launch's stream), whatever stream the next caller is on.
    static constexpr uint32_t kSchedSlots = 256;
    CUdeviceptr sched = 0;
    uint32_t sched_next = 0;
    CUevent sched_ev[kSchedSlots] = {};
    bool sched_ev_live[kSchedSlots] = {};
    // persistent events of the pipelined host-buffer entries (created once)
    static constexpr uint32_t kMaxChunks = 32;
    CUevent ev_in[kMaxChunks] = {}, ev_c[kMaxChunks] = {};
    // pinned bounce buffers for pageable callers: [0,1] upload double buffer, [2,3] download double buffer
    void* bounce[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t bounce_cap[4] = {0, 0, 0, 0};
    CUevent bounce_ev[4] = {};
    // grow-only staging buffers for host-pointer calls: q k v o lse do dq dk dv
    CUdeviceptr stage[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    size_t stage_cap[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

class Engine {
public:
    // "" on success. Idempotent.
    std::string init();
    void shutdown();
    bool ready() const { return ready_; }
    int device_count() const { return (int)devices_.size(); }
    // Devices are addressed by CUDA ordinal everywhere (== torch device index).
    Device* by_ordinal(int ordinal);
    const Device* first() const { return devices_.empty() ? nullptr : &devices_[0]; }
    CudaDriver& driver() { return drv_; }

    // Validation shared by every entry (mirrors attention_gpu.zig:372-404 widened to GQA,
    // D <= 128, Sq != Sk). Returns "" or the reason.
    static std::string validate(const AttnShape& s, int32_t dtype);

    // Asynchronous on `stream` of device `dev`. Device pointers. Returns "" or an error.
    std::string forward(int dev, CUstream stream, CUdeviceptr q, CUdeviceptr k, CUdeviceptr v, CUdeviceptr o,
                        CUdeviceptr lse, const AttnShape& s, int32_t dtype, float scale, bool causal, int32_t window);
    std::string backward(int dev, CUstream stream, CUdeviceptr q, CUdeviceptr k, CUdeviceptr v, CUdeviceptr o,
                         CUdeviceptr d_o, CUdeviceptr lse, CUdeviceptr dq, CUdeviceptr dk, CUdeviceptr dv,
                         const AttnShape& s, int32_t dtype, float scale, bool causal, int32_t window = -1);
    // Synchronous, host pointers, chunked + pipelined over (batch x kv-head) units.
    // `stage_code` receives the reference's failure stage (-2 alloc, -3 upload, -4 compute, -5 download).
    std::string forward_host(int dev, const void* q, const void* k, const void* v, void* o, float* lse,
                             const AttnShape& s, int32_t dtype, float scale, bool causal, int32_t window,
                             int* stage_code);
    std::string backward_host(int dev, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                              const float* lse, void* dq, void* dk, void* dv, const AttnShape& s, int32_t dtype,
                              float scale, bool causal, int* stage_code);
    std::string smoke_multiply(int dev, const float* in, float* out, uint32_t n);
    // RoPE of one or two tensors in one launch: x*/o*: [bh*, S*, D] of `dtype` (xk = 0: one tensor), cos/sin: [table_rows, D/2]
    // fp32; mode 0 = half-split (triton_flash.py:680-703), 1 = interleaved pairs (attention_f32.comp:98-111); sign -1 = inverse.
    std::string rope(int dev, CUstream stream, CUdeviceptr xq, CUdeviceptr oq, uint64_t bhq, uint32_t Sq, CUdeviceptr xk,
                     CUdeviceptr ok, uint64_t bhk, uint32_t Sk, CUdeviceptr cos, CUdeviceptr sin, uint32_t table_rows,
                     uint32_t D, int32_t mode, int32_t dtype, float sign);
    // RoPE prologue (one launch over Q and K into a stream-ordered workspace) + fused attention.
    std::string forward_rope(int dev, CUstream stream, CUdeviceptr q, CUdeviceptr k, CUdeviceptr v, CUdeviceptr o,
                             CUdeviceptr lse, CUdeviceptr cos, CUdeviceptr sin, uint32_t table_rows, int32_t mode,
                             const AttnShape& s, int32_t dtype, float scale, bool causal, int32_t window);
    // One call spanning several devices (inputs resident on src_dev): NVLink scatter of Q/K/V slabs, kernels, gather of O/LSE.
    std::string forward_spanning(int src_dev, CUstream stream, CUdeviceptr q, CUdeviceptr k, CUdeviceptr v, CUdeviceptr o,
                                 CUdeviceptr lse, const AttnShape& s, int32_t dtype, float scale, bool causal, int32_t window,
                                 const int32_t* devices, int32_t ndev, int32_t chunks, float* timings_ms);

    // Paged-KV decode (one query token per sequence; python/aule/triton_flash_amd.py:662-740): q/out [B,Hq,D],
    // k_cache/v_cache [num_blocks, block_size, Hkv, D], block_tables [B, max_blocks] i32, context_lens [B] i32.
    // max_context (an upper bound of context_lens, 0 = max_blocks * block_size) only sizes the split count.
    std::string paged_decode(int dev, CUstream stream, CUdeviceptr q, CUdeviceptr k_cache, CUdeviceptr v_cache,
                             CUdeviceptr block_tables, CUdeviceptr context_lens, CUdeviceptr out, uint32_t B,
                             uint32_t Hq, uint32_t Hkv, uint32_t D, uint32_t num_blocks, uint32_t block_size,
                             uint32_t max_blocks, uint32_t max_context, int32_t dtype, float scale, int32_t window);

    // Raw memory for the handle-table tensors (device 0).
    std::string mem_alloc(int dev, size_t bytes, CUdeviceptr* out);
    void mem_free(int dev, CUdeviceptr p);
    std::string copy_h2d(int dev, CUdeviceptr dst, const void* src, size_t bytes);
    std::string copy_d2h(int dev, void* dst, CUdeviceptr src, size_t bytes);
    std::string synchronize(int dev);

    // path (test / A-B hooks; 0 = the shipped configuration):
    //   bits 0-7   0 auto, 1 force the CUDA-core kernels, 16+v forward tuning variant v (bf16)
    //   bit 8      forward: no head pairing                     bit 9   forward: no L2-residency runs of the work-item order
    //   bits 10-11 backward timing hook: 1 = dK/dV kernel only, 2 = dQ kernel only (partial gradients, never for real use)
    //   bit 12     backward bring-up: the issuer waits for every MMA group (tools/bwd_trace.py serial)
    //   bit 13     backward: the v3 dK/dV kernel (P, dS staged through shared memory) instead of the transposed v4
    //   bit 14     forward: the v4 kernel (tuning builds only; an error otherwise)
    //   bit 16     backward: dK/dV and dQ kernels back to back on the caller's stream (no second stream)
    //   bit 15     forward v4: no cross-item prefetch of the next work item's first Q K^T
    //   bit 17     backward: fused kernel (dQ reduced into an fp32 accumulator; D = 128 only) -- see bwd_fused_ below
    //   bits 18-22 backward, fused kernel: (batch, kv-head) units per CTA run (0 = default)
    //   bit 23     backward: the v1 dQ kernel (tuning builds only; an error otherwise)
    //   bit 24     backward dK/dV kernel: try_wait side polls in the issuer's load pump (the pre-round-2 behaviour; A/B)
    //   bit 25     backward dK/dV kernel, head_dim <= 64: one S^T buffer instead of two (the earlier behaviour; A/B)
    //   bit 27     backward, fused kernel of bit 26: proxy fence on the consumer side (A/B)
    //   bit 28     backward, fused kernel of bit 26: proxy fence of the dS^T tile by a helper warp (A/B)
    //   bit 26     backward: fused kernel in 64-query half steps with TMA bulk reductions (attn_bwd_fused2_sm100.cu; D = 128)
    void set_kernel_path(int32_t p) {
        bwd_fused_ = (p >> 17) & 1; bwd_units_per_run_ = (p >> 18) & 31; bwd_dq_v1_ = (p >> 23) & 1; bwd_legacy_poll_ = (p >> 24) & 1; bwd_single_s_ = (p >> 25) & 1; bwd_fused2_ = (p >> 26) & 1; bwd_consumer_fence_ = (p >> 27) & 1; bwd_fencer_ = (p >> 28) & 1;
        pair_heads_enabled_ = !(p & 256); l2_runs_enabled_ = !(p & 512); cross_item_enabled_ = !(p & 32768); bwd_order_ = (p >> 10) & 3; bwd_serial_ = (p >> 12) & 3; fwd_v4_ = (p >> 14) & 1; bwd_two_streams_ = !(p & 65536); path_ = p & 255;
    }
    void set_trace_buffer(uint64_t dptr) { trace_ = dptr; }
    static bool log_enabled();     // AULE_LOG=1: report kernel-path choices (CUDA-core fallbacks) on stderr
    uint64_t launch_count() const { return launches_.load(); }
    std::string last_kernel() { std::lock_guard<std::mutex> g(name_mu_); return last_kernel_; }

private:
    std::string check(CUresult r, const char* what) const;
    std::string load_device(int ordinal);
    std::string ensure_stage(Device& d, int slot, size_t bytes);
    std::string ensure_bounce(Device& d, int slot, size_t bytes);
    bool host_pinned(const void* p) const;
    std::string launch(Device& d, CUfunction fn, const char* name, unsigned gx, unsigned gy, unsigned gz, unsigned bx,
                       unsigned smem, CUstream stream, void** params);
    std::string make_tmap(CUtensorMap* m, int32_t dtype, CUdeviceptr base, uint64_t bh, uint64_t S, uint32_t D,
                          bool mn_major_f32 = false, uint32_t box_rows = 128) const;
    std::string make_tmap_paged(CUtensorMap* m, int32_t dtype, CUdeviceptr base, uint32_t num_blocks, uint32_t block_size,
                                uint32_t Hkv, uint32_t D) const;

    CudaDriver drv_;
    std::vector<Device> devices_;
    bool ready_ = false;
    int32_t path_ = kAuto;
    bool pair_heads_enabled_ = true;
    bool l2_runs_enabled_ = true;
    bool cross_item_enabled_ = true;   // forward: next item's first Q K^T under the current item's last block (path bit 15 disables)
    int32_t bwd_order_ = 0;
    int32_t fwd_v4_ = 0;
    bool bwd_two_streams_ = true;
    int32_t bwd_fused_ = 0;
    int32_t bwd_dq_v1_ = 0;
    int32_t bwd_legacy_poll_ = 0;
    int32_t bwd_single_s_ = 0;
    int32_t bwd_fused2_ = 0;
    int32_t bwd_consumer_fence_ = 0;
    int32_t bwd_fencer_ = 0;
    int32_t bwd_units_per_run_ = 0;
    int32_t bwd_serial_ = 0;      // BwdParams::order. bit 0 (path bit 12), bring-up: the issuer waits for every MMA group
                                  // (tools/bwd_trace.py serial); bits 1-2 (path bits 13-14): polynomial-exp2 pairs of 4 (A/B)
    uint64_t trace_ = 0;
    std::atomic<uint64_t> launches_{0};
    std::string last_kernel_ = "none";
    // Thread safety (ctypes releases the GIL during calls): the device-pointer entries only share the scheduler-counter
    // ring (launch_mu_); the host-pointer entries share the staging buffers and the three internal streams and are
    // serialised per device (host_mu_).
    static constexpr int kMaxDevices = 64;
    std::mutex launch_mu_[kMaxDevices], host_mu_[kMaxDevices], name_mu_;
};

// RAII: make a device's primary context current for the duration of a call.
class CtxGuard {
public:
    CtxGuard(CudaDriver& drv, CUcontext ctx);
    ~CtxGuard();
private:
    CudaDriver& drv_;
    bool pushed_ = false;
};

}  // namespace aule
