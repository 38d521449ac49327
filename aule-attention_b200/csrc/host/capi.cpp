// libaule C ABI (include/aule.h).  Part 1 mirrors src/lib.zig export for export: global
// engine, 1024-slot tensor handle table (lib.zig:14-17), static error buffer (:19-26),
// the same return codes.  Part 2 is the device-pointer / dtype / scale / multi-GPU extension.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../../include/aule.h"
#include "engine.h"

namespace {

aule::Engine g_engine;

constexpr uint32_t kMaxTensors = 1024;      // lib.zig:16
struct TensorSlot {
    bool live = false;
    uint32_t shape[4] = {0, 0, 0, 0};
    uint32_t count = 0;                      // elements (32-bit each)
    CUdeviceptr ptr = 0;
};
TensorSlot g_tensors[kMaxTensors];

// lib.zig:20 keeps one static buffer; here one per calling thread (ctypes releases the GIL, so two Python threads can fail
// at the same time): the pointer aule_get_error() returns stays valid until the next failing call ON THAT THREAD.
thread_local char g_error[512];
thread_local size_t g_error_len = 0;
thread_local char g_name_buf[300];

void set_error(const char* fmt, ...) {       // lib.zig:23-25
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    g_error_len = n < 0 ? 0 : ((size_t)n >= sizeof(g_error) ? sizeof(g_error) - 1 : (size_t)n);
}

TensorSlot* slot_of(uint64_t handle) {
    if (handle == 0 || handle > kMaxTensors) return nullptr;
    TensorSlot* s = &g_tensors[handle - 1];
    return s->live ? s : nullptr;
}

int primary_device() {                        // handle-table tensors live on the first usable device
    const aule::Device* d = g_engine.first();
    return d ? d->ordinal : -1;
}

}  // namespace

extern "C" {

int32_t aule_init(void) {
    if (g_engine.ready()) return 0;           // lib.zig:66-69: idempotent
    std::string e = g_engine.init();
    if (!e.empty()) {
        set_error("Failed to initialize CUDA backend: %s", e.c_str());
        return -1;
    }
    return 0;
}

void aule_shutdown(void) {
    aule_tensor_clear_all();
    g_engine.shutdown();
}

const char* aule_get_error(void) {
    if (g_error_len == 0) return "No error";
    g_error[g_error_len] = 0;
    return g_error;
}

const char* aule_get_backend_name(void) { return g_engine.ready() ? "cuda-sm100" : "Not initialized"; }
int32_t aule_supports_backward(void) { return g_engine.ready() ? 1 : 0; }
int32_t aule_get_vendor(void) { return g_engine.ready() ? 2 : -1; }          // 2 = nvidia
int32_t aule_get_gpu_vendor(void) { return aule_get_vendor(); }
int32_t aule_is_amd_optimized(void) { return 0; }
int32_t aule_has_fp16(void) { return g_engine.ready() ? 1 : 0; }
int32_t aule_get_subgroup_size(void) { return g_engine.ready() ? 32 : 0; }
int32_t aule_set_shader_variant(uint8_t variant) {
    if (!g_engine.ready()) return -1;
    if (variant != 0) { set_error("Shader variant %u not available (single sm100 variant)", (unsigned)variant); return -2; }
    return 0;
}
int32_t aule_get_shader_variant(void) { return g_engine.ready() ? 0 : -1; }
int32_t aule_has_shader_variant(uint8_t variant) { return (g_engine.ready() && variant == 0) ? 1 : 0; }

int32_t aule_get_device_name(char* buffer, uint32_t buffer_len) {
    const aule::Device* d = g_engine.first();
    if (!g_engine.ready() || !d || !buffer || buffer_len == 0) return -1;
    size_t n = strlen(d->name);
    if (n >= buffer_len) n = buffer_len - 1;
    memcpy(buffer, d->name, n);
    buffer[n] = 0;
    return (int32_t)n;
}

// ---------------------------------------------------------------- host-pointer fp32 entries
int32_t aule_attention_forward(const float* query, const float* key, const float* value, float* output,
                               uint32_t batch_size, uint32_t num_heads, uint32_t seq_len, uint32_t head_dim,
                               int32_t causal) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{batch_size, num_heads, num_heads, seq_len, seq_len, head_dim};
    int code = 0;
    std::string e = g_engine.forward_host(primary_device(), query, key, value, output, nullptr, s, aule::kF32, 0.f,
                                          causal != 0, -1, &code);
    if (!e.empty()) { set_error("Attention failed: %s", e.c_str()); return code ? code : -4; }
    return 0;
}

int32_t aule_attention_forward_with_lse(const float* query, const float* key, const float* value, float* output,
                                        float* lse_out, uint32_t batch_size, uint32_t num_heads, uint32_t seq_len,
                                        uint32_t head_dim, int32_t causal) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{batch_size, num_heads, num_heads, seq_len, seq_len, head_dim};
    int code = 0;
    std::string e = g_engine.forward_host(primary_device(), query, key, value, output, lse_out, s, aule::kF32, 0.f,
                                          causal != 0, -1, &code);
    if (!e.empty()) { set_error("Forward with LSE failed: %s", e.c_str()); return code ? code : -4; }
    return 0;
}

int32_t aule_attention_backward(const float* query, const float* key, const float* value, const float* output,
                                const float* grad_output, const float* lse, float* grad_query, float* grad_key,
                                float* grad_value, uint32_t batch_size, uint32_t num_heads, uint32_t seq_len,
                                uint32_t head_dim, int32_t causal) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{batch_size, num_heads, num_heads, seq_len, seq_len, head_dim};
    int code = 0;
    std::string e = g_engine.backward_host(primary_device(), query, key, value, output, grad_output, lse, grad_query,
                                           grad_key, grad_value, s, aule::kF32, 0.f, causal != 0, &code);
    if (!e.empty()) { set_error("Backward pass failed: %s", e.c_str()); return code ? code : -4; }
    return 0;
}

// ---------------------------------------------------------------- handle-table tensors
uint64_t aule_tensor_create(uint32_t batch_size, uint32_t num_heads, uint32_t seq_len, uint32_t head_dim) {
    if (!g_engine.ready()) { set_error("Not initialized"); return 0; }
    uint32_t idx = kMaxTensors;
    for (uint32_t i = 0; i < kMaxTensors; ++i)
        if (!g_tensors[i].live) { idx = i; break; }
    if (idx == kMaxTensors) { set_error("Max tensors reached"); return 0; }
    const uint64_t count = (uint64_t)batch_size * num_heads * seq_len * head_dim;
    if (count == 0 || count > 0xffffffffull) { set_error("Create tensor failed: element count %llu out of range", (unsigned long long)count); return 0; }
    CUdeviceptr p = 0;
    std::string e = g_engine.mem_alloc(primary_device(), count * 4, &p);
    if (!e.empty()) { set_error("Create tensor failed: %s", e.c_str()); return 0; }
    TensorSlot& t = g_tensors[idx];
    t.live = true;
    t.shape[0] = batch_size; t.shape[1] = num_heads; t.shape[2] = seq_len; t.shape[3] = head_dim;
    t.count = (uint32_t)count;
    t.ptr = p;
    return (uint64_t)idx + 1;
}

uint64_t aule_tensor_create_u32(uint32_t b, uint32_t h, uint32_t s, uint32_t d) { return aule_tensor_create(b, h, s, d); }

void aule_tensor_destroy(uint64_t handle) {
    TensorSlot* t = slot_of(handle);
    if (!t) return;
    g_engine.mem_free(primary_device(), t->ptr);
    *t = TensorSlot{};
}

int32_t aule_tensor_upload(uint64_t handle, const float* data, uint32_t count) {
    if (!g_engine.ready()) return -1;
    TensorSlot* t = slot_of(handle);
    if (!t) return -1;
    if (count != t->count) { set_error("Upload failed: size mismatch (tensor has %u elements, got %u)", t->count, count); return -3; }
    std::string e = g_engine.copy_h2d(primary_device(), t->ptr, data, (size_t)count * 4);
    if (!e.empty()) { set_error("Upload failed: %s", e.c_str()); return -3; }
    return 0;
}

int32_t aule_tensor_download(uint64_t handle, float* output, uint32_t count) {
    if (!g_engine.ready()) return -1;
    TensorSlot* t = slot_of(handle);
    if (!t) return -1;
    if (count != t->count) { set_error("Download failed: size mismatch (tensor has %u elements, got %u)", t->count, count); return -3; }
    std::string e = g_engine.copy_d2h(primary_device(), output, t->ptr, (size_t)count * 4);
    if (!e.empty()) { set_error("Download failed: %s", e.c_str()); return -3; }
    return 0;
}

int32_t aule_tensor_download_u32(uint64_t handle, uint32_t* output, uint32_t count) {
    return aule_tensor_download(handle, reinterpret_cast<float*>(output), count);
}

uint32_t aule_tensor_size(uint64_t handle) {
    TensorSlot* t = slot_of(handle);
    return t ? t->count : 0;
}

uint32_t aule_tensor_count(void) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < kMaxTensors; ++i) n += g_tensors[i].live ? 1 : 0;
    return n;
}
uint32_t aule_tensor_max(void) { return kMaxTensors; }
void aule_tensor_clear_all(void) {
    for (uint32_t i = 0; i < kMaxTensors; ++i)
        if (g_tensors[i].live) aule_tensor_destroy((uint64_t)i + 1);
}

int32_t aule_attention_forward_gpu(uint64_t qh, uint64_t kh, uint64_t vh, uint64_t oh, uint64_t rot_cos,
                                   uint64_t rot_sin, int32_t causal, int32_t window_size) {
    if (!g_engine.ready()) return -1;
    TensorSlot *q = slot_of(qh), *k = slot_of(kh), *v = slot_of(vh), *o = slot_of(oh);
    if (!q || !k || !v || !o) return -1;
    // RoPE argument pairing of AttentionEngine.forward (attention_gpu.zig:406-424): both or neither
    TensorSlot *rc = nullptr, *rs = nullptr;
    if ((rot_cos != 0) != (rot_sin != 0)) { set_error("Attention failed: rot_cos and rot_sin must be given together"); return -3; }
    if (rot_cos != 0) {
        rc = slot_of(rot_cos); rs = slot_of(rot_sin);
        if (!rc || !rs) return -1;
    }
    // Shape rules of AttentionEngine.forward (attention_gpu.zig:372-400).
    if (k->shape[0] != q->shape[0] || k->shape[3] != q->shape[3]) { set_error("Attention failed: K batch/head_dim must match Q"); return -3; }
    for (int i = 0; i < 4; ++i) {
        if (v->shape[i] != k->shape[i]) { set_error("Attention failed: V shape must match K"); return -3; }
        if (o->shape[i] != q->shape[i]) { set_error("Attention failed: output shape must match Q"); return -3; }
    }
    aule::AttnShape s{q->shape[0], q->shape[1], k->shape[1], q->shape[2], k->shape[2], q->shape[3]};
    const int dev = primary_device();
    aule::Device* d = g_engine.by_ordinal(dev);
    std::string e;
    if (rc) {
        // The shader rotates interleaved pairs (2i, 2i+1) of Q and K by the angle of their row index, tables
        // [.., seq, head_dim/2] (attention_f32.comp:98-111, tests/test_rope_unit.py:46-47): convention 1 of the RoPE kernel.
        const uint32_t half = s.D / 2;
        if (half == 0 || rc->count != rs->count || rc->count % half != 0) { set_error("Attention failed: rot_cos / rot_sin must hold [seq, head_dim/2] values"); return -3; }
        e = g_engine.forward_rope(dev, d->s_compute, q->ptr, k->ptr, v->ptr, o->ptr, 0, rc->ptr, rs->ptr, rc->count / half, 1, s,
                                  aule::kF32, 0.f, causal != 0, window_size);
    } else {
        e = g_engine.forward(dev, d->s_compute, q->ptr, k->ptr, v->ptr, o->ptr, 0, s, aule::kF32, 0.f, causal != 0, window_size);
    }
    if (e.empty()) e = g_engine.synchronize(dev);       // the reference call is synchronous (attention_gpu.zig:456-469)
    if (!e.empty()) { set_error("Attention failed: %s", e.c_str()); return -3; }
    return 0;
}

// ---------------------------------------------------------------- out-of-scope exports (stubs)
// lib.zig:533-566 -> AttentionEngine.forwardPaged (attention_gpu.zig:484-653): the reference takes ordinary contiguous
// K/V tensors, copies them into its private 32-token block pool and runs attention over the pool -- the paging is an
// internal storage detail, the result is aule_attention_forward_gpu's.  Here the same tensors go straight to the
// fused kernel (no private pool to copy into); error text / code follow lib.zig:561-563.
int32_t aule_attention_forward_paged(uint64_t q, uint64_t k, uint64_t v, uint64_t output, uint64_t rot_cos,
                                     uint64_t rot_sin, int32_t causal, int32_t window_size) {
    int32_t rc = aule_attention_forward_gpu(q, k, v, output, rot_cos, rot_sin, causal, window_size);
    if (rc == -3) {
        std::string msg = aule_get_error();
        const std::string from = "Attention failed";
        if (msg.compare(0, from.size(), from) == 0) msg = "PagedAttention failed" + msg.substr(from.size());
        set_error("%s", msg.c_str());
    }
    return rc;
}
int32_t aule_spatial_sort(uint64_t, uint64_t, uint64_t, uint32_t) {
    set_error("aule_spatial_sort: spatial sort is outside the B200 hot path (unsupported)");
    return -10;
}
int32_t aule_attention_forward_gravity(uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, int32_t,
                                       uint32_t, int32_t) {
    set_error("aule_attention_forward_gravity: gravity attention is outside the B200 hot path (unsupported)");
    return -10;
}

// ---------------------------------------------------------------- Part 2: B200 extension
int32_t aule_attention_forward_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o, uint64_t lse_or_0, uint32_t B,
                                    uint32_t Hq, uint32_t Hkv, uint32_t Sq, uint32_t Sk, uint32_t D, int32_t dtype,
                                    float scale, int32_t causal, int32_t window, int32_t device, uint64_t cu_stream) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{B, Hq, Hkv, Sq, Sk, D};
    std::string e = g_engine.forward(device, (CUstream)cu_stream, q, k, v, o, lse_or_0, s, dtype, scale, causal != 0, window);
    if (!e.empty()) { set_error("Attention failed: %s", e.c_str()); return -4; }
    return 0;
}

int32_t aule_attention_backward_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o, uint64_t d_o, uint64_t lse,
                                     uint64_t dq, uint64_t dk, uint64_t dv, uint32_t B, uint32_t Hq, uint32_t Hkv,
                                     uint32_t Sq, uint32_t Sk, uint32_t D, int32_t dtype, float scale, int32_t causal,
                                     int32_t device, uint64_t cu_stream) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{B, Hq, Hkv, Sq, Sk, D};
    std::string e = g_engine.backward(device, (CUstream)cu_stream, q, k, v, o, d_o, lse, dq, dk, dv, s, dtype, scale, causal != 0);
    if (!e.empty()) { set_error("Backward pass failed: %s", e.c_str()); return -4; }
    return 0;
}

int32_t aule_attention_backward_window_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o, uint64_t d_o, uint64_t lse,
                                            uint64_t dq, uint64_t dk, uint64_t dv, uint32_t B, uint32_t Hq, uint32_t Hkv,
                                            uint32_t Sq, uint32_t Sk, uint32_t D, int32_t dtype, float scale, int32_t causal,
                                            int32_t window, int32_t device, uint64_t cu_stream) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{B, Hq, Hkv, Sq, Sk, D};
    std::string e = g_engine.backward(device, (CUstream)cu_stream, q, k, v, o, d_o, lse, dq, dk, dv, s, dtype, scale, causal != 0, window);
    if (!e.empty()) { set_error("Backward pass failed: %s", e.c_str()); return -4; }
    return 0;
}

int32_t aule_attention_forward_host(const void* q, const void* k, const void* v, void* o, float* lse_or_null,
                                    uint32_t B, uint32_t Hq, uint32_t Hkv, uint32_t Sq, uint32_t Sk, uint32_t D,
                                    int32_t dtype, float scale, int32_t causal, int32_t window, int32_t device) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{B, Hq, Hkv, Sq, Sk, D};
    int code = 0;
    std::string e = g_engine.forward_host(device, q, k, v, o, lse_or_null, s, dtype, scale, causal != 0, window, &code);
    if (!e.empty()) { set_error("Attention failed: %s", e.c_str()); return code ? code : -4; }
    return 0;
}

int32_t aule_rope_dptr(uint64_t x, uint64_t out, uint64_t cos, uint64_t sin, uint32_t B, uint32_t H, uint32_t S,
                       uint32_t D, uint32_t table_rows, int32_t interleaved, int32_t dtype, int32_t inverse, int32_t device,
                       uint64_t cu_stream) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    std::string e = g_engine.rope(device, (CUstream)cu_stream, x, out, (uint64_t)B * H, S, 0, 0, 0, 0, cos, sin, table_rows, D,
                                  interleaved ? 1 : 0, dtype, inverse ? -1.f : 1.f);
    if (!e.empty()) { set_error("RoPE failed: %s", e.c_str()); return -4; }
    return 0;
}

int32_t aule_attention_forward_rope_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o, uint64_t lse_or_0, uint64_t cos,
                                         uint64_t sin, uint32_t table_rows, int32_t interleaved, uint32_t B, uint32_t Hq,
                                         uint32_t Hkv, uint32_t Sq, uint32_t Sk, uint32_t D, int32_t dtype, float scale,
                                         int32_t causal, int32_t window, int32_t device, uint64_t cu_stream) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{B, Hq, Hkv, Sq, Sk, D};
    std::string e = g_engine.forward_rope(device, (CUstream)cu_stream, q, k, v, o, lse_or_0, cos, sin, table_rows,
                                          interleaved ? 1 : 0, s, dtype, scale, causal != 0, window);
    if (!e.empty()) { set_error("Attention failed: %s", e.c_str()); return -4; }
    return 0;
}

int32_t aule_attention_forward_spanning_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o, uint64_t lse_or_0, uint32_t B,
                                             uint32_t Hq, uint32_t Hkv, uint32_t Sq, uint32_t Sk, uint32_t D, int32_t dtype,
                                             float scale, int32_t causal, int32_t window, int32_t src_device,
                                             uint64_t cu_stream, const int32_t* devices, int32_t num_devices, int32_t chunks,
                                             float* timings_ms_or_null) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{B, Hq, Hkv, Sq, Sk, D};
    std::string e = g_engine.forward_spanning(src_device, (CUstream)cu_stream, q, k, v, o, lse_or_0, s, dtype, scale, causal != 0,
                                              window, devices, num_devices, chunks, timings_ms_or_null);
    if (!e.empty()) { set_error("Attention failed: %s", e.c_str()); return -4; }
    return 0;
}

int32_t aule_attention_backward_host(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                     const float* lse, void* dq, void* dk, void* dv, uint32_t B, uint32_t Hq, uint32_t Hkv,
                                     uint32_t Sq, uint32_t Sk, uint32_t D, int32_t dtype, float scale, int32_t causal,
                                     int32_t device) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    aule::AttnShape s{B, Hq, Hkv, Sq, Sk, D};
    int code = 0;
    std::string e = g_engine.backward_host(device, q, k, v, o, d_o, lse, dq, dk, dv, s, dtype, scale, causal != 0, &code);
    if (!e.empty()) { set_error("Backward pass failed: %s", e.c_str()); return code ? code : -4; }
    return 0;
}

int32_t aule_attention_paged_decode_dptr(uint64_t q, uint64_t k_cache, uint64_t v_cache, uint64_t block_tables,
                                         uint64_t context_lens, uint64_t out, uint32_t B, uint32_t Hq, uint32_t Hkv,
                                         uint32_t D, uint32_t num_blocks, uint32_t block_size,
                                         uint32_t max_blocks_per_seq, uint32_t max_context_len, int32_t dtype,
                                         float scale, int32_t window, int32_t device, uint64_t cu_stream) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    std::string e = g_engine.paged_decode(device, (CUstream)cu_stream, q, k_cache, v_cache, block_tables, context_lens,
                                          out, B, Hq, Hkv, D, num_blocks, block_size, max_blocks_per_seq,
                                          max_context_len, dtype, scale, window);
    if (!e.empty()) { set_error("PagedAttention failed: %s", e.c_str()); return -4; }
    return 0;
}

int32_t aule_device_count(void) { return g_engine.ready() ? g_engine.device_count() : -1; }
int32_t aule_get_sm_count(int32_t device) {
    aule::Device* d = g_engine.ready() ? g_engine.by_ordinal(device) : nullptr;
    return d ? d->sm_count : -1;
}
int32_t aule_synchronize(int32_t device) {
    std::string e = g_engine.synchronize(device);
    if (!e.empty()) { set_error("Synchronize failed: %s", e.c_str()); return -4; }
    return 0;
}
uint64_t aule_launch_count(void) { return g_engine.launch_count(); }
const char* aule_last_kernel(void) {
    snprintf(g_name_buf, sizeof(g_name_buf), "%s", g_engine.last_kernel().c_str());
    return g_name_buf;
}
int32_t aule_set_kernel_path(int32_t path) {
    g_engine.set_kernel_path(path);
    return 0;
}
int32_t aule_set_trace_buffer(uint64_t dptr) {
    g_engine.set_trace_buffer(dptr);
    return 0;
}
int32_t aule_smoke_multiply(const float* in, float* out, uint32_t n) {
    if (!g_engine.ready()) { set_error("Library not initialized. Call aule_init() first."); return -1; }
    std::string e = g_engine.smoke_multiply(primary_device(), in, out, n);
    if (!e.empty()) { set_error("Smoke kernel failed: %s", e.c_str()); return -4; }
    return 0;
}
const char* aule_version(void) { return "aule-b200 0.2.0 (abi 0.5.0+dptr2)"; }

}  // extern "C"
