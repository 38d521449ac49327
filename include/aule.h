/*
 * libaule C ABI  --  the drop-in boundary of the B200-native attention engine.
 *
 * Part 1 re-declares, symbol for symbol, the C ABI the reference exports from
 * src/lib.zig (so /root/reference/python/aule/vulkan.py loads this library
 * unmodified through ctypes.CDLL, vulkan.py:205-207,:224-406).  Each entry
 * cites the reference export it replaces.
 *
 * Part 2 is the extension the reference does not have and BASELINE.json's
 * north_star requires: raw device pointers (PyTorch tensor.data_ptr()),
 * bf16/fp16, explicit scale, GQA, D=128, a caller-supplied CUstream, and a
 * pipelined host-buffer entry used for end-to-end measurements.
 *
 * Conventions (same as the reference, lib.zig:12-26,:124-130):
 *   - all tensors are contiguous row-major [B, H, S, D];
 *   - 0 = success, negative = failure, message via aule_get_error();
 *   - global state, single-threaded by contract;
 *   - no CPU fallback: every compute entry fails with an error when no
 *     sm_100 device / CUDA driver is present.
 */
#ifndef AULE_H_
#define AULE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define AULE_API __declspec(dllexport)
#else
#define AULE_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------ */
/* Part 1: the reference's ABI (src/lib.zig)                                 */
/* ------------------------------------------------------------------------ */

/* lib.zig:59-93   0 ok / -1 fail; idempotent. Opens the CUDA driver (dlopen
 * libcuda.so.1), retains the primary context of every visible sm_100 device,
 * loads the embedded sm_100a cubin (cuModuleLoadData). */
AULE_API int32_t aule_init(void);
/* lib.zig:105-122  destroys all handle-table tensors, then the contexts. */
AULE_API void aule_shutdown(void);
/* lib.zig:124-130  static 512-byte buffer; "No error" if none. */
AULE_API const char* aule_get_error(void);
/* lib.zig:133-141 */
AULE_API const char* aule_get_backend_name(void);
/* lib.zig:96-103   1 when the backward kernels are available. */
AULE_API int32_t aule_supports_backward(void);
/* lib.zig:144-163 / :274-293   0 other,1 amd,2 nvidia,3 intel,4 apple,-1 not init */
AULE_API int32_t aule_get_vendor(void);
AULE_API int32_t aule_get_gpu_vendor(void);
/* lib.zig:242-271  bytes written or -1 */
AULE_API int32_t aule_get_device_name(char* buffer, uint32_t buffer_len);
/* lib.zig:168-197,:296-309 */
AULE_API int32_t aule_is_amd_optimized(void);
AULE_API int32_t aule_has_fp16(void);
AULE_API int32_t aule_get_subgroup_size(void);
/* lib.zig:202-239  single "sm100" variant: id 0 accepted, others -2 */
AULE_API int32_t aule_set_shader_variant(uint8_t variant);
AULE_API int32_t aule_get_shader_variant(void);
AULE_API int32_t aule_has_shader_variant(uint8_t variant);

/* lib.zig:312-367  host fp32 pointers, MHA, Sq == Sk, scale = 1/sqrt(D).
 * codes: -1 not-init, -2 alloc, -3 upload, -4 compute, -5 download */
AULE_API int32_t aule_attention_forward(const float* query, const float* key, const float* value,
                                        float* output, uint32_t batch_size, uint32_t num_heads,
                                        uint32_t seq_len, uint32_t head_dim, int32_t causal);
/* lib.zig:765-852  also writes LSE [B,H,S] (natural log, scaled scores) */
AULE_API int32_t aule_attention_forward_with_lse(const float* query, const float* key,
                                                 const float* value, float* output, float* lse_out,
                                                 uint32_t batch_size, uint32_t num_heads,
                                                 uint32_t seq_len, uint32_t head_dim, int32_t causal);
/* lib.zig:639-762 */
AULE_API int32_t aule_attention_backward(const float* query, const float* key, const float* value,
                                         const float* output, const float* grad_output,
                                         const float* lse, float* grad_query, float* grad_key,
                                         float* grad_value, uint32_t batch_size, uint32_t num_heads,
                                         uint32_t seq_len, uint32_t head_dim, int32_t causal);

/* lib.zig:409-455  u64 handle = slot+1, 0 = failure; at most 1024 live (:17).
 * Tensors are TRUE device memory here (the reference uses host-visible
 * memory, gpu_tensor.zig:50). Element type is 32-bit (f32 / u32). */
AULE_API uint64_t aule_tensor_create(uint32_t batch_size, uint32_t num_heads, uint32_t seq_len,
                                     uint32_t head_dim);
AULE_API uint64_t aule_tensor_create_u32(uint32_t batch_size, uint32_t num_heads, uint32_t seq_len,
                                         uint32_t head_dim);
AULE_API void aule_tensor_destroy(uint64_t handle);
/* lib.zig:457-494  -1 bad handle / not init, -3 size mismatch */
AULE_API int32_t aule_tensor_upload(uint64_t handle, const float* data, uint32_t count);
AULE_API int32_t aule_tensor_download(uint64_t handle, float* output, uint32_t count);
AULE_API int32_t aule_tensor_download_u32(uint64_t handle, uint32_t* output, uint32_t count);
/* lib.zig:626-634 */
AULE_API uint32_t aule_tensor_size(uint64_t handle);
/* lib.zig:383-407 */
AULE_API uint32_t aule_tensor_count(void);
AULE_API uint32_t aule_tensor_max(void);
AULE_API void aule_tensor_clear_all(void);

/* lib.zig:496-529  shapes come from the tensors => GQA and Sq != Sk work.
 * rot_cos/rot_sin must be 0 (RoPE prologue is a SURVEY 8f "next" row): -3.
 * window_size -1 = full. codes: -1 bad handle, -3 compute error */
AULE_API int32_t aule_attention_forward_gpu(uint64_t q, uint64_t k, uint64_t v, uint64_t output,
                                            uint64_t rot_cos, uint64_t rot_sin, int32_t causal,
                                            int32_t window_size);

/* lib.zig:533-566 (AttentionEngine.forwardPaged, attention_gpu.zig:484-653): same tensors and result as
 * aule_attention_forward_gpu -- the reference copies K/V into a private 32-token block pool first, which is an
 * internal storage detail; here they go straight to the fused kernel.  codes: -1 bad handle, -3 compute error.
 * (The serving-side paged KV cache with caller-owned block tables is aule_attention_paged_decode_dptr below.) */
AULE_API int32_t aule_attention_forward_paged(uint64_t q, uint64_t k, uint64_t v, uint64_t output,
                                              uint64_t rot_cos, uint64_t rot_sin, int32_t causal,
                                              int32_t window_size);
/* Out-of-scope product features (SURVEY 2.1 rows 13,14). Exported because
 * vulkan.py:300-316 sets their prototypes unconditionally; they return -10
 * ("unsupported") and set the error string.  lib.zig:568-624 */
AULE_API int32_t aule_spatial_sort(uint64_t keys, uint64_t values, uint64_t indices,
                                   uint32_t sort_dim);
AULE_API int32_t aule_attention_forward_gravity(uint64_t q, uint64_t k, uint64_t v, uint64_t output,
                                                uint64_t rot_cos, uint64_t rot_sin, uint64_t indices,
                                                int32_t causal, uint32_t max_attend,
                                                int32_t window_size);

/* ------------------------------------------------------------------------ */
/* Part 2: B200 extension (replaces the slot python/aule/__init__.py:185-207  */
/* fills with Triton: FlashAttentionTritonFunc, triton_flash.py:386-526)     */
/* ------------------------------------------------------------------------ */

/* AULE_DTYPE_F32_TF32: fp32 tensors in memory whose forward MAY run on the tensor cores as tf32 -- what the reference's GPU
 * path does with fp32 inputs (tl.dot on fp32 operands, triton_flash.py:405-411, Triton's default allow_tf32); taken for
 * head_dim <= 64 (error <= 1e-3 relative to the output scale), exactly AULE_DTYPE_F32 everywhere else and in every backward.
 * AULE_DTYPE_F32 itself is always exact fp32 arithmetic (the Vulkan shaders' behaviour, attention_f32.comp). */
enum { AULE_DTYPE_F32 = 0, AULE_DTYPE_BF16 = 1, AULE_DTYPE_F16 = 2, AULE_DTYPE_F32_TF32 = 3 };

/* Fused forward on raw device pointers (CUdeviceptr), asynchronous on
 * `cu_stream` of `device` (no implicit sync).
 *   q,o: [B,Hq,Sq,D]  k,v: [B,Hkv,Sk,D]  lse_or_0: [B,Hq,Sq] fp32 or 0
 *   scale <= 0  => 1/sqrt(D)      (triton_flash.py:394-395)
 *   causal      => key j visible to query i iff j <= i (top-left, :187)
 *   window      => -1 full; W>0 causal keeps 0 <= i-j < W, bidirectional keeps |i-j| <= W/2 (the Vulkan shader's
 *                  convention, attention_f32.comp:173-183; the generic Triton kernel's W (triton_flash.py:190-194:
 *                  i-j <= W, |i-j| <= W) is W+1 causal / 2W bidirectional here)
 *   GQA         => kv_head = q_head / (Hq/Hkv)          (triton_flash.py:95-96)
 * bf16/fp16 with D <= 128, D % 8 == 0 run the tcgen05/TMA kernel (D is zero-padded to 64 / 128 by the TMA boxes,
 * like BLOCK_K = next_power_of_2(D) in triton_flash.py:446); AULE_DTYPE_F32_TF32 with D <= 64 runs the kind::tf32 variant of
 * the same kernel; fp32 and D % 8 != 0 run the fp32-accumulate CUDA-core kernel (AULE_LOG=1 reports that choice on stderr).
 * Replaces _flash_attn_fwd_kernel launch, triton_flash.py:448-464. */
AULE_API int32_t aule_attention_forward_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o,
                                             uint64_t lse_or_0, uint32_t B, uint32_t Hq, uint32_t Hkv,
                                             uint32_t Sq, uint32_t Sk, uint32_t D, int32_t dtype,
                                             float scale, int32_t causal, int32_t window,
                                             int32_t device, uint64_t cu_stream);

/* Backward on raw device pointers. dq: [B,Hq,Sq,D], dk/dv: [B,Hkv,Sk,D] in
 * `dtype`; dK/dV are summed over the q-heads of a GQA group on-device
 * (deterministic, no atomics).  Replaces _compute_delta_kernel +
 * _flash_attn_bwd_kernel launches, triton_flash.py:495-524. */
AULE_API int32_t aule_attention_backward_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o,
                                              uint64_t d_o, uint64_t lse, uint64_t dq, uint64_t dk,
                                              uint64_t dv, uint32_t B, uint32_t Hq, uint32_t Hkv,
                                              uint32_t Sq, uint32_t Sk, uint32_t D, int32_t dtype,
                                              float scale, int32_t causal, int32_t device,
                                              uint64_t cu_stream);

/* Backward of a forward that used a sliding window (same `window` meaning as aule_attention_forward_dptr).  The
 * reference's backward ignores the window (triton_flash.py:313-319: window_size is never saved), which makes its gradients
 * wrong for windowed attention; here the mask is applied.  Runs the deterministic CUDA-core kernels (the tensor-core
 * backward covers causal / full masks); window <= 0 is aule_attention_backward_dptr. */
AULE_API int32_t aule_attention_backward_window_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o, uint64_t d_o,
                                                     uint64_t lse, uint64_t dq, uint64_t dk, uint64_t dv, uint32_t B,
                                                     uint32_t Hq, uint32_t Hkv, uint32_t Sq, uint32_t Sk, uint32_t D,
                                                     int32_t dtype, float scale, int32_t causal, int32_t window,
                                                     int32_t device, uint64_t cu_stream);

/* Host-buffer forward for any dtype: stages through device memory with the
 * H2D copy of batch b+1 / D2H copy of batch b-1 overlapped with the kernel of
 * batch b (3 streams).  Page-locked caller buffers are copied directly; pageable ones (NumPy arrays) go through the
 * library's own pinned bounce buffers, double-buffered per chunk.  Synchronous.  lse may be NULL.
 * Thread-safe: host-pointer entries are serialised per device, device-pointer entries may run concurrently.
 * This is the call bench.py times for the end-to-end ("e2e") figure. */
AULE_API int32_t aule_attention_forward_host(const void* q, const void* k, const void* v, void* o,
                                             float* lse_or_null, uint32_t B, uint32_t Hq,
                                             uint32_t Hkv, uint32_t Sq, uint32_t Sk, uint32_t D,
                                             int32_t dtype, float scale, int32_t causal,
                                             int32_t window, int32_t device);

/* Host-buffer backward for any dtype (q,k,v,o,dO,lse in; dq,dk,dv out), staged through device memory.  Synchronous.
 * The typed twin of aule_attention_backward (lib.zig:639-762), used by bench.py for the forward+backward e2e figure. */
AULE_API int32_t aule_attention_backward_host(const void* q, const void* k, const void* v, const void* o,
                                              const void* d_o, const float* lse, void* dq, void* dk, void* dv,
                                              uint32_t B, uint32_t Hq, uint32_t Hkv, uint32_t Sq, uint32_t Sk,
                                              uint32_t D, int32_t dtype, float scale, int32_t causal, int32_t device);

/* Rotary position embedding on raw device pointers.  Two pairing conventions, as in the reference:
 *   interleaved == 0, half-split (Triton path, python/aule/triton_flash.py:680-703 apply_rope_separate):
 *     out[d] = x[d] cos[s,d] - x[d+D/2] sin[s,d],  out[d+D/2] = x[d+D/2] cos[s,d] + x[d] sin[s,d]
 *   interleaved != 0, adjacent pairs (Vulkan shader, shaders/attention_f32.comp:98-111; tests/test_rope_unit.py:76-85):
 *     out[2d] = x[2d] cos[s,d] - x[2d+1] sin[s,d],  out[2d+1] = x[2d] sin[s,d] + x[2d+1] cos[s,d]
 * x/out: [B,H,S,D] of `dtype`, cos/sin: [table_rows, D/2] fp32 with table_rows >= S (checked: a short table is an error,
 * never an out-of-bounds read); inverse != 0 applies the transpose (backward pass). */
AULE_API int32_t aule_rope_dptr(uint64_t x, uint64_t out, uint64_t cos, uint64_t sin, uint32_t B, uint32_t H,
                                uint32_t S, uint32_t D, uint32_t table_rows, int32_t interleaved, int32_t dtype,
                                int32_t inverse, int32_t device, uint64_t cu_stream);

/* RoPE prologue + fused attention behind one call (flash_attention_rope, triton_flash.py:561-603; the handle form is
 * aule_attention_forward_gpu with rot_cos / rot_sin, lib.zig:496-529).  Q and K are rotated by ONE launch that reads each
 * of them once and writes the rotated copy once into a stream-ordered workspace (K must be rotated once per key, not once
 * per (query block, key) pair, so the rotation cannot live in the attention kernel's K/V stream), then the fused kernel
 * runs on the copies.  Arguments as aule_attention_forward_dptr + the tables of aule_rope_dptr. */
AULE_API int32_t aule_attention_forward_rope_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o, uint64_t lse_or_0,
                                                  uint64_t cos, uint64_t sin, uint32_t table_rows, int32_t interleaved,
                                                  uint32_t B, uint32_t Hq, uint32_t Hkv, uint32_t Sq, uint32_t Sk,
                                                  uint32_t D, int32_t dtype, float scale, int32_t causal,
                                                  int32_t window, int32_t device, uint64_t cu_stream);

/* One call spanning several devices of the box (SURVEY 8e "spanning call"): q,k,v,o (and lse) live on `src_device`;
 * the (batch, kv-head) units are split contiguously over devices[0..num_devices) (src_device must be one of them).
 * Every other device receives its Q/K/V slabs over NVLink (cuMemcpyPeerAsync), runs the fused kernel and returns its
 * O / LSE slab; `chunks` sub-ranges per device overlap copy-in, kernel and copy-out (1 = strictly serial phases).
 * The caller's stream waits for the gathered result.  timings_ms_or_null -> float[4]: scatter, kernel, gather (max over
 * devices of each phase's device-timed duration) and the total on the source stream (the call then blocks). */
AULE_API int32_t aule_attention_forward_spanning_dptr(uint64_t q, uint64_t k, uint64_t v, uint64_t o, uint64_t lse_or_0,
                                                      uint32_t B, uint32_t Hq, uint32_t Hkv, uint32_t Sq, uint32_t Sk,
                                                      uint32_t D, int32_t dtype, float scale, int32_t causal,
                                                      int32_t window, int32_t src_device, uint64_t cu_stream,
                                                      const int32_t* devices, int32_t num_devices, int32_t chunks,
                                                      float* timings_ms_or_null);

/* Paged-KV decode on raw device pointers: one query token per sequence against a block-table
 * KV cache in the vLLM layout.  Replaces flash_attention_paged_amd / _paged_attention_fwd_amd
 * (python/aule/triton_flash_amd.py:662-740 / :544-660), same argument meaning:
 *   q, out        [B, Hq, D]                              `dtype` (bf16 | f16)
 *   k_cache/v_cache [num_blocks, block_size, Hkv, D]      `dtype`, contiguous, 16-byte aligned
 *   block_tables  [B, max_blocks_per_seq] int32           page of tokens [i*block_size, (i+1)*block_size)
 *   context_lens  [B] int32                               tokens of the sequence that are live (<= max_blocks*block_size)
 *   scale <= 0 => 1/sqrt(D) (:699-700);  window > 0 keeps keys with (context_len-1-pos) < window (:618-621), -1 = all
 *   kv head of q head h = h / (Hq/Hkv) (:573-575), Hq/Hkv <= 16;  D in {64,128};  block_size % 16 == 0, <= 256
 *   max_context_len: an upper bound of context_lens known on the host (the reference reads it back with .item(),
 *   :711); only sizes the split-KV grid, 0 = max_blocks_per_seq*block_size.  A sequence with context_len 0 gets zeros.
 * Asynchronous on `cu_stream`; HBM-bound (reads each live K/V byte once per kv head).  -4 + error string on failure. */
AULE_API int32_t aule_attention_paged_decode_dptr(uint64_t q, uint64_t k_cache, uint64_t v_cache,
                                                  uint64_t block_tables, uint64_t context_lens, uint64_t out,
                                                  uint32_t B, uint32_t Hq, uint32_t Hkv, uint32_t D,
                                                  uint32_t num_blocks, uint32_t block_size,
                                                  uint32_t max_blocks_per_seq, uint32_t max_context_len,
                                                  int32_t dtype, float scale, int32_t window, int32_t device,
                                                  uint64_t cu_stream);

/* Device bookkeeping. */
AULE_API int32_t aule_device_count(void);               /* sm_100 devices usable, -1 not init */
AULE_API int32_t aule_get_sm_count(int32_t device);     /* 148 on B200 */
AULE_API int32_t aule_synchronize(int32_t device);      /* cuCtxSynchronize on that device */
/* Launch accounting for tests / bench ("gpu_launches"): number of kernels this
 * library has launched since init, and the name of the last one. */
AULE_API uint64_t aule_launch_count(void);
AULE_API const char* aule_last_kernel(void);
/* Test hooks. aule_set_kernel_path: 0 = automatic choice, 1 = force the CUDA-core
 * (fp32-accumulate) kernels for every dtype -- used to cross-check the tensor-core kernel
 * on the GPU itself; higher bits select A/B variants of the tensor-core kernels (forward tuning
 * variants, head pairing, L2 runs, cross-item prefetch, backward kernel generation / timing split;
 * the bit map is in csrc/host/engine.h, Engine::set_kernel_path).  aule_smoke_multiply: out[i] = 2*in[i] through the whole
 * module-load / launch / copy path (the tests/test_multiply.zig analogue,
 * src/compute_pipeline.zig:203-254 + shaders/test.comp). */
AULE_API int32_t aule_set_kernel_path(int32_t path);
/* Bring-up hook: a device buffer (>= 3 x 4096 u64, zeroed by the caller) into which CTA 0 of the backward kernels
 * records (tag << 48 | clock64) events of its issuer thread and of one thread per compute group; 0 disables.
 * Read by tools/bwd_trace.py to draw the per-step pipeline timeline. */
AULE_API int32_t aule_set_trace_buffer(uint64_t dptr);
AULE_API int32_t aule_smoke_multiply(const float* in, float* out, uint32_t n);
/* Library / ABI version string. */
AULE_API const char* aule_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AULE_H_ */
