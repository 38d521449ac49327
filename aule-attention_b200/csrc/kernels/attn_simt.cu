// CUDA-core fused attention (fp32 math) for sm_100a: forward (+LSE) and a
// deterministic two-kernel backward.  This is the engine's fp32 path (config A,
// the legacy fp32 C ABI) and the path for head dims the tensor-core kernel does
// not cover; bf16/fp16 inputs with D in {64,128} take attn_fwd_sm100.cu instead.
//
// Replaces (semantics, not code): shaders/attention_f32.comp:1-225,
// attention_forward_f32.comp:1-189, attention_backward_f32.comp:1-234 and the
// fp32 branch of python/aule/triton_flash.py:62-350.
//
// Work decomposition (all three kernels): a CTA owns 32 consecutive rows of one
// (batch, head); each row is shared by 8 lanes ("slices"), lane s holding the
// float4 groups d = 4*s + 32*i of the head dimension, so a (row, key) dot product
// is a per-lane partial sum + 3 xor-shuffles, and shared-memory reads of a key row
// are one 128-byte conflict-free wavefront broadcast to the 4 rows of a warp.
// HBM traffic: Q/O once, K/V once per 32-row block (L2-resident across blocks).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "kernel_params.h"

namespace {

constexpr int ROWS = 32;      // rows per CTA
constexpr int LANES = 8;      // lanes per row
constexpr int TILE = 32;      // keys (or queries) staged per shared-memory tile
constexpr int MAXG = 4;       // float4 groups per lane -> D <= 128
constexpr int DP = 128;       // padded row length in shared memory (floats)

using aule_kp::SimtParams;

template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ float ld(const void* p, size_t i) { return ((const float*)p)[i]; }
    static __device__ __forceinline__ void st(void* p, size_t i, float x) { ((float*)p)[i] = x; }
};
template <> struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const void* p, size_t i) { return __bfloat162float(((const __nv_bfloat16*)p)[i]); }
    static __device__ __forceinline__ void st(void* p, size_t i, float x) { ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(x); }
};
template <> struct Elem<__half> {
    static __device__ __forceinline__ float ld(const void* p, size_t i) { return __half2float(((const __half*)p)[i]); }
    static __device__ __forceinline__ void st(void* p, size_t i, float x) { ((__half*)p)[i] = __float2half_rn(x); }
};

__device__ __forceinline__ float lane8_sum(float x) {
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    return x;
}

// Stage `TILE` rows [row0, row0+TILE) of a [S, D] matrix into smem[TILE][DP] as fp32,
// zero-filling rows >= S and columns >= D.
template <typename T>
__device__ __forceinline__ void stage_tile(float (*sm)[DP], const void* base, size_t mat_off,
                                           uint32_t row0, uint32_t S, uint32_t D, uint32_t dpad) {
    for (uint32_t idx = threadIdx.x; idx < TILE * dpad; idx += blockDim.x) {
        const uint32_t r = idx / dpad, c = idx - r * dpad;
        float x = 0.f;
        if (row0 + r < S && c < D) x = Elem<T>::ld(base, mat_off + (size_t)(row0 + r) * D + c);
        sm[r][c] = x;
    }
}

// Row fragment (this lane's float4 groups) of row `row` of a [S,D] matrix.
template <typename T>
__device__ __forceinline__ void load_frag(float (&f)[MAXG][4], const void* base, size_t mat_off,
                                          uint32_t row, uint32_t S, uint32_t D, int slice) {
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t d = 4 * slice + 32 * g + e;
            f[g][e] = (row < S && d < D) ? Elem<T>::ld(base, mat_off + (size_t)row * D + d) : 0.f;
        }
}

template <typename T>
__device__ __forceinline__ void store_frag(const float (&f)[MAXG][4], void* base, size_t mat_off,
                                           uint32_t row, uint32_t S, uint32_t D, int slice, float mul) {
    if (row >= S) return;
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t d = 4 * slice + 32 * g + e;
            if (d < D) Elem<T>::st(base, mat_off + (size_t)row * D + d, f[g][e] * mul);
        }
}

__device__ __forceinline__ float frag_dot(const float (&a)[MAXG][4], const float* srow, int slice, int ngroups) {
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
        if (g < ngroups) {
            const float4 x = *reinterpret_cast<const float4*>(srow + 4 * slice + 32 * g);
            acc = fmaf(a[g][0], x.x, acc); acc = fmaf(a[g][1], x.y, acc);
            acc = fmaf(a[g][2], x.z, acc); acc = fmaf(a[g][3], x.w, acc);
        }
    return acc;
}

__device__ __forceinline__ void frag_axpy(float (&a)[MAXG][4], float w, const float* srow, int slice, int ngroups) {
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
        if (g < ngroups) {
            const float4 x = *reinterpret_cast<const float4*>(srow + 4 * slice + 32 * g);
            a[g][0] = fmaf(w, x.x, a[g][0]); a[g][1] = fmaf(w, x.y, a[g][1]);
            a[g][2] = fmaf(w, x.z, a[g][2]); a[g][3] = fmaf(w, x.w, a[g][3]);
        }
}

// visible(i, j): key j contributes to query i.  Sliding window (the Vulkan shader's convention, attention_f32.comp:173-183):
// causal keeps 0 <= i-j < W, bidirectional keeps |i-j| <= W/2.
__device__ __forceinline__ bool visible(uint32_t i, uint32_t j, uint32_t Sk, int causal, int window) {
    if (j >= Sk) return false;
    if (causal && j > i) return false;
    if (window > 0) {
        const int64_t d = (int64_t)i - (int64_t)j;
        if (causal) { if (d >= window) return false; }
        else if (d > window / 2 || -d > window / 2) return false;
    }
    return true;
}

// ---------------------------------------------------------------- forward
template <typename T>
__device__ void fwd_body(const SimtParams& p) {
    __shared__ __align__(16) float sK[TILE][DP];
    __shared__ __align__(16) float sV[TILE][DP];
    const int slice = threadIdx.x & (LANES - 1);
    const int rloc = threadIdx.x / LANES;
    const uint32_t b = blockIdx.z, hq = blockIdx.y, hk = hq / (p.Hq / p.Hkv);
    const uint32_t row0 = blockIdx.x * ROWS, row = row0 + rloc;
    const uint32_t dpad = (p.D + 31) & ~31u;
    const int ngroups = dpad / 32;
    const size_t qoff = ((size_t)b * p.Hq + hq) * p.Sq * p.D;
    const size_t koff = ((size_t)b * p.Hkv + hk) * p.Sk * p.D;

    float qf[MAXG][4], acc[MAXG][4];
    load_frag<T>(qf, p.q, qoff, row, p.Sq, p.D, slice);
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[g][e] = 0.f;
    float m = -INFINITY, l = 0.f;

    uint32_t kv_end = p.Sk, kv_begin = 0;
    if (p.causal) kv_end = min(p.Sk, row0 + ROWS);
    if (p.window > 0 && p.causal && row0 + 1 > (uint32_t)p.window) kv_begin = ((row0 + 1 - p.window) / TILE) * TILE;

    for (uint32_t kb = kv_begin; kb < kv_end; kb += TILE) {
        __syncthreads();
        stage_tile<T>(sK, p.k, koff, kb, p.Sk, p.D, dpad);
        stage_tile<T>(sV, p.v, koff, kb, p.Sk, p.D, dpad);
        __syncthreads();
        float s[TILE];
        float tmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < TILE; ++j) {
            float d = lane8_sum(frag_dot(qf, sK[j], slice, ngroups)) * p.scale;
            d = visible(row, kb + j, p.Sk, p.causal, p.window) ? d : -INFINITY;
            s[j] = d;
            tmax = fmaxf(tmax, d);
        }
        const float m_new = fmaxf(m, tmax);
        if (m_new == -INFINITY) continue;               // nothing visible yet for this row
        const float corr = __expf(m - m_new);           // m == -inf -> 0
        l *= corr;
#pragma unroll
        for (int g = 0; g < MAXG; ++g)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[g][e] *= corr;
#pragma unroll
        for (int j = 0; j < TILE; ++j) {
            const float pj = expf(s[j] - m_new);        // -inf -> 0
            l += pj;
            frag_axpy(acc, pj, sV[j], slice, ngroups);
        }
        m = m_new;
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    store_frag<T>(acc, p.o, qoff, row, p.Sq, p.D, slice, inv);
    if (p.lse && slice == 0 && row < p.Sq)
        p.lse[((size_t)b * p.Hq + hq) * p.Sq + row] = (l > 0.f) ? m + logf(l) : -INFINITY;
}

// ---------------------------------------------------------------- backward, kernel 1: delta + dQ
template <typename T>
__device__ void bwd_dq_body(const SimtParams& p) {
    __shared__ __align__(16) float sK[TILE][DP];
    __shared__ __align__(16) float sV[TILE][DP];
    const int slice = threadIdx.x & (LANES - 1);
    const int rloc = threadIdx.x / LANES;
    const uint32_t b = blockIdx.z, hq = blockIdx.y, hk = hq / (p.Hq / p.Hkv);
    const uint32_t row0 = blockIdx.x * ROWS, row = row0 + rloc;
    const uint32_t dpad = (p.D + 31) & ~31u;
    const int ngroups = dpad / 32;
    const size_t qoff = ((size_t)b * p.Hq + hq) * p.Sq * p.D;
    const size_t koff = ((size_t)b * p.Hkv + hk) * p.Sk * p.D;
    const size_t roff = ((size_t)b * p.Hq + hq) * p.Sq;

    float qf[MAXG][4], dof[MAXG][4], acc[MAXG][4];
    load_frag<T>(qf, p.q, qoff, row, p.Sq, p.D, slice);
    load_frag<T>(dof, p.d_o, qoff, row, p.Sq, p.D, slice);
    float dlt = 0.f;                                     // Delta_i = sum_d O_id dO_id (triton_flash.py:353-379)
    {
        float of[MAXG][4];
        load_frag<T>(of, p.o, qoff, row, p.Sq, p.D, slice);
#pragma unroll
        for (int g = 0; g < MAXG; ++g)
#pragma unroll
            for (int e = 0; e < 4; ++e) { dlt = fmaf(of[g][e], dof[g][e], dlt); acc[g][e] = 0.f; }
        dlt = lane8_sum(dlt);
        if (slice == 0 && row < p.Sq) p.delta[roff + row] = dlt;   // consumed by the dK/dV kernel
    }
    const float lse = (row < p.Sq) ? p.lse[roff + row] : 0.f;

    uint32_t kv_end = p.Sk;
    if (p.causal) kv_end = min(p.Sk, row0 + ROWS);
    for (uint32_t kb = 0; kb < kv_end; kb += TILE) {
        __syncthreads();
        stage_tile<T>(sK, p.k, koff, kb, p.Sk, p.D, dpad);
        stage_tile<T>(sV, p.v, koff, kb, p.Sk, p.D, dpad);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < TILE; ++j) {
            const float s = lane8_sum(frag_dot(qf, sK[j], slice, ngroups)) * p.scale;
            const float dp = lane8_sum(frag_dot(dof, sV[j], slice, ngroups));
            const bool vis = visible(row, kb + j, p.Sk, p.causal, p.window) && row < p.Sq;
            const float pj = vis ? expf(s - lse) : 0.f;            // P = exp(S - LSE)   (:321); a row without a visible key never gets here
            const float ds = pj * (dp - dlt) * p.scale;            // dS = P o (dP - Delta) * scale (:330)
            frag_axpy(acc, ds, sK[j], slice, ngroups);             // dQ += dS K (:336)
        }
    }
    store_frag<T>(acc, p.dq, qoff, row, p.Sq, p.D, slice, 1.f);
}

// ---------------------------------------------------------------- backward, kernel 2: dK, dV
// CTA = 32 key rows of one (batch, kv head); loops over the q-heads of the GQA group and
// over query tiles from the diagonal on (triton_flash.py:286-294), so the group sum of
// dK/dV (triton_flash.py:345-347 does it with atomics) is a plain register accumulation.
template <typename T>
__device__ void bwd_dkv_body(const SimtParams& p) {
    __shared__ __align__(16) float sQ[TILE][DP];
    __shared__ __align__(16) float sDO[TILE][DP];
    __shared__ float sLse[TILE], sDelta[TILE];
    const int slice = threadIdx.x & (LANES - 1);
    const int rloc = threadIdx.x / LANES;
    const uint32_t b = blockIdx.z, hk = blockIdx.y, group = p.Hq / p.Hkv;
    const uint32_t key0 = blockIdx.x * ROWS, key = key0 + rloc;
    const uint32_t dpad = (p.D + 31) & ~31u;
    const int ngroups = dpad / 32;
    const size_t koff = ((size_t)b * p.Hkv + hk) * p.Sk * p.D;

    float kf[MAXG][4], vf[MAXG][4], dk[MAXG][4], dv[MAXG][4];
    load_frag<T>(kf, p.k, koff, key, p.Sk, p.D, slice);
    load_frag<T>(vf, p.v, koff, key, p.Sk, p.D, slice);
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dk[g][e] = 0.f; dv[g][e] = 0.f; }

    const uint32_t q_begin = p.causal ? (key0 / TILE) * TILE : 0;   // queries i >= key only
    for (uint32_t gq = 0; gq < group; ++gq) {
        const uint32_t hq = hk * group + gq;
        const size_t qoff = ((size_t)b * p.Hq + hq) * p.Sq * p.D;
        const size_t roff = ((size_t)b * p.Hq + hq) * p.Sq;
        for (uint32_t qb = q_begin; qb < p.Sq; qb += TILE) {
            __syncthreads();
            stage_tile<T>(sQ, p.q, qoff, qb, p.Sq, p.D, dpad);
            stage_tile<T>(sDO, p.d_o, qoff, qb, p.Sq, p.D, dpad);
            if (threadIdx.x < TILE) {
                const uint32_t i = qb + threadIdx.x;
                sLse[threadIdx.x] = i < p.Sq ? p.lse[roff + i] : 0.f;
                sDelta[threadIdx.x] = i < p.Sq ? p.delta[roff + i] : 0.f;
            }
            __syncthreads();
#pragma unroll 4
            for (int i = 0; i < TILE; ++i) {
                const float s = lane8_sum(frag_dot(kf, sQ[i], slice, ngroups)) * p.scale;
                const float dp = lane8_sum(frag_dot(vf, sDO[i], slice, ngroups));
                const bool vis = (qb + i < p.Sq) && visible(qb + i, key, p.Sk, p.causal, p.window);
                const float pj = vis ? expf(s - sLse[i]) : 0.f;
                frag_axpy(dv, pj, sDO[i], slice, ngroups);                  // dV += P^T dO (:324)
                const float ds = pj * (dp - sDelta[i]) * p.scale;
                frag_axpy(dk, ds, sQ[i], slice, ngroups);                   // dK += dS^T Q (:333)
            }
        }
    }
    store_frag<T>(dk, p.dk, koff, key, p.Sk, p.D, slice, 1.f);
    store_frag<T>(dv, p.dv, koff, key, p.Sk, p.D, slice, 1.f);
}

}  // namespace

#define AULE_SIMT_KERNELS(SUFFIX, TYPE)                                                              \
    extern "C" __global__ void __launch_bounds__(ROWS * LANES) aule_fwd_simt_##SUFFIX(const SimtParams p) { fwd_body<TYPE>(p); }      \
    extern "C" __global__ void __launch_bounds__(ROWS * LANES) aule_bwd_dq_simt_##SUFFIX(const SimtParams p) { bwd_dq_body<TYPE>(p); } \
    extern "C" __global__ void __launch_bounds__(ROWS * LANES) aule_bwd_dkv_simt_##SUFFIX(const SimtParams p) { bwd_dkv_body<TYPE>(p); }

AULE_SIMT_KERNELS(f32, float)
AULE_SIMT_KERNELS(bf16, __nv_bfloat16)
AULE_SIMT_KERNELS(f16, __half)

// Launch-path smoke kernel (analogue of shaders/test.comp driven by
// src/compute_pipeline.zig / tests/test_multiply.zig): out[i] = 2 * in[i].
extern "C" __global__ void aule_smoke_multiply(const float* in, float* out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 2.f * in[i];
}

// dtype conversion helpers used by the legacy fp32 host ABI when it stages
// through the tensor-core path, and by the host-buffer entry.
extern "C" __global__ void aule_cvt_f32_to_bf16(const float* in, __nv_bfloat16* out, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}
extern "C" __global__ void aule_cvt_bf16_to_f32(const __nv_bfloat16* in, float* out, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = __bfloat162float(in[i]);
}

// ---------------------------------------------------------------------------------------------------
// Rotary position embedding of Q and K in ONE launch (the prologue of flash_attention_rope,
// python/aule/triton_flash.py:561-603): each tensor is read once and written once.  Two pairing conventions:
//   mode 0, half-split (Triton path, triton_flash.py:680-703 apply_rope_separate; rotate_half(x) = [-x2, x1]):
//       out[d]       = x[d]       * cos[s,d] - x[d + D/2] * sin[s,d]
//       out[d + D/2] = x[d + D/2] * cos[s,d] + x[d]       * sin[s,d]          d < D/2
//   mode 1, interleaved pairs (Vulkan shader, shaders/attention_f32.comp:98-111; tests/test_rope_unit.py:76-85):
//       out[2d]     = x[2d] * cos[s,d] - x[2d+1] * sin[s,d]
//       out[2d + 1] = x[2d] * sin[s,d] + x[2d+1] * cos[s,d]
// cos/sin: [table_rows, D/2] fp32, position s = row index inside its sequence (the host checks table_rows >= S).
// `sign` = -1 applies the transpose (inverse rotation), which is what the backward pass needs.
// Memory-bound; fp32 math; four pairs per thread (8- or 16-byte accesses) when D % 8 == 0, one pair otherwise.
template <typename T>
__device__ __forceinline__ void ld4f(const T* p, float (&f)[4]);
template <>
__device__ __forceinline__ void ld4f<float>(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void ld4f<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}
template <>
__device__ __forceinline__ void ld4f<__half>(const __half* p, float (&f)[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void st4f(T* p, const float (&f)[4]);
template <>
__device__ __forceinline__ void st4f<float>(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <>
__device__ __forceinline__ void st4f<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&a); v.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = v;
}
template <>
__device__ __forceinline__ void st4f<__half>(__half* p, const float (&f)[4]) {
    __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&a); v.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = v;
}

using aule_kp::RopeParams;

template <typename T>
__device__ __forceinline__ void rope_body(const RopeParams& p) {
    const uint32_t half = p.D / 2;
    const bool vec = (p.D % 8) == 0;
    const uint32_t per_row = vec ? half / 4 : half;              // work items per row
    const uint64_t nq = p.rows_q * per_row, total = nq + p.rows_k * per_row;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const bool second = i >= nq;
        const uint64_t w = second ? i - nq : i;
        const T* x = reinterpret_cast<const T*>(second ? p.xk : p.xq);
        T* out = reinterpret_cast<T*>(second ? p.ok : p.oq);
        const uint32_t S = second ? p.Sk : p.Sq;
        const uint64_t row = w / per_row;
        const uint32_t u = (uint32_t)(w - row * per_row);
        const uint32_t s = (uint32_t)(row % S);
        if (vec) {
            const uint32_t d = 4 * u;                            // first of four pairs
            float c[4], sv[4], a[4], b[4], ra[4], rb[4];
            ld4f<float>(p.cs + (size_t)s * half + d, c);
            ld4f<float>(p.sn + (size_t)s * half + d, sv);
            if (p.mode == 0) {
                ld4f<T>(x + row * p.D + d, a);
                ld4f<T>(x + row * p.D + d + half, b);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ra[e] = a[e] * c[e] - b[e] * (p.sign * sv[e]);
                    rb[e] = b[e] * c[e] + a[e] * (p.sign * sv[e]);
                }
                st4f<T>(out + row * p.D + d, ra);
                st4f<T>(out + row * p.D + d + half, rb);
            } else {
                ld4f<T>(x + row * p.D + 2 * d, a);               // pairs (a0,a1) (a2,a3) (b0,b1) (b2,b3)
                ld4f<T>(x + row * p.D + 2 * d + 4, b);
                ra[0] = a[0] * c[0] - a[1] * (p.sign * sv[0]); ra[1] = a[0] * (p.sign * sv[0]) + a[1] * c[0];
                ra[2] = a[2] * c[1] - a[3] * (p.sign * sv[1]); ra[3] = a[2] * (p.sign * sv[1]) + a[3] * c[1];
                rb[0] = b[0] * c[2] - b[1] * (p.sign * sv[2]); rb[1] = b[0] * (p.sign * sv[2]) + b[1] * c[2];
                rb[2] = b[2] * c[3] - b[3] * (p.sign * sv[3]); rb[3] = b[2] * (p.sign * sv[3]) + b[3] * c[3];
                st4f<T>(out + row * p.D + 2 * d, ra);
                st4f<T>(out + row * p.D + 2 * d + 4, rb);
            }
        } else {
            const uint32_t d = u;
            const float c = p.cs[(size_t)s * half + d], sv = p.sign * p.sn[(size_t)s * half + d];
            const uint64_t i1 = p.mode == 0 ? row * p.D + d : row * p.D + 2 * d, i2 = p.mode == 0 ? i1 + half : i1 + 1;
            const float x1 = Elem<T>::ld(x, i1), x2 = Elem<T>::ld(x, i2);
            Elem<T>::st(out, i1, x1 * c - x2 * sv);
            Elem<T>::st(out, i2, x2 * c + x1 * sv);
        }
    }
}
extern "C" __global__ void aule_rope_f32(const RopeParams p) { rope_body<float>(p); }
extern "C" __global__ void aule_rope_bf16(const RopeParams p) { rope_body<__nv_bfloat16>(p); }
extern "C" __global__ void aule_rope_f16(const RopeParams p) { rope_body<__half>(p); }
