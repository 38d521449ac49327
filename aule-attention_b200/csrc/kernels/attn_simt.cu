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

// visible(i, j): key j contributes to query i.
__device__ __forceinline__ bool visible(uint32_t i, uint32_t j, uint32_t Sk, int causal, int window) {
    if (j >= Sk) return false;
    if (causal && j > i) return false;
    if (window > 0) {
        const int64_t d = (int64_t)i - (int64_t)j;
        if (d >= window) return false;
        if (!causal && -d >= window) return false;
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
            const bool vis = visible(row, kb + j, p.Sk, p.causal, -1) && row < p.Sq;
            const float pj = vis ? expf(s - lse) : 0.f;            // P = exp(S - LSE)   (:321)
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
                const bool vis = (qb + i < p.Sq) && visible(qb + i, key, p.Sk, p.causal, -1);
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
// Rotary position embedding, half-split convention of the reference's Triton path
// (python/aule/triton_flash.py:680-703 apply_rope_separate; rotate_half(x) = [-x2, x1]):
//   out[d]       = x[d]       * cos[s,d] - x[d + D/2] * sin[s,d]
//   out[d + D/2] = x[d + D/2] * cos[s,d] + x[d]       * sin[s,d]          d < D/2,  cos/sin: [S, D/2] fp32
// `sign` = -1 applies the transpose (inverse rotation), which is what the backward pass needs.
// Memory-bound elementwise pass (x read once, out written once); fp32 math.
template <typename T>
__device__ __forceinline__ void rope_body(const T* x, T* out, const float* cs, const float* sn, uint64_t rows,
                                          uint32_t S, uint32_t D, float sign) {
    const uint32_t half = D / 2;
    const uint64_t total = rows * half;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = i / half;
        const uint32_t d = (uint32_t)(i - row * half);
        const uint32_t s = (uint32_t)(row % S);
        const float c = cs[(size_t)s * half + d], sv = sign * sn[(size_t)s * half + d];
        const float x1 = Elem<T>::ld(x, row * D + d), x2 = Elem<T>::ld(x, row * D + d + half);
        Elem<T>::st(out, row * D + d, x1 * c - x2 * sv);
        Elem<T>::st(out, row * D + d + half, x2 * c + x1 * sv);
    }
}
extern "C" __global__ void aule_rope_f32(const float* x, float* out, const float* cs, const float* sn, uint64_t rows, uint32_t S, uint32_t D, float sign) { rope_body(x, out, cs, sn, rows, S, D, sign); }
extern "C" __global__ void aule_rope_bf16(const __nv_bfloat16* x, __nv_bfloat16* out, const float* cs, const float* sn, uint64_t rows, uint32_t S, uint32_t D, float sign) { rope_body(x, out, cs, sn, rows, S, D, sign); }
extern "C" __global__ void aule_rope_f16(const __half* x, __half* out, const float* cs, const float* sn, uint64_t rows, uint32_t S, uint32_t D, float sign) { rope_body(x, out, cs, sn, rows, S, D, sign); }
