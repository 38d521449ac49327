// Paged-KV decode attention for sm_100a.
//
// What it computes follows the reference's decode kernel python/aule/triton_flash_amd.py:544-660
// (_paged_attention_fwd_amd) and its wrapper :662-740 (flash_attention_paged_amd): one query token per sequence,
//   out[b,h,:] = softmax_j(scale * q[b,h,:] . K[b,j,h_kv,:]) V[b,j,h_kv,:],   j < context_lens[b],
//   K[b,j] = k_cache[block_tables[b, j / block_size], j % block_size],   h_kv = h / (Hq / Hkv)    (:573-575)
// with the optional sliding window "(context_len - 1 - j) < window_size" (:618-621).
//
// How it is built here is unrelated to the reference (one Triton program per (b, h) that walks the whole block
// table and re-reads K/V once per q-head of a GQA group).  This path is HBM-bound: every byte of the live KV cache
// must cross the memory bus once and nothing else should.
//   * CTA = (sequence b, kv head, split).  All Hq/Hkv (<= 16) q-heads of the group are the M rows of one
//     m16n8k16 tensor-core tile, so K/V are read ONCE per group, not once per q-head.
//   * A producer warp walks the block table (32 entries per coalesced load, prefetched one batch ahead) and streams
//     16-token K and V tiles into an 8-stage shared-memory ring with 5-D TMA (box = 64 x D/64 x 1 head x 16 tokens x
//     1 page, 128-byte swizzle) -- a token's D*2 bytes are contiguous in the cache and arrive as one request.
//   * Four consumer warps each own every fourth tile: S = Q K^T (mma.sync, K fragments by ldmatrix), masked online
//     softmax in the log2 domain, O += P V (V fragments by ldmatrix.trans; P never leaves registers).
//   * The warps' (m, l, O) are merged through shared memory; with nsplit > 1 the CTA writes an unnormalised
//     partial to a workspace and aule_paged_combine_* merges the splits (flash-decoding).
// tcgen05 is deliberately not used: the M dimension is the GQA group size (<= 16 rows), a 128-row UMMA tile would
// waste 8-32x of the tensor pipe and TMEM round trips for nothing -- the kernel's ceiling is HBM bandwidth, and
// mma.sync at < 10 % of its peak keeps up with it.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "kernel_params.h"
#include "sm100_ptx.cuh"

namespace paged100 {
using namespace sm100;
using aule_kp::PagedCfg;
using aule_kp::PagedParams;

__device__ __forceinline__ void tma_load_5d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3, int32_t c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Byte offset of the 16-byte chunk `cg` (8 elements, cg < D/8) of token `tok` inside a [16][D/64][128 B] tile
// written by TMA with CU_TENSOR_MAP_SWIZZLE_128B (chunk index XOR (128-byte row index & 7)).
template <int D>
__device__ __forceinline__ uint32_t tile_off(uint32_t tok, uint32_t cg) {
    constexpr uint32_t HALVES = D / 64;
    const uint32_t r = tok * HALVES + (cg >> 3);
    return r * 128u + (((cg & 7u) ^ (r & 7u)) << 4);
}

template <int D, bool BF16>
__device__ __forceinline__ void paged_body(const CUtensorMap* tmK, const CUtensorMap* tmV, const PagedParams& p) {
    using Cfg = PagedCfg<D>;
    constexpr int NS = Cfg::NS, TOK = Cfg::TOK, NW = Cfg::CONSUMERS;
    constexpr int KSTEPS = D / 16, NTILES = D / 8;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t ring = smem_u32(smem + Cfg::OFF_RING);
    float* red_o = reinterpret_cast<float*>(smem + Cfg::OFF_RED_O);
    float* red_ml = reinterpret_cast<float*>(smem + Cfg::OFF_RED_ML);
    const uint32_t bar_full = smem_u32(smem + Cfg::OFF_BAR), bar_empty = bar_full + NS * 8;

    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t split = blockIdx.x % p.nsplit;
    const uint32_t hkv = (blockIdx.x / p.nsplit) % p.Hkv;
    const uint32_t b = blockIdx.x / (p.nsplit * p.Hkv);
    const uint32_t G = p.Hq / p.Hkv;

    // this CTA's tile range: tiles of 16 tokens that intersect [lo, ctx), split evenly
    const int32_t ctx = max(p.context_lens[b], 0);
    const int32_t lo = (p.window > 0) ? max(ctx - p.window, 0) : 0;
    const uint32_t tile_first = (uint32_t)lo / TOK, tile_end = ((uint32_t)ctx + TOK - 1) / TOK;
    const uint32_t per = (tile_end - tile_first + p.nsplit - 1) / p.nsplit;
    const uint32_t t0 = min(tile_first + split * per, tile_end), t1 = min(t0 + per, tile_end);
    const uint32_t ntiles = t1 - t0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, 1);
        }
        fence_mbar_init();
        tma_prefetch_desc(tmK);
        tma_prefetch_desc(tmV);
    }
    __syncthreads();

    if (warp == NW) {
        // ------------------------------------------------------------------ producer: ONE thread
        // (A warp-wide loop with shuffles measured producer-bound: ~20 % of all issue slots went to broadcasting
        // coordinates into uniform registers and to a runtime modulo per tile.)  The thread keeps the next PF block
        // table entries in registers, loaded one group ahead, and advances (page, tile-in-page) incrementally.
        if (lane == 0 && ntiles > 0) {
            constexpr int PF = 16;
            const int32_t* table = p.block_tables + (size_t)b * p.max_blocks;
            const uint32_t tiles_per_page = p.block_size / TOK;
            uint32_t pg = t0 / tiles_per_page;                  // first page of the range
            uint32_t sub = t0 - pg * tiles_per_page;            // first 16-token tile inside it
            const uint32_t pg_end = (t1 + tiles_per_page - 1) / tiles_per_page;
            int32_t cur[PF], nxt[PF];
#pragma unroll
            for (int j = 0; j < PF; ++j) cur[j] = (pg + j < pg_end && pg + j < p.max_blocks) ? table[pg + j] : -1;
            uint32_t i = 0;                                     // tile counter of this CTA
            while (i < ntiles) {
#pragma unroll
                for (int j = 0; j < PF; ++j)
                    nxt[j] = (pg + PF + j < pg_end && pg + PF + j < p.max_blocks) ? table[pg + PF + j] : -1;
#pragma unroll
                for (int j = 0; j < PF; ++j) {
                    const int32_t page = cur[j];
                    for (; sub < tiles_per_page && i < ntiles; ++sub, ++i) {
                        const uint32_t s = i % NS;
                        if (i >= NS) mbar_wait(bar_empty + s * 8, ((i / NS) - 1) & 1);
                        const uint32_t dst = ring + s * Cfg::STAGE_BYTES;
                        mbar_expect_tx(bar_full + s * 8, Cfg::STAGE_BYTES);
                        // an out-of-range page index is out of bounds for the tensor map and reads as zeros
                        tma_load_5d(dst, tmK, bar_full + s * 8, 0, 0, (int32_t)hkv, (int32_t)(sub * TOK), page);
                        tma_load_5d(dst + Cfg::TILE_BYTES, tmV, bar_full + s * 8, 0, 0, (int32_t)hkv, (int32_t)(sub * TOK), page);
                    }
                    sub = 0;
                }
                pg += PF;
#pragma unroll
                for (int j = 0; j < PF; ++j) cur[j] = nxt[j];
            }
        }
    } else {
        // ------------------------------------------------------------------ consumer warps
        const uint32_t g = lane >> 2, t = lane & 3;
        // Q fragments (A operand, rows = q-heads of the group, zero rows past G), pre-scaled later through scale_log2
        uint32_t qa[KSTEPS][4];
        {
            const uint32_t* q32 = reinterpret_cast<const uint32_t*>(p.q);
            const size_t row0 = ((size_t)b * p.Hq + (size_t)hkv * G + g) * (D / 2);
            const size_t row1 = row0 + (size_t)8 * (D / 2);
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                qa[ks][0] = (g < G) ? q32[row0 + ks * 8 + t] : 0u;
                qa[ks][1] = (g + 8 < G) ? q32[row1 + ks * 8 + t] : 0u;
                qa[ks][2] = (g < G) ? q32[row0 + ks * 8 + 4 + t] : 0u;
                qa[ks][3] = (g + 8 < G) ? q32[row1 + ks * 8 + 4 + t] : 0u;
            }
        }
        float o[NTILES][4];
#pragma unroll
        for (int n = 0; n < NTILES; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g + 8; l is this thread's partial sum

        for (uint32_t i = warp; i < ntiles; i += NW) {
            const uint32_t s = i % NS;
            const uint32_t kt = ring + s * Cfg::STAGE_BYTES, vt = kt + Cfg::TILE_BYTES;
            const int32_t pos0 = (int32_t)((t0 + i) * TOK);
            const bool ragged = pos0 < lo || pos0 + TOK > ctx;
            mbar_wait(bar_full + s * 8, (i / NS) & 1);
            if (ragged) {
                // Tokens outside [lo, ctx) get P = 0, but 0 x NaN garbage in an uninitialised page would still poison
                // the accumulator: clear their V rows (D*2 bytes = 16 or 8 lanes of 16 B per token).
                constexpr uint32_t LPT = D / 8;                      // lanes per token row
                for (uint32_t tk = lane / LPT; tk < TOK; tk += 32 / LPT) {
                    const int32_t pos = pos0 + (int32_t)tk;
                    if (pos < lo || pos >= ctx)
                        *reinterpret_cast<uint4*>(smem + Cfg::OFF_RING + s * Cfg::STAGE_BYTES + Cfg::TILE_BYTES +
                                                  tk * (D * 2) + (lane % LPT) * 16) = make_uint4(0, 0, 0, 0);
                }
                __syncwarp();
            }
            // ---- S[h][tok] = Q K^T : 2 n-tiles of 8 tokens
            float sc[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
                for (int q4 = 0; q4 < D / 32; ++q4) {
                    uint32_t kb[4];   // matrices: d-chunks 4q4 .. 4q4+3 of tokens 8j .. 8j+7
                    ldsm_x4(kb, kt + tile_off<D>(8 * j + (lane & 7), 4 * q4 + (lane >> 3)));
                    mma16816<BF16>(sc[j], qa[2 * q4], kb[0], kb[1]);
                    mma16816<BF16>(sc[j], qa[2 * q4 + 1], kb[2], kb[3]);
                }
            }
            // ---- mask + online softmax (log2 domain)
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int32_t pos = pos0 + 8 * j + 2 * (int32_t)t + e;
                    const bool ok = !ragged || (pos >= lo && pos < ctx);
                    sc[j][e] = ok ? sc[j][e] * p.scale_log2 : -INFINITY;
                    sc[j][2 + e] = ok ? sc[j][2 + e] * p.scale_log2 : -INFINITY;
                    mx0 = fmaxf(mx0, sc[j][e]);
                    mx1 = fmaxf(mx1, sc[j][2 + e]);
                }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
            const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0, ms1 = (mn1 == -INFINITY) ? 0.f : mn1;
            const float a0 = ex2(m0 - ms0), a1 = ex2(m1 - ms1);     // exp2(-inf) = 0 on the first tile
            m0 = mn0; m1 = mn1;
            uint32_t pa[4];
            {
                const float p00 = ex2(sc[0][0] - ms0), p01 = ex2(sc[0][1] - ms0), p02 = ex2(sc[0][2] - ms1), p03 = ex2(sc[0][3] - ms1);
                const float p10 = ex2(sc[1][0] - ms0), p11 = ex2(sc[1][1] - ms0), p12 = ex2(sc[1][2] - ms1), p13 = ex2(sc[1][3] - ms1);
                l0 = l0 * a0 + (p00 + p01 + p10 + p11);
                l1 = l1 * a1 + (p02 + p03 + p12 + p13);
                pa[0] = pack2<BF16>(p00, p01); pa[1] = pack2<BF16>(p02, p03);
                pa[2] = pack2<BF16>(p10, p11); pa[3] = pack2<BF16>(p12, p13);
            }
            if (__any_sync(0xffffffffu, a0 != 1.f || a1 != 1.f)) {
#pragma unroll
                for (int n = 0; n < NTILES; ++n) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }
            }
            // ---- O[h][d] += P V : one k-step of 16 tokens, D/8 n-tiles
#pragma unroll
            for (int n2 = 0; n2 < NTILES / 2; ++n2) {
                uint32_t vb[4];   // matrices: (tok 0-7, chunk 2n2) (tok 8-15, chunk 2n2) (tok 0-7, chunk 2n2+1) (tok 8-15, chunk 2n2+1)
                ldsm_x4_t(vb, vt + tile_off<D>(((lane >> 3) & 1) * 8 + (lane & 7), 2 * n2 + (lane >> 4)));
                mma16816<BF16>(o[2 * n2], pa, vb[0], vb[1]);
                mma16816<BF16>(o[2 * n2 + 1], pa, vb[2], vb[3]);
            }
            if (ragged) fence_proxy_async_smem();               // our generic-proxy stores precede the next TMA write
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + s * 8);
        }
        // ---- publish this warp's state
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        float* wo = red_o + warp * 16 * D;
#pragma unroll
        for (int n = 0; n < NTILES; ++n) {
            *reinterpret_cast<float2*>(wo + g * D + 8 * n + 2 * t) = make_float2(o[n][0], o[n][1]);
            *reinterpret_cast<float2*>(wo + (g + 8) * D + 8 * n + 2 * t) = make_float2(o[n][2], o[n][3]);
        }
        if (t == 0) {
            red_ml[(warp * 16 + g) * 2] = m0; red_ml[(warp * 16 + g) * 2 + 1] = l0;
            red_ml[(warp * 16 + g + 8) * 2] = m1; red_ml[(warp * 16 + g + 8) * 2 + 1] = l1;
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------------- merge the warps, write out
    for (uint32_t idx = threadIdx.x; idx < G * D; idx += blockDim.x) {
        const uint32_t h = idx / D, d = idx % D;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; ++w) M = fmaxf(M, red_ml[(w * 16 + h) * 2]);
        float L = 0.f, acc = 0.f;
        if (M != -INFINITY) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float f = ex2(red_ml[(w * 16 + h) * 2] - M);
                L += red_ml[(w * 16 + h) * 2 + 1] * f;
                acc += red_o[(w * 16 + h) * D + d] * f;
            }
        }
        const size_t row = (size_t)b * p.Hq + (size_t)hkv * G + h;
        if (p.nsplit == 1) {
            const float r = (L > 0.f) ? acc / L : 0.f;           // empty context: zeros (the reference divides 0 by 0)
            if constexpr (BF16) reinterpret_cast<__nv_bfloat16*>(p.out)[row * D + d] = __float2bfloat16_rn(r);
            else                reinterpret_cast<__half*>(p.out)[row * D + d] = __float2half_rn(r);
        } else {
            p.ws_o[(row * p.nsplit + split) * D + d] = acc;
            if (d == 0) {
                p.ws_ml[(row * p.nsplit + split) * 2] = M;
                p.ws_ml[(row * p.nsplit + split) * 2 + 1] = L;
            }
        }
    }
}

// Merge the nsplit partials of one (sequence, q-head): grid = B*Hq, block = D threads.
template <bool BF16>
__device__ __forceinline__ void combine_body(const PagedParams& p, uint32_t D) {
    const size_t row = blockIdx.x;
    const uint32_t d = threadIdx.x;
    const float* ml = p.ws_ml + row * p.nsplit * 2;
    float M = -INFINITY;
    for (uint32_t s = 0; s < p.nsplit; ++s) M = fmaxf(M, ml[2 * s]);
    float L = 0.f, acc = 0.f;
    if (M != -INFINITY)
        for (uint32_t s = 0; s < p.nsplit; ++s) {
            const float f = ex2(ml[2 * s] - M);
            L += ml[2 * s + 1] * f;
            acc += p.ws_o[(row * p.nsplit + s) * D + d] * f;
        }
    const float r = (L > 0.f) ? acc / L : 0.f;
    if constexpr (BF16) reinterpret_cast<__nv_bfloat16*>(p.out)[row * D + d] = __float2bfloat16_rn(r);
    else                reinterpret_cast<__half*>(p.out)[row * D + d] = __float2half_rn(r);
}

}  // namespace paged100

#define AULE_PAGED100(NAME, DD, BF)                                                                          \
    extern "C" __global__ void __launch_bounds__(aule_kp::PagedCfg<DD>::THREADS, 2)                          \
        NAME(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,               \
             const aule_kp::PagedParams p) {                                                                 \
        paged100::paged_body<DD, BF>(&tmK, &tmV, p);                                                         \
    }
AULE_PAGED100(aule_paged_sm100_bf16_d64, 64, true)
AULE_PAGED100(aule_paged_sm100_bf16_d128, 128, true)
AULE_PAGED100(aule_paged_sm100_f16_d64, 64, false)
AULE_PAGED100(aule_paged_sm100_f16_d128, 128, false)
extern "C" __global__ void aule_paged_combine_bf16(const aule_kp::PagedParams p, uint32_t D) { paged100::combine_body<true>(p, D); }
extern "C" __global__ void aule_paged_combine_f16(const aule_kp::PagedParams p, uint32_t D) { paged100::combine_body<false>(p, D); }
