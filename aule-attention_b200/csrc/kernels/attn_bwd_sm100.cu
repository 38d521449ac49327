// Fused attention backward for sm_100a (B200): TMA -> SMEM -> tcgen05.mma -> TMEM.   (kernel v1)
//
// Replaces (behaviour, not code): python/aule/triton_flash.py:242-379 (_flash_attn_bwd_kernel,
// _compute_delta_kernel) and shaders/attention_backward_f32.comp:1-234.  Same loop nest as the reference
// (KV block outer, query blocks from the diagonal on, triton_flash.py:262-294) and the same math (:321-347):
//     P = exp(S*scale - LSE)   dP = dO V^T   dS = P o (dP - Delta)
//     dV += P^T dO             dK += dS^T Q  dQ += dS K          (dK, dQ scaled by `scale` at the end)
// Differences by design: all five contractions run on the tensor cores; dK/dV of a GQA group are summed in
// TMEM by the one CTA that owns the KV block (the reference uses global atomics, :345-347); dQ is reduced in
// FP32 (vector red.global.add into a workspace, converted once) where the reference uses atomics in the
// input dtype (:335-338).
//
// One 288-thread CTA per (KV block of 128 keys, kv head, batch), heaviest (lowest block) first.      (kernel v2)
//   warps 0-3 / 4-7  compute: thread == (query row, key half): S,dP TMEM -> P,dS (16-bit) -> swizzled SMEM; dQ
//                    partial TMEM -> fp32 vector reductions into the workspace
//   warp  8          issuer: one elected thread issues every TMA load and every tcgen05.mma
// K_j and V_j stay resident in SMEM; dV [256,384) and dK [384,512) accumulate in TMEM across every q-head of the
// group and every query block; S [0,128) and dP [128,256) are rewritten per step and the dQ partial reuses dP's
// columns, so that S(i+1) = Q(i+1) K^T (Q is double-buffered) runs while the compute warps drain dQ(i).
// Every operand tile is the same [128 rows][64 elements] 128B-swizzled sub-tile, which serves as a K-major
// operand (Q, K, V, dO, dS for S / dP / dQ) and, read through an MN-major descriptor, as its own transpose
// (P^T, dS^T, and Q, dO, K as [k][n] B operands) -- no data is ever transposed.
// Step pipeline (issuer):  S(i) | wait dO(i), dQ(i-1) drained: dP(i) | wait P,dS(i): dV,dK,dQ(i) ; S(i+1) |
//                          wait all MMAs(i): load dO(i+1), Q(i+2).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sm100_ptx.cuh"
#include "kernel_params.h"

namespace bwd100 {
using namespace sm100;
using aule_kp::BwdParams;
template <int D> using Cfg = aule_kp::BwdCfg<D>;

template <int D, bool BF16>
__device__ __forceinline__ void bwd_body(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                         const CUtensorMap* tmdO, const CUtensorMap* tmdK, const CUtensorMap* tmdV,
                                         const BwdParams& p) {
    using C = Cfg<D>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barriers
    const uint32_t bar_kv = sb + C::OFF_BAR;        // K_j, V_j landed (once)
    const uint32_t bar_q0 = bar_kv + 8;             // Q buffer 0 / 1 landed (+8)
    const uint32_t bar_do = bar_kv + 24;            // dO landed
    const uint32_t bar_sdp = bar_kv + 32;           // S and dP complete (commit)
    const uint32_t bar_pds = bar_kv + 40;           // compute -> issuer: P, dS in SMEM (256 arrivals)
    const uint32_t bar_d = bar_kv + 48;             // dV, dK, dQ MMAs complete (commit)
    const uint32_t bar_dqf = bar_kv + 56;           // compute -> issuer: dQ drained from TMEM (256 arrivals)
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_SLOT);
    const uint32_t sK = sb + C::OFF_K, sV = sb + C::OFF_V, sQ0 = sb + C::OFF_Q, sdO = sb + C::OFF_DO,
                   sP = sb + C::OFF_P, sdS = sb + C::OFF_DS;

    if (threadIdx.x == 0) {
        if (sb & 1023u) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }
        mbar_init(bar_kv, 1); mbar_init(bar_q0, 1); mbar_init(bar_q0 + 8, 1); mbar_init(bar_do, 1);
        mbar_init(bar_sdp, 1); mbar_init(bar_pds, 256); mbar_init(bar_d, 1); mbar_init(bar_dqf, 256);
        fence_mbar_init();
        tma_prefetch_desc(tmQ); tma_prefetch_desc(tmK); tma_prefetch_desc(tmV); tma_prefetch_desc(tmdO);
    }
    if (warp == 8) tmem_alloc<512>(sb + C::OFF_TMEM_SLOT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 384;

    // ---- which KV block
    const uint32_t per = p.Hkv * p.B;
    const uint32_t jb = blockIdx.x / per;                            // KV block (0 = heaviest under causal)
    const uint32_t bhk = blockIdx.x - jb * per;                      // b * Hkv + hk
    const uint32_t b = bhk / p.Hkv, hk = bhk - b * p.Hkv;
    const uint32_t group = p.Hq / p.Hkv;
    const uint32_t key0 = jb * 128;
    const uint32_t nqb = (p.Sq + 127) / 128;
    const uint32_t i_begin = p.causal ? jb : 0;                      // query blocks with rows >= key0 (top-left causal)
    const uint32_t steps_per_head = (i_begin < nqb) ? (nqb - i_begin) : 0;
    const uint32_t nsteps = steps_per_head * group;

    if (warp == 8) {
        // ===================================================== issuer
        if (elect_one()) {
            constexpr uint64_t HI_K = smem_desc_hi(16, 1024);                // K-major SW128
            constexpr uint64_t HI_MN = smem_desc_hi(C::CHUNK_BYTES, 1024);   // MN-major SW128 (64-wide chunks 16 KB apart)
            constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
            constexpr uint32_t HI_MN_HI = uint32_t(HI_MN >> 32), HI_MN_LO = uint32_t(HI_MN);
            auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
            constexpr uint32_t ID_KK = instr_desc_f16(BF16, 128, 128, false);                    // A K-major, B K-major, N=128
            constexpr uint32_t ID_MNMN = instr_desc_f16(BF16, 128, D, true) | (1u << 15);        // A MN-major, B MN-major, N=D
            constexpr uint32_t ID_KMN = instr_desc_f16(BF16, 128, D, true);                      // A K-major,  B MN-major, N=D
            auto coords = [&](uint32_t step, int32_t& row, int32_t& bh) {
                const uint32_t g = step / steps_per_head, i = i_begin + (step - g * steps_per_head);
                row = (int32_t)(i * 128);
                bh = (int32_t)(b * p.Hq + hk * group + g);
            };
            auto load_q = [&](uint32_t step) {
                int32_t row, bh; coords(step, row, bh);
                const uint32_t bar = bar_q0 + 8 * (step & 1), dst = sQ0 + (step & 1) * C::TILE_BYTES;
                mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                for (int c = 0; c < C::CHUNKS; ++c) tma_load_3d(dst + c * C::CHUNK_BYTES, tmQ, bar, c * 64, row, bh);
            };
            auto load_do = [&](uint32_t step) {
                int32_t row, bh; coords(step, row, bh);
                mbar_expect_tx(bar_do, C::TILE_BYTES);
#pragma unroll
                for (int c = 0; c < C::CHUNKS; ++c) tma_load_3d(sdO + c * C::CHUNK_BYTES, tmdO, bar_do, c * 64, row, bh);
            };
            auto issue_s = [&](uint32_t step) {                              // S = Q K^T
                const uint32_t sQ = sQ0 + (step & 1) * C::TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + COL_S, mk(HI_K_HI, (HI_K_LO | (sQ >> 4)) + off), mk(HI_K_HI, (HI_K_LO | (sK >> 4)) + off), ID_KK, kk > 0);
                }
            };
            mbar_expect_tx(bar_kv, 2 * C::TILE_BYTES);
#pragma unroll
            for (int c = 0; c < C::CHUNKS; ++c) {
                tma_load_3d(sK + c * C::CHUNK_BYTES, tmK, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
                tma_load_3d(sV + c * C::CHUNK_BYTES, tmV, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
            }
            if (nsteps > 0) { load_q(0); load_do(0); }
            if (nsteps > 1) load_q(1);
            mbar_wait(bar_kv, 0);
            if (nsteps > 0) {
                mbar_wait(bar_q0, 0);
                tc_fence_after();
                issue_s(0);
            }
            for (uint32_t step = 0; step < nsteps; ++step) {
                const uint32_t sQ = sQ0 + (step & 1) * C::TILE_BYTES;
                // ---- dP = dO V^T (needs dO(i) and the dP/dQ columns drained by the compute warps)
                mbar_wait(bar_do, step & 1);
                if (step > 0) mbar_wait(bar_dqf, (step - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + COL_DP, mk(HI_K_HI, (HI_K_LO | (sdO >> 4)) + off), mk(HI_K_HI, (HI_K_LO | (sV >> 4)) + off), ID_KK, kk > 0);
                }
                mma_commit(bar_sdp);
                // ---- dV += P^T dO, dK += dS^T Q, dQ = dS K
                mbar_wait(bar_pds, step & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)                                // K = 128 query rows, 16 per step
                    mma_ss(tmem + COL_DV, mk(HI_MN_HI, (HI_MN_LO | (sP >> 4)) + kk * 128), mk(HI_MN_HI, (HI_MN_LO | (sdO >> 4)) + kk * 128),
                           ID_MNMN, (step > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    mma_ss(tmem + COL_DK, mk(HI_MN_HI, (HI_MN_LO | (sdS >> 4)) + kk * 128), mk(HI_MN_HI, (HI_MN_LO | (sQ >> 4)) + kk * 128),
                           ID_MNMN, (step > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {                              // K = 128 keys; dQ lands in dP's columns
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + COL_DP, mk(HI_K_HI, (HI_K_LO | (sdS >> 4)) + off), mk(HI_MN_HI, (HI_MN_LO | (sK >> 4)) + kk * 128), ID_KMN, kk > 0);
                }
                mma_commit(bar_d);
                // ---- S(i+1) queues right behind (S columns were consumed before bar_pds; Q(i+1) has its own buffer)
                if (step + 1 < nsteps) {
                    mbar_wait(bar_q0 + 8 * ((step + 1) & 1), ((step + 1) >> 1) & 1);
                    tc_fence_after();
                    issue_s(step + 1);
                }
                // ---- buffers of step i are free once its MMAs are complete
                mbar_wait(bar_d, step & 1);
                if (step + 1 < nsteps) load_do(step + 1);
                if (step + 2 < nsteps) load_q(step + 2);
            }
        }
    } else {
        // ===================================================== compute warps
        const uint32_t h = warp >> 2;                                // key half: columns [64h, 64h+64)
        const uint32_t r = (warp & 3) * 32 + lane;                   // row of the tile == TMEM lane
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        // row statistics of the NEXT step are fetched one step ahead (their global-load latency was ~20 % of the kernel)
        float lse2_n = 0.f, delta_n = 0.f;
        auto fetch_stats = [&](uint32_t step) {
            const uint32_t g = step / steps_per_head, i = i_begin + (step - g * steps_per_head);
            const uint32_t row = i * 128 + r;
            const size_t off = ((size_t)b * p.Hq + hk * group + g) * p.Sq;
            const bool ok = row < p.Sq;
            lse2_n = ok ? p.lse[off + row] * 1.4426950408889634f : 0.f;
            delta_n = ok ? p.delta[off + row] : 0.f;
        };
        if (nsteps > 0) fetch_stats(0);
        for (uint32_t step = 0; step < nsteps; ++step) {
            const uint32_t g = step / steps_per_head, i = i_begin + (step - g * steps_per_head);
            const uint32_t hq = hk * group + g;
            const uint32_t row = i * 128 + r;                        // global query row of this thread
            const size_t stat_off = ((size_t)b * p.Hq + hq) * p.Sq;
            const bool row_ok = row < p.Sq;
            const float lse2 = lse2_n, delta = delta_n;
            if (step + 1 < nsteps) fetch_stats(step + 1);
            const bool diag = p.causal && (i * 128 < key0 + 128);    // block touches the diagonal

            // ---- P = exp2(S*scale_log2 - LSE*log2e), dS = P o (dP - Delta): 16-bit into swizzled SMEM
            mbar_wait(bar_sdp, step & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t s[32], dp[32];
                tmem_ld32(tmem + lane_addr + COL_S + 64 * h + c * 32, s);
                tmem_ld32(tmem + lane_addr + COL_DP + 64 * h + c * 32, dp);
                tmem_wait_ld();
                uint32_t pp[16], ds[16];
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float p0 = ex2(fmaf(__uint_as_float(s[e]), p.scale_log2, -lse2));
                    float p1 = ex2(fmaf(__uint_as_float(s[e + 1]), p.scale_log2, -lse2));
                    const uint32_t col0 = key0 + 64 * h + c * 32 + e;
                    const bool v0 = row_ok && col0 < p.Sk && !(diag && col0 > row);
                    const bool v1 = row_ok && col0 + 1 < p.Sk && !(diag && col0 + 1 > row);
                    p0 = v0 ? p0 : 0.f;
                    p1 = v1 ? p1 : 0.f;
                    const float d0 = p0 * (__uint_as_float(dp[e]) - delta);
                    const float d1 = p1 * (__uint_as_float(dp[e + 1]) - delta);
                    pp[e / 2] = pack2<BF16>(p0, p1);
                    ds[e / 2] = pack2<BF16>(d0, d1);
                }
                const uint32_t rowoff = h * C::CHUNK_BYTES + r * 128;            // chunk h = keys [64h, 64h+64)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t a = rowoff + (((c * 4 + u) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + a), "r"(pp[4 * u]), "r"(pp[4 * u + 1]), "r"(pp[4 * u + 2]), "r"(pp[4 * u + 3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdS + a), "r"(ds[4 * u]), "r"(ds[4 * u + 1]), "r"(ds[4 * u + 2]), "r"(ds[4 * u + 3]) : "memory");
                }
            }
            fence_proxy_async_smem();                                // generic-proxy writes -> visible to the MMA (async proxy)
            tc_fence_before();
            mbar_arrive(bar_pds);

            // ---- dQ partial (dP's columns, this half's D/2 of them) -> fp32 workspace, 16-byte vector reductions
            mbar_wait(bar_d, step & 1);
            tc_fence_after();
            {
                // tcgen05.ld is warp-collective (.sync.aligned): every lane executes it, only the reductions are
                // predicated on the row being in range (ragged last query block).
                float* dst = p.dq_ws + (stat_off + (row_ok ? row : 0)) * D + (D / 2) * h;
#pragma unroll 1
                for (int c = 0; c < D / 64; ++c) {
                    // (draining all columns first and issuing the reductions back-to-back was measured 3.5x SLOWER:
                    //  the burst of 16 red.v4 per thread then sits in front of the next step's proxy fence)
                    uint32_t q[32];
                    tmem_ld32(tmem + lane_addr + COL_DP + (D / 2) * h + c * 32, q);
                    tmem_wait_ld();
                    if (row_ok) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * 32 + e), "f"(__uint_as_float(q[e])),
                                         "f"(__uint_as_float(q[e + 1])), "f"(__uint_as_float(q[e + 2])), "f"(__uint_as_float(q[e + 3])) : "memory");
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_dqf);
        }
    }

    // ---- epilogue: dV, dK (x scale) -> 16-bit -> swizzled SMEM (P / dS buffers) -> TMA store
    __syncthreads();                                                 // every MMA is complete (the last bar_d was waited on)
    tc_fence_after();
    if (warp < 8) {
        const uint32_t h = warp >> 2, r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const uint32_t col = (which ? COL_DK : COL_DV) + (D / 2) * h;
            const uint32_t sbuf = which ? sdS : sP;
            const float mul = which ? p.scale : 1.f;
#pragma unroll 1
            for (int c = 0; c < D / 64; ++c) {
                uint32_t o[32];
                if (nsteps > 0) {
                    tmem_ld32(tmem + lane_addr + col + c * 32, o);
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = 0u;           // no visible query touches this KV block
                }
                const uint32_t dcol = (D / 2) * h + c * 32;          // first output column of this chunk
                const uint32_t chunk = dcol / 64, unit0 = (dcol % 64) / 8;
                const uint32_t rowbase = sbuf + chunk * C::CHUNK_BYTES + r * 128;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t v0 = pack2<BF16>(__uint_as_float(o[8 * u + 0]) * mul, __uint_as_float(o[8 * u + 1]) * mul);
                    const uint32_t v1 = pack2<BF16>(__uint_as_float(o[8 * u + 2]) * mul, __uint_as_float(o[8 * u + 3]) * mul);
                    const uint32_t v2 = pack2<BF16>(__uint_as_float(o[8 * u + 4]) * mul, __uint_as_float(o[8 * u + 5]) * mul);
                    const uint32_t v3 = pack2<BF16>(__uint_as_float(o[8 * u + 6]) * mul, __uint_as_float(o[8 * u + 7]) * mul);
                    const uint32_t addr = rowbase + (((unit0 + u) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
                }
            }
        }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < C::CHUNKS; ++c) {
            tma_store_3d(tmdV, sP + c * C::CHUNK_BYTES, c * 64, (int32_t)key0, (int32_t)bhk);
            tma_store_3d(tmdK, sdS + c * C::CHUNK_BYTES, c * 64, (int32_t)key0, (int32_t)bhk);
        }
        tma_store_commit();
        tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace bwd100

#define AULE_BWD100(NAME, DD, BF)                                                                        \
    extern "C" __global__ void __launch_bounds__(288, 1) NAME(const __grid_constant__ CUtensorMap tmQ,    \
                                                              const __grid_constant__ CUtensorMap tmK,    \
                                                              const __grid_constant__ CUtensorMap tmV,    \
                                                              const __grid_constant__ CUtensorMap tmdO,   \
                                                              const __grid_constant__ CUtensorMap tmdK,   \
                                                              const __grid_constant__ CUtensorMap tmdV,   \
                                                              const aule_kp::BwdParams p) {               \
        bwd100::bwd_body<DD, BF>(&tmQ, &tmK, &tmV, &tmdO, &tmdK, &tmdV, p);                               \
    }

AULE_BWD100(aule_bwd_sm100_bf16_d128, 128, true)
AULE_BWD100(aule_bwd_sm100_bf16_d64, 64, true)
AULE_BWD100(aule_bwd_sm100_f16_d128, 128, false)
AULE_BWD100(aule_bwd_sm100_f16_d64, 64, false)

// Delta_i = sum_d O_id dO_id (triton_flash.py:353-379), one warp per row; also zeroes nothing else.
template <typename T>
__device__ __forceinline__ void delta_body(const T* o, const T* d_o, float* delta, uint64_t rows, uint32_t D) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const uint32_t lane = threadIdx.x & 31;
    float acc = 0.f;
    for (uint32_t d = lane * 2; d < D; d += 64) {
        const float2 a = make_float2((float)o[row * D + d], (float)o[row * D + d + 1]);
        const float2 g = make_float2((float)d_o[row * D + d], (float)d_o[row * D + d + 1]);
        acc = fmaf(a.x, g.x, fmaf(a.y, g.y, acc));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) delta[row] = acc;
}
extern "C" __global__ void aule_bwd_delta_bf16(const __nv_bfloat16* o, const __nv_bfloat16* d_o, float* delta, uint64_t rows, uint32_t D) {
    delta_body(o, d_o, delta, rows, D);
}
extern "C" __global__ void aule_bwd_delta_f16(const __half* o, const __half* d_o, float* delta, uint64_t rows, uint32_t D) {
    delta_body(o, d_o, delta, rows, D);
}
// dQ = (dtype)(scale * workspace)
extern "C" __global__ void aule_bwd_dq_convert_bf16(const float* ws, __nv_bfloat16* dq, uint64_t n, float scale) {
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (uint64_t)gridDim.x * blockDim.x * 2)
        *reinterpret_cast<__nv_bfloat162*>(dq + i) = __floats2bfloat162_rn(ws[i] * scale, ws[i + 1] * scale);
}
extern "C" __global__ void aule_bwd_dq_convert_f16(const float* ws, __half* dq, uint64_t n, float scale) {
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (uint64_t)gridDim.x * blockDim.x * 2)
        *reinterpret_cast<__half2*>(dq + i) = __floats2half2_rn(ws[i] * scale, ws[i + 1] * scale);
}
