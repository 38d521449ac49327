// Attention backward for sm_100a (B200), fused form in 64-query half steps: ONE kernel computes dK, dV and dQ, and the
// dQ partials leave the SM through asynchronous TMA bulk reductions.                          (kernel v6, EXPERIMENTAL, opt-in)
//
// Status: parity-green (tests/test_gpu_r2.py::test_fused2_backward_kernel_vs_oracle) but measured SLOWER than the shipped
// two-kernel backward (config C/2: 1.71-1.79 ms against 1.59-1.64 ms in the same runs, profiles/r2_s2_fused2.md), so it is
// reached only through aule_set_kernel_path bit 26.  It is kept because it is the first form in this repository that
// issues 5 GEMMs per (query block, key block) pair AND gets the dQ partials out without stalling the compute warps;
// what still limits it is written down in profiles/r2_s2_fused2.md.
//
// Replaces (behaviour, not code): python/aule/triton_flash.py:242-347 (_flash_attn_bwd_kernel).  Like the reference
// (tl.atomic_add, :335-339) dQ is accumulated across key blocks with floating-point additions in global memory, so dQ is
// not bit-reproducible run to run (dK / dV are); the deterministic two-kernel backward of attn_bwd_sm100.cu is the default.
//
// Why this form.  The two-kernel backward issues 7 GEMMs and two exp passes per pair for 5 counted; the first fused kernel
// (attn_bwd_fused_sm100.cu) issues 5 and one but drains every dQ partial with red.global from the compute warps' registers,
// which an SM sustains at only 23-25 B/clk (tools/microbench/red_rate.cu): 2900 cycles of a 6100-cycle step in which those
// warps do nothing else.  The TMA form (cp.reduce.async.bulk) runs at the same rate but asynchronously -- it needs the
// partial in shared memory, though, and at 128-query steps shared memory is full.  Halving the query tile halves the
// Q / dO stages and the dS^T tile, which pays for two 32 KB staging tiles:
//
//   CTA = (key block j of 128 keys, kv head, batch); K_j, V_j resident; half step h = 64 queries of (q-head g, block i).
//     S^T  = K_j Q_h^T           SS  M=128 keys, N=64   -> TMEM [0,64)     -> P^T = exp2(S^T c - LSE) packed in place
//     dP^T = V_j dO_h^T          SS                     -> TMEM [64,128) / [128,192) (even / odd h)
//     dV  += P^T dO_h            TS  N=128, K=64        -> TMEM [256,384)
//     dS^T = P^T o (dP^T - Delta) -> 16-bit -> SMEM tile [128 keys][64 q]
//     dK  += dS^T Q_h            SS  (A = tile K-major, B = Q_h MN-major)              -> TMEM [384,512)
//     dQ_h^T = K_j^T dS^T        SS  (A = K_j MN-major, B = tile MN-major) M=128 d, N=64 -> TMEM [192,256)
//     dQ_h^T -> registers -> staging tile [64 q][128 d] fp32 (lane = d: conflict-free) -> bulk reductions (4 KB pieces)
//       into the fp32 accumulator [B,Hq,Sq,D] (64 consecutive query rows are contiguous there), issued by a reducer warp.
//   P / dS warps per half step:  dS(h) | P(h+1);   drain warps: dQ^T(h);   tensor pipe:  dP^T(h+1) | dQ^T(h) dK(h) | dV(h+1) | S^T(h+2).
//   Warps: 0-15 P / dS (thread = key row x 16 query columns), 16-19 drain (thread = d x 64 query columns), 20 = MMA +
//   TMA-load issuer, 21 = reducer, 22 = optional fence helper.  The drain is a latency chain of its own (wait, tcgen05.ld,
//   64 stores, proxy fence): on separate warps it runs beside the P / dS chain instead of inside it (first version, 16
//   symmetric warps: 2630 cycles per half step).  Variant with the drain warps issuing red.global themselves (no staging,
//   4-deep rings): 2.38 ms -- four warps cannot keep the reduction path busy (experiments/attn_bwd_fused2_red_global_drain.cu.txt).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sm100_ptx.cuh"
#include "kernel_params.h"

namespace bwd100f2 {
using namespace sm100;
using aule_kp::BwdParams;
using bwd100::Tracer;

__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
                 : "memory");
}

template <bool BF16>
__device__ __forceinline__ void bwd_fused2_body(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                                const CUtensorMap* tmdO, const CUtensorMap* tmdK, const CUtensorMap* tmdV,
                                                const BwdParams& p) {
    constexpr int D = 128;
    using C = aule_kp::BwdF2Cfg<D>;
    constexpr int NQ = C::NQ, NDO = C::NDO;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_kv = sb + C::OFF_BAR;            // K_j, V_j landed (once)
    const uint32_t bar_qfull0 = bar_kv + 8;             // Q stage s landed (+8s)
    const uint32_t bar_qfree0 = bar_qfull0 + 8 * NQ;    // dK(h) complete (commit): Q stage free (+8s)
    const uint32_t bar_dofull0 = bar_qfree0 + 8 * NQ;   // dO stage s landed (+8s)
    const uint32_t bar_dofree0 = bar_dofull0 + 8 * NDO; // dV(h) complete (commit): dO stage free (+8s)
    const uint32_t bar_s = bar_dofree0 + 8 * NDO;       // S^T(h) complete (commit)
    const uint32_t bar_dp = bar_s + 8;                  // dP^T(h) complete (commit)
    const uint32_t bar_dq = bar_dp + 8;                 // dQ^T(h) complete (commit)
    const uint32_t bar_dsfree = bar_dq + 8;             // dQ^T(h) and dK(h) complete (commit): the dS^T tile may be rewritten
    const uint32_t bar_p = bar_dsfree + 8;              // P / dS warps -> issuer: P^T(h) in TMEM (16 arrivals)
    const uint32_t bar_ds = bar_p + 8;                  // -> issuer: dS^T(h) in SMEM, dP^T(h) in registers (16 P / dS warps, or 1: the fence helper)
    const uint32_t bar_dqfree = bar_ds + 8;             // drain warps -> issuer: dQ^T(h) in registers (4 arrivals)
    const uint32_t bar_done = bar_dqfree + 8;           // every MMA complete (commit)
    const uint32_t bar_stat0 = bar_done + 8;            // publishers -> everyone: statistics of block step m in buffer m&1 (4 arrivals) (+8)
    const uint32_t bar_stgfull0 = bar_stat0 + 16;       // drain warps -> reducer: staging tile h&1 written (4 arrivals) (+8)
    const uint32_t bar_stgfree0 = bar_stgfull0 + 16;    // reducer -> drain warps: the bulk reductions have read staging tile h&1 (+8)
    const uint32_t bar_dp1 = bar_stgfree0 + 16;         // dP^T(h) complete, odd h (dP^T is double-buffered: one barrier per buffer)
    const uint32_t bar_dsw = bar_dp1 + 8;               // P / dS warps -> fencer warp: dS^T(h) written to SMEM (16 arrivals)
    static_assert(8 * (1 + 2 * NQ + 2 * NDO + 8 + 2 + 2 + 2 + 2) <= C::BAR_BYTES, "barrier area too small");
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_SLOT);
    float* stat = reinterpret_cast<float*>(smem + C::OFF_STAT);      // [2][lse2 x128 | delta x128] per 128-query block step
    const uint32_t sK = sb + C::OFF_K, sV = sb + C::OFF_V, sQ0 = sb + C::OFF_Q, sdO0 = sb + C::OFF_DO, sdS = sb + C::OFF_DS;
    const uint32_t sStg0 = sb + C::OFF_STG;

    if (threadIdx.x == 0) {
        if (sb & 1023u) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }
        mbar_init(bar_kv, 1);
        for (int i = 0; i < NQ; ++i) { mbar_init(bar_qfull0 + 8 * i, 1); mbar_init(bar_qfree0 + 8 * i, 1); }
        for (int i = 0; i < NDO; ++i) { mbar_init(bar_dofull0 + 8 * i, 1); mbar_init(bar_dofree0 + 8 * i, 1); }
        mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_dq, 1); mbar_init(bar_dsfree, 1); mbar_init(bar_done, 1);
        mbar_init(bar_p, 16); mbar_init(bar_ds, (p.order & 32) ? 1 : 16); mbar_init(bar_dqfree, 4); mbar_init(bar_dsw, 16);
        mbar_init(bar_stat0, 4); mbar_init(bar_stat0 + 8, 4);
        mbar_init(bar_stgfull0, 4); mbar_init(bar_stgfull0 + 8, 4);
        mbar_init(bar_stgfree0, 1); mbar_init(bar_stgfree0 + 8, 1); mbar_init(bar_dp1, 1);
        fence_mbar_init();
        tma_prefetch_desc(tmQ); tma_prefetch_desc(tmK); tma_prefetch_desc(tmV); tma_prefetch_desc(tmdO);
    }
    constexpr uint32_t W_ISSUER = 20, W_REDUCER = 21, W_FENCER = 22;  // warps 0-15 P / dS, 16-19 drain
    if (warp == W_ISSUER) tmem_alloc<512>(sb + C::OFF_TMEM_SLOT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM: S^T [0,64) | dP^T even h [64,128) | dP^T odd h [128,192) | dQ^T [192,256) | dV [256,384) | dK [384,512).
    // dP^T is double-buffered so that dP^T(h+1) does not wait for "dP^T(h) in registers": it is issued a whole half step before
    // it is needed (single-buffered, the P / dS warps waited ~520 cycles per half step for it behind dQ^T(h) and dK(h)).
    constexpr uint32_t COL_S = 0, COL_DP = 64, COL_DQ = 192, COL_DV = 256, COL_DK = 384;

    // ---- which KV block.  CTAs come in runs of U (batch, kv-head) units; inside a run the key block is the slow index
    //      (block 0 = heaviest under causal first), the unit the fast one.
    const uint32_t per = p.Hkv * p.B;
    const uint32_t nkb = (p.Sk + 127) / 128;
    const uint32_t U = (p.units_per_run == 0 || p.units_per_run > per) ? per : p.units_per_run;
    const uint32_t run = blockIdx.x / (U * nkb);
    const uint32_t in_run = blockIdx.x - run * (U * nkb);
    const uint32_t u_here = min(U, per - run * U);                   // the last run may be short
    const uint32_t jb = in_run / u_here;
    const uint32_t bhk = run * U + (in_run - jb * u_here);           // b * Hkv + hk
    const uint32_t b = bhk / p.Hkv, hk = bhk - b * p.Hkv;
    const uint32_t group = p.Hq / p.Hkv;
    const uint32_t key0 = jb * 128;
    const uint32_t nqb = (p.Sq + 127) / 128;
    const uint32_t i_begin = p.causal ? jb : 0;                      // query blocks with rows >= key0 (top-left causal)
    const uint32_t blocks_per_head = (i_begin < nqb) ? (nqb - i_begin) : 0;
    const uint32_t nsteps = blocks_per_head * group * 2;             // half steps: (g, i, half) with half fastest
    // BwdParams::order bit 4 (A/B): the generic->async proxy fence after the shared-memory writes is executed once by the
    // CONSUMER thread (MMA issuer / reducer) after it has acquired the writers' mbarrier, instead of by every writer before
    // its arrive -- fence.proxy.async costs the writers ~400 cycles on their critical chain.
    const bool consumer_fence = (p.order & 16) != 0;
    // BwdParams::order bit 5: a helper warp (22) executes that fence for the dS^T tile -- the 16 writer warps arrive on bar_dsw
    // right after their stores, the helper acquires it, fences and arrives on bar_ds for the MMA issuer: the ~500-cycle fence
    // then sits neither in the P / dS chain nor in the issuer's instruction stream.
    const bool fencer = (p.order & 32) != 0;

    if (warp == W_ISSUER) {
        // ===================================================== issuer
        if (elect_one() && nsteps > 0) {
            constexpr uint64_t HI_K = smem_desc_hi(16, 1024);                          // K-major SW128
            constexpr uint64_t HI_MNK = smem_desc_hi(C::KV_CHUNK_BYTES, 1024);         // MN-major SW128, 64-wide chunks 16 KB apart (K_j)
            constexpr uint64_t HI_MNQ = smem_desc_hi(C::Q_CHUNK_BYTES, 1024);          // MN-major SW128, chunks 8 KB apart (Q / dO / dS^T tiles)
            constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
            constexpr uint32_t HI_MNK_HI = uint32_t(HI_MNK >> 32), HI_MNK_LO = uint32_t(HI_MNK);
            constexpr uint32_t HI_MNQ_HI = uint32_t(HI_MNQ >> 32), HI_MNQ_LO = uint32_t(HI_MNQ);
            auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
            constexpr uint32_t ID_S = instr_desc_f16(BF16, 128, 64, false);                    // A, B K-major, N = 64 queries
            constexpr uint32_t ID_KMN = instr_desc_f16(BF16, 128, D, true);                    // A K-major (or TMEM), B MN-major, N = D
            constexpr uint32_t ID_MNMN = instr_desc_f16(BF16, 128, 64, true) | (1u << 15);     // A, B MN-major, M = D, N = 64 queries
            // (q-head, query block, half) of the next Q / dO load
            uint32_t ql = 0, ql_st = 0, ql_use = 0, ql_g = 0, ql_i = i_begin, ql_h = 0;
            uint32_t dl = 0, dl_st = 0, dl_use = 0, dl_g = 0, dl_i = i_begin, dl_h = 0;
            // L2 prefetch cursors, PF tiles ahead of the loads: a Q / dO tile is requested only ~1.5 half steps before its MMA
            // (ring depth), which does not cover a DRAM miss under this kernel's L2 traffic (trace: S^T(h+2) waited ~1500
            // cycles for Q on every other half step)
            constexpr uint32_t PF = 4;
            uint32_t pq = 0, pq_g = 0, pq_i = i_begin, pq_h = 0, pd = 0, pd_g = 0, pd_i = i_begin, pd_h = 0;
            for (uint32_t t = 0; t < PF; ++t) {
                ++pq; if (++pq_h == 2) { pq_h = 0; if (++pq_i == nqb) { pq_i = i_begin; ++pq_g; } }
                ++pd; if (++pd_h == 2) { pd_h = 0; if (++pd_i == nqb) { pd_i = i_begin; ++pd_g; } }
            }
            auto pump = [&]() {
                if (ql < nsteps && (ql_use == 0 || mbar_test(bar_qfree0 + 8 * ql_st, (ql_use - 1) & 1))) {
                    const uint32_t bar = bar_qfull0 + 8 * ql_st, dst = sQ0 + ql_st * C::Q_TILE_BYTES;
                    mbar_expect_tx(bar, C::Q_TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        tma_load_3d(dst + c * C::Q_CHUNK_BYTES, tmQ, bar, c * 64, (int32_t)(ql_i * 128 + ql_h * 64), (int32_t)(b * p.Hq + hk * group + ql_g));
                    ++ql;
                    if (++ql_h == 2) { ql_h = 0; if (++ql_i == nqb) { ql_i = i_begin; ++ql_g; } }
                    if (++ql_st == NQ) { ql_st = 0; ++ql_use; }
                    if (pq < nsteps) {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            tma_prefetch_3d(tmQ, c * 64, (int32_t)(pq_i * 128 + pq_h * 64), (int32_t)(b * p.Hq + hk * group + pq_g));
                        ++pq; if (++pq_h == 2) { pq_h = 0; if (++pq_i == nqb) { pq_i = i_begin; ++pq_g; } }
                    }
                }
                if (dl < nsteps && (dl_use == 0 || mbar_test(bar_dofree0 + 8 * dl_st, (dl_use - 1) & 1))) {
                    const uint32_t bar = bar_dofull0 + 8 * dl_st, dst = sdO0 + dl_st * C::Q_TILE_BYTES;
                    mbar_expect_tx(bar, C::Q_TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        tma_load_3d(dst + c * C::Q_CHUNK_BYTES, tmdO, bar, c * 64, (int32_t)(dl_i * 128 + dl_h * 64), (int32_t)(b * p.Hq + hk * group + dl_g));
                    ++dl;
                    if (++dl_h == 2) { dl_h = 0; if (++dl_i == nqb) { dl_i = i_begin; ++dl_g; } }
                    if (++dl_st == NDO) { dl_st = 0; ++dl_use; }
                    if (pd < nsteps) {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            tma_prefetch_3d(tmdO, c * 64, (int32_t)(pd_i * 128 + pd_h * 64), (int32_t)(b * p.Hq + hk * group + pd_g));
                        ++pd; if (++pd_h == 2) { pd_h = 0; if (++pd_i == nqb) { pd_i = i_begin; ++pd_g; } }
                    }
                }
            };
            auto wait = [&](uint32_t bar, uint32_t parity) {                 // blocking wait that keeps the loads flowing
                while (!mbar_try_wait<0>(bar, parity)) pump();
            };
            auto issue_s = [&](uint32_t k) {                                 // S^T(k) = K_j Q^T  (A = K_j, B = Q as [n = query][k = d])
                const uint32_t sQ = sQ0 + (k % NQ) * C::Q_TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t offa = ((kk / 4) * C::KV_CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    const uint32_t offb = ((kk / 4) * C::Q_CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + COL_S, mk(HI_K_HI, (HI_K_LO | (sK >> 4)) + offa), mk(HI_K_HI, (HI_K_LO | (sQ >> 4)) + offb), ID_S, kk > 0);
                }
                mma_commit(bar_s);
            };
            auto issue_dp = [&](uint32_t k) {                                // dP^T(k) = V_j dO^T
                const uint32_t sdO = sdO0 + (k % NDO) * C::Q_TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t offa = ((kk / 4) * C::KV_CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    const uint32_t offb = ((kk / 4) * C::Q_CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + COL_DP + 64 * (k & 1), mk(HI_K_HI, (HI_K_LO | (sV >> 4)) + offa), mk(HI_K_HI, (HI_K_LO | (sdO >> 4)) + offb), ID_S, kk > 0);
                }
                mma_commit((k & 1) ? bar_dp1 : bar_dp);
            };
            auto issue_dv = [&](uint32_t k) {                                // dV += P^T(k) dO (K = 64 queries, A = P^T in TMEM: 8 columns per k-step)
                const uint32_t sdO = sdO0 + (k % NDO) * C::Q_TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    mma_ts(tmem + COL_DV, tmem + COL_S + 16 * kk, mk(HI_MNQ_HI, (HI_MNQ_LO | (sdO >> 4)) + kk * 128), ID_KMN, (k > 0 || kk > 0) ? 1u : 0u);
                mma_commit(bar_dofree0 + 8 * (k % NDO));
            };
            mbar_expect_tx(bar_kv, 2 * C::KV_TILE_BYTES);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                tma_load_3d(sK + c * C::KV_CHUNK_BYTES, tmK, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
                tma_load_3d(sV + c * C::KV_CHUNK_BYTES, tmV, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
            }
            for (int t = 0; t < NQ; ++t) pump();                             // fill both rings
            Tracer tr(p.trace, 0, true);
            wait(bar_kv, 0);
            wait(bar_qfull0, 0);
            tc_fence_after();
            issue_s(0);
            wait(bar_dofull0, 0);
            tc_fence_after();
            issue_dp(0);
            wait(bar_p, 0);
            tc_fence_after();
            issue_dv(0);
            if (nsteps > 1) {
                wait(bar_qfull0 + 8 * (1 % NQ), 0);
                tc_fence_after();
                issue_s(1);                                                  // overwrites P^T(0): after dV(0) in the pipe
            }
            for (uint32_t h = 0; h < nsteps; ++h) {
                const uint32_t qst = h % NQ;
                if (h + 1 < nsteps) {                                        // dP^T(h+1): its buffer held dP^T(h-1), consumed before bar_ds(h-1)
                    wait(bar_dofull0 + 8 * ((h + 1) % NDO), ((h + 1) / NDO) & 1);
                    tr.ev(12, h);
                    tc_fence_after();
                    issue_dp(h + 1);
                }
                tr.ev(10, h);
                wait(bar_ds, h & 1);                                         // dS^T(h) in SMEM, dP^T(h) consumed; dQ^T(h-1) was drained before
                tr.ev(11, h);
                if (consumer_fence) fence_proxy_async_smem();
                // dQ^T(h) overwrites the columns of dQ^T(h-1): the drain warps run on their own and must have it in registers
                // (without this wait a slow drain is lapped on bar_dq -- found as a hang under ncu's replay passes)
                if (h > 0) wait(bar_dqfree, (h - 1) & 1);
                tc_fence_after();
                {
                    const uint32_t sQ = sQ0 + qst * C::Q_TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)                           // dQ^T(h) = K_j^T dS^T (K = 128 keys)
                        mma_ss(tmem + COL_DQ, mk(HI_MNK_HI, (HI_MNK_LO | (sK >> 4)) + kk * 128), mk(HI_MNQ_HI, (HI_MNQ_LO | (sdS >> 4)) + kk * 128),
                               ID_MNMN, kk > 0);
                    mma_commit(bar_dq);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)                           // dK += dS^T Q (K = 64 queries, A = the dS^T tile K-major)
                        mma_ss(tmem + COL_DK, mk(HI_K_HI, (HI_K_LO | (sdS >> 4)) + kk * 2), mk(HI_MNQ_HI, (HI_MNQ_LO | (sQ >> 4)) + kk * 128),
                               ID_KMN, (h > 0 || kk > 0) ? 1u : 0u);
                    mma_commit(bar_qfree0 + 8 * qst);
                    mma_commit(bar_dsfree);
                }
                if (h + 1 < nsteps) {
                    const uint32_t k = h + 1;
                    wait(bar_p, k & 1);                                      // P^T(h+1) in TMEM
                    tr.ev(13, h);
                    tc_fence_after();
                    issue_dv(k);
                    if (h + 2 < nsteps) {
                        const uint32_t k2 = h + 2;
                        wait(bar_qfull0 + 8 * (k2 % NQ), (k2 / NQ) & 1);
                        tr.ev(16, h);
                        tc_fence_after();
                        issue_s(k2);                                         // overwrites P^T(h+1): after dV(h+1) in the pipe
                    }
                }
            }
            mma_commit(bar_done);
            wait(bar_done, 0);
        }
    } else if (warp == W_REDUCER) {
        // ===================================================== reducer: the bulk reductions of a half step
        if (elect_one() && nsteps > 0) {
            uint32_t g = 0, i = i_begin, half = 0;
            for (uint32_t h = 0; h < nsteps; ++h) {
                const uint32_t bsel = (C::STG_BUFS == 2) ? (h & 1) : 0u;
                mbar_wait(bar_stgfull0 + 8 * bsel, (C::STG_BUFS == 2) ? ((h >> 1) & 1) : (h & 1));            // the 4 drain warps stored (and proxy-fenced) their part
                if (consumer_fence) fence_proxy_async_smem();
                const uint32_t q0 = i * 128 + half * 64;
                const uint32_t rows = q0 < p.Sq ? min(64u, p.Sq - q0) : 0u;
                // The tile leaves in PIECE-row pieces, each a bulk group of its own, with at most two pieces in flight, so that a
                // Q / dO tile load never queues behind one 32 KB reduction (~1400 cycles at 23-25 B/clk) in the SM's TMA unit.
                // (Tile loads arrive 1500-3000 cycles late on every other half step -- profiles/r2_s2_fused2_tma_reduce_trace.txt;
                // the pieces did not change that, so the queueing is not only here.  Kept: it costs nothing measurable.)
                constexpr uint32_t PIECE = 8;                                // query rows per bulk reduction (4 KB)
                float* dst = p.dq_acc + (((size_t)b * p.Hq + hk * group + g) * p.Sq + q0) * D;
                for (uint32_t r0 = 0; r0 < rows; r0 += PIECE) {
                    const uint32_t nr = min(PIECE, rows - r0);
                    bulk_reduce_add_f32(dst + (size_t)r0 * D, sStg0 + bsel * C::STG_BYTES + r0 * (D * 4), nr * D * 4);
                    tma_store_commit();
                    tma_store_wait_read<1>();
                }
                tma_store_wait_read<0>();                                    // this tile has been read: the drain warps may refill it
                mbar_arrive(bar_stgfree0 + 8 * bsel);
                if (++half == 2) { half = 0; if (++i == nqb) { i = i_begin; ++g; } }
            }
            tma_store_wait_read<0>();                                        // the CTA must outlive the reads of its shared memory
        }
    } else if (warp == W_FENCER) {
        if (fencer && elect_one()) {
            for (uint32_t h = 0; h < nsteps; ++h) {
                mbar_wait(bar_dsw, h & 1);
                tc_fence_after();
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(bar_ds);
            }
        }
    } else if (warp >= 16) {
        // ===================================================== drain warps 16-19: thread == (head-dim index d, all 64 query columns)
        // dQ^T(h): TMEM -> registers -> staging tile [q][d] fp32 (consecutive lanes = consecutive d: conflict-free) -> the
        // reducer's bulk reduction.  Their wait / load / store / fence chain runs beside the P / dS chain of warps 0-15, not in it.
        const uint32_t r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const uint32_t tDQ = tmem + lane_addr + COL_DQ;
        Tracer tr(p.trace, 2, warp == 16 && lane == 0);
        for (uint32_t h = 0; h < nsteps; ++h) {
            tr.ev(30, h);
            mbar_wait(bar_dq, h & 1);
            tr.ev(26, h);
            tc_fence_after();
            const uint32_t bsel = (C::STG_BUFS == 2) ? (h & 1) : 0u;
            const uint32_t base = sStg0 + bsel * C::STG_BYTES + r * 4;
            uint32_t dq[32];
            tmem_ld32(tDQ, dq);
            tmem_wait_ld();
            tr.ev(28, h);
            if constexpr (C::STG_BUFS == 2) {
                if (h >= 2) mbar_wait(bar_stgfree0 + 8 * bsel, ((h >> 1) - 1) & 1);   // the reduction of half step h-2 has read this tile
            } else {
                if (h >= 1) mbar_wait(bar_stgfree0, (h - 1) & 1);                     // one tile: the reduction of half step h-1 has read it
            }
            tr.ev(29, h);
#pragma unroll
            for (int e = 0; e < 32; ++e)
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + e * (D * 4)), "r"(dq[e]) : "memory");
            tmem_ld32(tDQ + 32, dq);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_dqfree);
#pragma unroll
            for (int e = 0; e < 32; ++e)
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + (32 + e) * (D * 4)), "r"(dq[e]) : "memory");
            if (!consumer_fence) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_stgfull0 + 8 * bsel);
            tr.ev(27, h);
        }
    } else {
        // ===================================================== P / dS warps 0-15: thread == (key row, 16 query columns)
        const uint32_t qg = warp >> 2;                               // query columns [16qg, 16qg+16) of the half step
        const uint32_t r = (warp & 3) * 32 + lane;                   // TMEM lane == key row of S^T / dP^T
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const uint32_t key = key0 + r;
        const bool key_ok = key < p.Sk;
        const uint32_t tS = tmem + lane_addr + COL_S + 16 * qg, tDP = tmem + lane_addr + COL_DP + 16 * qg;
        const uint32_t ds_row = sdS + r * 128;                       // this thread's 32 bytes: 16-byte units (2qg + u) ^ (r&7)
        // ---- column statistics (LSE, Delta of the 128 queries of a block step m = h >> 1): warps 0-3 publish them ahead
        const uint32_t t128 = threadIdx.x;                           // < 128 for the publishers
        float lse_n = 0.f, delta_n = 0.f;
        uint32_t g_n = 0, i_n = i_begin;
        const uint32_t nblocks = nsteps >> 1;
        auto fetch_stats = [&]() {
            const uint32_t row = i_n * 128 + t128;
            const size_t off = ((size_t)b * p.Hq + hk * group + g_n) * p.Sq;
            const bool ok = row < p.Sq;
            lse_n = ok ? p.lse[off + row] : 0.f;
            delta_n = ok ? p.delta[off + row] : 0.f;
            if (++i_n == nqb) { i_n = i_begin; ++g_n; }
        };
        auto publish = [&](uint32_t m) {                             // statistics of block step m -> buffer m&1, one arrival per warp
            float* sn = stat + (m & 1) * 256;
            sn[t128] = lse_n * 1.4426950408889634f; sn[128 + t128] = delta_n;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_stat0 + 8 * (m & 1));
        };
        if (warp < 4 && nblocks > 0) {
            fetch_stats(); publish(0);
            if (nblocks > 1) { fetch_stats(); publish(1); }
            if (nblocks > 2) fetch_stats();                          // block step 2, published during half step 2
        }
        const float2 cc = make_float2(p.scale_log2, p.scale_log2);
        float pv[16];                                                // P^T of the half step whose dS^T comes next (fp32)
        uint32_t i_p = i_begin;                                      // query block of the next P phase
        Tracer tr(p.trace, 1, warp == 0 && lane == 0);
        // P phase of half step k
        auto p_phase = [&](uint32_t k) {
            const uint32_t m = k >> 1, half = k & 1;
            if (half == 0) mbar_wait(bar_stat0 + 8 * (m & 1), (m >> 1) & 1);    // statistics of block step m are visible
            const uint32_t qrow0 = i_p * 128 + 64 * half;            // first query of the half step
            const uint32_t q0 = qrow0 + 16 * qg;                     // first query of this thread's columns
            const bool diag = p.causal && (qrow0 < key0 + 128);
            const bool masked = diag || !key_ok || key0 + 128 > p.Sk || qrow0 + 64 > p.Sq;
            uint32_t alive = 0xffffu;
            if (masked) {
                const int64_t first = diag ? (int64_t)key - (int64_t)q0 : 0, last = (int64_t)p.Sq - 1 - (int64_t)q0;
                const uint32_t lo_m = first <= 0 ? 0xffffu : (first > 15 ? 0u : ((0xffffu << (int)first) & 0xffffu));
                const uint32_t hi_m = last >= 15 ? 0xffffu : (last < 0 ? 0u : (0xffffu >> (15 - (int)last)));
                alive = key_ok ? (lo_m & hi_m) : 0u;
            }
            if (half == 1) { if (++i_p == nqb) i_p = i_begin; }
            const float* sc = stat + (m & 1) * 256 + 64 * half + 16 * qg;        // this thread's 16 lse2
            mbar_wait(bar_s, k & 1);
            tc_fence_after();
            uint32_t sreg[16];
            tmem_ld16(tS, sreg);
            tmem_wait_ld();
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
                const float4 l4 = *reinterpret_cast<const float4*>(sc + 4 * e4);       // broadcast
                const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(sreg[4 * e4]), __uint_as_float(sreg[4 * e4 + 1])), cc, make_float2(-l4.x, -l4.y));
                const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(sreg[4 * e4 + 2]), __uint_as_float(sreg[4 * e4 + 3])), cc, make_float2(-l4.z, -l4.w));
                float2 v0, v1;
                if ((e4 & 1) == 0) { v0 = ex2_emu2(x0); } else { v0.x = ex2(x0.x); v0.y = ex2(x0.y); }   // 1 pair in 4 on the FMA pipe
                v1.x = ex2(x1.x); v1.y = ex2(x1.y);
                pv[4 * e4] = v0.x; pv[4 * e4 + 1] = v0.y; pv[4 * e4 + 2] = v1.x; pv[4 * e4 + 3] = v1.y;
            }
            if (masked) {
#pragma unroll
                for (int e = 0; e < 16; ++e) pv[e] = (alive & (1u << e)) ? pv[e] : 0.f;
            }
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) pk[e] = pack2<BF16>(pv[2 * e], pv[2 * e + 1]);
            tmem_st8(tS, pk);                                        // packed P^T over the first 8 of this thread's 16 columns
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p);
        };
        if (nsteps > 0) p_phase(0);
        for (uint32_t h = 0; h < nsteps; ++h) {
            const uint32_t m = h >> 1, half = h & 1;
            // ---- (a) dS^T(h) = P^T o (dP^T - delta[query]) -> 16-bit -> the shared-memory tile
            tr.ev(20, h);
            mbar_wait(half ? bar_dp1 : bar_dp, (h >> 1) & 1);
            tr.ev(21, h);
            tc_fence_after();
            {
                const float* sd = stat + (m & 1) * 256 + 128 + 64 * half + 16 * qg;       // this thread's 16 deltas
                uint32_t dp[16], pk[8];
                tmem_ld16(tDP + 64 * half, dp);
                tmem_wait_ld();
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    const float4 d4 = *reinterpret_cast<const float4*>(sd + 4 * e4);  // broadcast
                    const float2 a0 = __fmul2_rn(make_float2(pv[4 * e4], pv[4 * e4 + 1]),
                                                 __fadd2_rn(make_float2(__uint_as_float(dp[4 * e4]), __uint_as_float(dp[4 * e4 + 1])), make_float2(-d4.x, -d4.y)));
                    const float2 a1 = __fmul2_rn(make_float2(pv[4 * e4 + 2], pv[4 * e4 + 3]),
                                                 __fadd2_rn(make_float2(__uint_as_float(dp[4 * e4 + 2]), __uint_as_float(dp[4 * e4 + 3])), make_float2(-d4.z, -d4.w)));
                    pk[2 * e4] = pack2<BF16>(a0.x, a0.y);
                    pk[2 * e4 + 1] = pack2<BF16>(a1.x, a1.y);
                }
                tr.ev(22, h);
                if (h > 0) mbar_wait(bar_dsfree, (h - 1) & 1);       // dQ^T(h-1) and dK(h-1) have read the tile
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t a = ds_row + (((2 * qg + u) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * u]), "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3]) : "memory");
                }
                if (!consumer_fence && !fencer) fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(fencer ? bar_dsw : bar_ds);
                // dQ^T(h-1) / dK(h-1) complete means every warp finished dS(h-1): when h is even that was the last reader of
                // the statistics of block step m-1, whose buffer now takes block step m+1
                if (half == 0 && h >= 2 && warp < 4 && m + 1 < nblocks) {
                    publish(m + 1);
                    if (m + 2 < nblocks) fetch_stats();
                }
            }
            tr.ev(23, h);
            // ---- (b) P^T(h+1)
            if (h + 1 < nsteps) p_phase(h + 1);
            tr.ev(25, h);
        }
    }

    // ---- epilogue: dV, dK (x scale) -> 16-bit -> swizzled SMEM (Q stages 0-1 / dO stages 0-1, free now) -> TMA store
    __syncthreads();                                                 // every MMA is complete (the issuer waited on bar_done)
    tc_fence_after();
    if (warp < 16) {
        const uint32_t hh = (warp >> 2) & 1, r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const int which = warp >> 3;                                 // warps 0-7: dV, warps 8-15: dK
        const uint32_t col = (which ? COL_DK : COL_DV) + (D / 2) * hh;
        const uint32_t sbuf = which ? sdO0 : sQ0;
        const float mul = which ? p.scale : 1.f;
#pragma unroll 1
        for (int c = 0; c < D / 64; ++c) {
            uint32_t o[32];
            if (nsteps > 0) {
                tmem_ld32(tmem + lane_addr + col + c * 32, o);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) o[e] = 0u;               // no visible query touches this KV block
            }
            const uint32_t dcol = (D / 2) * hh + c * 32;             // first output column of this chunk
            const uint32_t chunk = dcol / 64, unit0 = (dcol % 64) / 8;
            const uint32_t rowbase = sbuf + chunk * C::KV_CHUNK_BYTES + r * 128;
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
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            tma_store_3d(tmdV, sQ0 + c * C::KV_CHUNK_BYTES, c * 64, (int32_t)key0, (int32_t)bhk);
            tma_store_3d(tmdK, sdO0 + c * C::KV_CHUNK_BYTES, c * 64, (int32_t)key0, (int32_t)bhk);
        }
        tma_store_commit();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ISSUER) tmem_dealloc<512>(tmem);
    if (threadIdx.x == 0) tma_store_wait_read<0>();                  // the CTA only has to outlive the reads of its shared memory
}

}  // namespace bwd100f2

#define AULE_BWD100_FUSED2(NAME, BF)                                                                     \
    extern "C" __global__ void __launch_bounds__(736, 1) NAME(const __grid_constant__ CUtensorMap tmQ,    \
                                                              const __grid_constant__ CUtensorMap tmK,    \
                                                              const __grid_constant__ CUtensorMap tmV,    \
                                                              const __grid_constant__ CUtensorMap tmdO,   \
                                                              const __grid_constant__ CUtensorMap tmdK,   \
                                                              const __grid_constant__ CUtensorMap tmdV,   \
                                                              const aule_kp::BwdParams p) {               \
        bwd100f2::bwd_fused2_body<BF>(&tmQ, &tmK, &tmV, &tmdO, &tmdK, &tmdV, p);                          \
    }
AULE_BWD100_FUSED2(aule_bwd_fused2_sm100_bf16_d128, true)
AULE_BWD100_FUSED2(aule_bwd_fused2_sm100_f16_d128, false)
