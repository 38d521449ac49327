// Attention backward for sm_100a (B200), fused form: ONE kernel computes dK, dV and dQ.                (kernel v5)
//
// Replaces (behaviour, not code): python/aule/triton_flash.py:242-347 (_flash_attn_bwd_kernel) -- like the reference it
// accumulates dQ across key blocks with floating-point reductions in global memory (tl.atomic_add, :335-339), so dQ is
// NOT bit-reproducible run to run (dK / dV are); the deterministic two-kernel backward of attn_bwd_sm100.cu stays
// available (aule_set_kernel_path bit 17 / AULE_DETERMINISTIC=1) and remains the path for D = 64.
// Why: the two-kernel form issues 7 GEMMs and two exp passes per (query block, key block) pair for 5 counted; this one
// issues 5 and one, and P / dS are computed once.
//
// CTA = (key block j of 128 keys, kv head, batch); K_j, V_j resident in SMEM; loop over the q-heads of the GQA group and
// the query blocks i at or below the diagonal.  Score tiles are TRANSPOSED (rows = TMEM lanes = keys, columns = queries):
//     S^T  = K_j Q_i^T          SS   -> TMEM [0,128)    -> P^T = exp2(S^T c - LSE_i) packed in place (A operand of dV)
//     dP^T = V_j dO_i^T         SS   -> TMEM [128,256)
//     dV  += P^T dO_i           TS   -> TMEM [256,384)
//     dS^T = P^T o (dP^T - Delta_i) -> 16-bit -> SMEM tile [128 keys][128 queries] (the only staged tile)
//     dK  += dS^T Q_i           SS (A = the dS^T tile K-major, B = Q_i MN-major)            -> TMEM [384,512)
//     dQ_i^T = K_j^T dS^T       SS (A = K_j MN-major, B = the dS^T tile MN-major)           -> TMEM [128,256) (dP^T is dead)
// dQ_i^T comes out with rows = head-dim index, columns = queries: a warp's 32 lanes hold 32 consecutive d of one query row,
// so every red.global.add.f32 of the drain is one fully coalesced 128-byte reduction (the row-major form would scatter a
// warp's reduction over 32 lines: 32x the LSU wavefronts).  The accumulator is fp32 [B,Hq,Sq,D], zeroed by the host and
// converted (x scale) by aule_bwd_dq_convert_*.  CTAs are ordered in runs of `units_per_run` (batch, kv-head) units,
// heaviest key block first inside a run, so that the accumulator rows a run reduces into stay L2-resident.
//
// Per step s the compute warps run   dS(s) | P(s+1) first half | drain dQ(s) | P(s+1) second half   and the tensor pipe
//   dQ^T(s) dK(s) | dV(s+1) even k-steps | dP^T(s+1) | dV(s+1) odd k-steps | S^T(s+2).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sm100_ptx.cuh"
#include "kernel_params.h"

namespace bwd100f {
using namespace sm100;
using aule_kp::BwdParams;
using bwd100::Tracer;

template <bool BF16>
__device__ __forceinline__ void bwd_fused_body(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                               const CUtensorMap* tmdO, const CUtensorMap* tmdK, const CUtensorMap* tmdV,
                                               const BwdParams& p) {
    constexpr int D = 128;
    using C = aule_kp::BwdFCfg<D>;
    constexpr int NQ = C::NQ, NDO = C::NDO;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_kv = sb + C::OFF_BAR;            // K_j, V_j landed (once)
    const uint32_t bar_qfull0 = bar_kv + 8;             // Q stage s landed (+8s)
    const uint32_t bar_qfree0 = bar_qfull0 + 8 * NQ;    // dK(i) complete (commit): Q stage free (+8s)
    const uint32_t bar_dofull0 = bar_qfree0 + 8 * NQ;   // dO stage s landed (+8s)
    const uint32_t bar_dofree0 = bar_dofull0 + 8 * NDO; // dV(i) complete (commit): dO stage free (+8s)
    const uint32_t bar_s = bar_dofree0 + 8 * NDO;       // S^T(i) complete (commit)
    const uint32_t bar_dp = bar_s + 8;                  // dP^T(i) complete (commit)
    const uint32_t bar_dq = bar_dp + 8;                 // dQ^T(i) complete (commit)
    const uint32_t bar_dsfree = bar_dq + 8;             // dQ^T(i) and dK(i) complete (commit): the dS^T tile may be rewritten
    const uint32_t bar_p = bar_dsfree + 8;              // compute -> issuer: first halves of P^T(i) in TMEM (16 arrivals)
    const uint32_t bar_pb = bar_p + 8;                  // compute -> issuer: second halves
    const uint32_t bar_ds = bar_pb + 8;                 // compute -> issuer: dS^T(i) in SMEM, dP^T(i) in registers (16)
    const uint32_t bar_dqfree = bar_ds + 8;             // compute -> issuer: dQ^T(i) in registers, its columns may take dP^T(i+1) (16)
    const uint32_t bar_done = bar_dqfree + 8;           // every MMA complete (commit)
    const uint32_t bar_stat0 = bar_done + 8;            // publishers -> everyone: statistics of step k are in buffer k&1 (4 arrivals) (+8)
    static_assert(8 * (1 + 2 * NQ + 2 * NDO + 9 + 2) <= C::BAR_BYTES, "barrier area too small");
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_SLOT);
    float* stat = reinterpret_cast<float*>(smem + C::OFF_STAT);      // [2][lse2 x128 | delta x128]
    const uint32_t sK = sb + C::OFF_K, sV = sb + C::OFF_V, sQ0 = sb + C::OFF_Q, sdO0 = sb + C::OFF_DO, sdS = sb + C::OFF_DS;

    if (threadIdx.x == 0) {
        if (sb & 1023u) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }
        mbar_init(bar_kv, 1);
        for (int i = 0; i < NQ; ++i) { mbar_init(bar_qfull0 + 8 * i, 1); mbar_init(bar_qfree0 + 8 * i, 1); }
        for (int i = 0; i < NDO; ++i) { mbar_init(bar_dofull0 + 8 * i, 1); mbar_init(bar_dofree0 + 8 * i, 1); }
        mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_dq, 1); mbar_init(bar_dsfree, 1); mbar_init(bar_done, 1);
        mbar_init(bar_p, 16); mbar_init(bar_pb, 16); mbar_init(bar_ds, 16); mbar_init(bar_dqfree, 16);
        mbar_init(bar_stat0, 4); mbar_init(bar_stat0 + 8, 4);
        fence_mbar_init();
        tma_prefetch_desc(tmQ); tma_prefetch_desc(tmK); tma_prefetch_desc(tmV); tma_prefetch_desc(tmdO);
    }
    if (warp == 16) tmem_alloc<512>(sb + C::OFF_TMEM_SLOT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 384;

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
    const uint32_t steps_per_head = (i_begin < nqb) ? (nqb - i_begin) : 0;
    const uint32_t nsteps = steps_per_head * group;

    if (warp == 16) {
        // ===================================================== issuer
        if (elect_one() && nsteps > 0) {
            constexpr uint64_t HI_K = smem_desc_hi(16, 1024);                // K-major SW128
            constexpr uint64_t HI_MN = smem_desc_hi(C::CHUNK_BYTES, 1024);   // MN-major SW128 (64-wide chunks 16 KB apart)
            constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
            constexpr uint32_t HI_MN_HI = uint32_t(HI_MN >> 32), HI_MN_LO = uint32_t(HI_MN);
            auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
            constexpr uint32_t ID_KK = instr_desc_f16(BF16, 128, 128, false);                  // A, B K-major, N = 128 queries
            constexpr uint32_t ID_KMN = instr_desc_f16(BF16, 128, D, true);                    // A K-major (or TMEM), B MN-major, N = D
            constexpr uint32_t ID_MNMN = instr_desc_f16(BF16, 128, 128, true) | (1u << 15);    // A, B MN-major, M = D, N = 128 queries
            uint32_t ql = 0, ql_st = 0, ql_use = 0, ql_g = 0, ql_i = i_begin;
            uint32_t dl = 0, dl_st = 0, dl_use = 0, dl_g = 0, dl_i = i_begin;
            auto pump = [&]() {
                if (ql < nsteps && (ql_use == 0 || mbar_test(bar_qfree0 + 8 * ql_st, (ql_use - 1) & 1))) {
                    const uint32_t bar = bar_qfull0 + 8 * ql_st, dst = sQ0 + ql_st * C::TILE_BYTES;
                    mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        tma_load_3d(dst + c * C::CHUNK_BYTES, tmQ, bar, c * 64, (int32_t)(ql_i * 128), (int32_t)(b * p.Hq + hk * group + ql_g));
                    ++ql;
                    if (++ql_i == nqb) { ql_i = i_begin; ++ql_g; }
                    if (++ql_st == NQ) { ql_st = 0; ++ql_use; }
                    if (ql < nsteps) {                                       // the tile after this one: warm it in L2 (2-stage ring)
#pragma unroll
                        for (int c = 0; c < C::CHUNKS; ++c)
                            tma_prefetch_3d(tmQ, c * 64, (int32_t)(ql_i * 128), (int32_t)(b * p.Hq + hk * group + ql_g));
                    }
                }
                if (dl < nsteps && (dl_use == 0 || mbar_test(bar_dofree0 + 8 * dl_st, (dl_use - 1) & 1))) {
                    const uint32_t bar = bar_dofull0 + 8 * dl_st, dst = sdO0 + dl_st * C::TILE_BYTES;
                    mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        tma_load_3d(dst + c * C::CHUNK_BYTES, tmdO, bar, c * 64, (int32_t)(dl_i * 128), (int32_t)(b * p.Hq + hk * group + dl_g));
                    ++dl;
                    if (++dl_i == nqb) { dl_i = i_begin; ++dl_g; }
                    if (++dl_st == NDO) { dl_st = 0; ++dl_use; }
                }
            };
            auto wait = [&](uint32_t bar, uint32_t parity) {                 // blocking wait that keeps the loads flowing
                while (!mbar_try_wait<0>(bar, parity)) pump();
            };
            auto issue_s = [&](uint32_t k) {                                 // S^T(k) = K_j Q^T  (A = K_j, B = Q as [n = query][k = d])
                const uint32_t sQ = sQ0 + (k % NQ) * C::TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + COL_S, mk(HI_K_HI, (HI_K_LO | (sK >> 4)) + off), mk(HI_K_HI, (HI_K_LO | (sQ >> 4)) + off), ID_KK, kk > 0);
                }
                mma_commit(bar_s);
            };
            auto issue_dp = [&](uint32_t k) {                                // dP^T(k) = V_j dO^T
                const uint32_t sdO = sdO0 + (k % NDO) * C::TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + COL_DP, mk(HI_K_HI, (HI_K_LO | (sV >> 4)) + off), mk(HI_K_HI, (HI_K_LO | (sdO >> 4)) + off), ID_KK, kk > 0);
                }
                mma_commit(bar_dp);
            };
            auto issue_dv_half = [&](uint32_t k, int half) {                 // dV += P^T(k) dO (K = 128 queries, A = P^T in TMEM): k-steps of one parity
                const uint32_t sdO = sdO0 + (k % NDO) * C::TILE_BYTES;
#pragma unroll
                for (int kk = half; kk < 8; kk += 2)
                    mma_ts(tmem + COL_DV, tmem + COL_S + 32 * (kk >> 1) + 8 * (kk & 1),
                           mk(HI_MN_HI, (HI_MN_LO | (sdO >> 4)) + kk * 128), ID_KMN, (k > 0 || kk > 0) ? 1u : 0u);
                if (half) mma_commit(bar_dofree0 + 8 * (k % NDO));
            };
            mbar_expect_tx(bar_kv, 2 * C::TILE_BYTES);
#pragma unroll
            for (int c = 0; c < C::CHUNKS; ++c) {
                tma_load_3d(sK + c * C::CHUNK_BYTES, tmK, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
                tma_load_3d(sV + c * C::CHUNK_BYTES, tmV, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
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
            issue_dv_half(0, 0);
            wait(bar_pb, 0);
            tc_fence_after();
            issue_dv_half(0, 1);
            if (nsteps > 1) {
                wait(bar_qfull0 + 8 * (1 % NQ), 0);
                tc_fence_after();
                issue_s(1);
            }
            for (uint32_t s = 0; s < nsteps; ++s) {
                const uint32_t qst = s % NQ;
                tr.ev(10, s);
                wait(bar_ds, s & 1);                                         // dS^T(s) in SMEM, dP^T(s) consumed
                tr.ev(11, s);
                tc_fence_after();
                {
                    const uint32_t sQ = sQ0 + qst * C::TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)                           // dQ^T(s) = K_j^T dS^T (K = 128 keys) over the dP^T columns
                        mma_ss(tmem + COL_DP, mk(HI_MN_HI, (HI_MN_LO | (sK >> 4)) + kk * 128), mk(HI_MN_HI, (HI_MN_LO | (sdS >> 4)) + kk * 128),
                               ID_MNMN, kk > 0);
                    mma_commit(bar_dq);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {                         // dK += dS^T Q (K = 128 queries, A = the dS^T tile K-major)
                        const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                        mma_ss(tmem + COL_DK, mk(HI_K_HI, (HI_K_LO | (sdS >> 4)) + off), mk(HI_MN_HI, (HI_MN_LO | (sQ >> 4)) + kk * 128),
                               ID_KMN, (s > 0 || kk > 0) ? 1u : 0u);
                    }
                    mma_commit(bar_qfree0 + 8 * qst);
                    mma_commit(bar_dsfree);
                }
                if (s + 1 < nsteps) {
                    const uint32_t k = s + 1;
                    wait(bar_dofull0 + 8 * (k % NDO), (k / NDO) & 1);
                    tr.ev(12, s);
                    wait(bar_p, k & 1);                                      // first halves of P^T(s+1)
                    tr.ev(13, s);
                    tc_fence_after();
                    issue_dv_half(k, 0);
                    wait(bar_dqfree, s & 1);                                 // dQ^T(s) drained
                    tr.ev(14, s);
                    tc_fence_after();
                    issue_dp(k);
                    wait(bar_pb, k & 1);                                     // second halves
                    tr.ev(15, s);
                    tc_fence_after();
                    issue_dv_half(k, 1);
                    if (s + 2 < nsteps) {
                        const uint32_t k2 = s + 2;
                        wait(bar_qfull0 + 8 * (k2 % NQ), (k2 / NQ) & 1);
                        tr.ev(16, s);
                        tc_fence_after();
                        issue_s(k2);                                         // overwrites P^T(s+1): after dV(s+1) in the pipe
                    }
                }
            }
            mma_commit(bar_done);
            wait(bar_done, 0);
        }
    } else {
        // ===================================================== compute warps: thread == (key row | d index, query quarter)
        const uint32_t qt = warp >> 2;                               // query quarter: columns [32qt, 32qt+32)
        const uint32_t r = (warp & 3) * 32 + lane;                   // TMEM lane: key row of S^T / dP^T, head-dim index of dQ^T
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const uint32_t key = key0 + r;
        const bool key_ok = key < p.Sk;
        const uint32_t tS = tmem + lane_addr + COL_S + 32 * qt, tDP = tmem + lane_addr + COL_DP + 32 * qt;
        const uint32_t ds_row = sdS + (qt >> 1) * C::CHUNK_BYTES + r * 128;      // this thread's 64 bytes: units ((qt&1)*4 + u) ^ (r&7)
        // ---- column statistics (LSE_i, Delta_i of the 128 queries of a step): warps 0-3 publish them two steps ahead
        const uint32_t t128 = threadIdx.x;                           // < 128 for the publishers
        float lse_n = 0.f, delta_n = 0.f;
        uint32_t g_n = 0, i_n = i_begin;
        auto fetch_stats = [&]() {
            const uint32_t row = i_n * 128 + t128;
            const size_t off = ((size_t)b * p.Hq + hk * group + g_n) * p.Sq;
            const bool ok = row < p.Sq;
            lse_n = ok ? p.lse[off + row] : 0.f;
            delta_n = ok ? p.delta[off + row] : 0.f;
            if (++i_n == nqb) { i_n = i_begin; ++g_n; }
        };
        auto publish = [&](uint32_t k) {                             // statistics of step k -> buffer k&1, one arrival per warp
            float* sn = stat + (k & 1) * 256;
            sn[t128] = lse_n * 1.4426950408889634f; sn[128 + t128] = delta_n;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_stat0 + 8 * (k & 1));
        };
        if (warp < 4 && nsteps > 0) {
            fetch_stats(); publish(0);
            if (nsteps > 1) { fetch_stats(); publish(1); }
            if (nsteps > 2) fetch_stats();                           // step 2, published during step 0
        }
        Tracer tr(p.trace, 1 + (qt & 1), (warp == 0 || warp == 4) && lane == 0);
        const float2 cc = make_float2(p.scale_log2, p.scale_log2);
        float pv[32];                                                // P^T of the step whose dS^T comes next (fp32)
        uint32_t i_p = i_begin;                                      // query block of the next P phase
        uint32_t i_d = i_begin, g_d = 0;                             // (query block, q-head of the group) of the next drain
        uint32_t alive = 0xffffffffu;
        bool masked = false;
        // P phase of step k, one half (16 query columns): publish, so that half of dV runs under the other half's math
        auto p_begin = [&](uint32_t k) {
            mbar_wait(bar_stat0 + 8 * (k & 1), (k >> 1) & 1);        // statistics of step k are visible
            const uint32_t q0 = i_p * 128 + 32 * qt;                 // first query of this thread's columns
            const bool diag = p.causal && (i_p * 128 < key0 + 128);
            masked = diag || !key_ok || key0 + 128 > p.Sk || i_p * 128 + 128 > p.Sq;
            alive = 0xffffffffu;
            if (masked) {
                const int64_t first = diag ? (int64_t)key - (int64_t)q0 : 0, last = (int64_t)p.Sq - 1 - (int64_t)q0;
                const uint32_t lo_m = first <= 0 ? 0xffffffffu : (first > 31 ? 0u : (0xffffffffu << (int)first));
                const uint32_t hi_m = last >= 31 ? 0xffffffffu : (last < 0 ? 0u : (0xffffffffu >> (31 - (int)last)));
                alive = key_ok ? (lo_m & hi_m) : 0u;
            }
            if (++i_p == nqb) i_p = i_begin;
            mbar_wait(bar_s, k & 1);
            tc_fence_after();
        };
        auto p_half = [&](uint32_t k, int half) {
            // (each half loads its own 16 S^T columns: the packed P^T of the first half only overwrites columns [0,8), and
            //  holding the second half's S^T in registers across the drain made the loop spill)
            const float* sc = stat + (k & 1) * 256 + 32 * qt + 16 * half;        // this half's 16 lse2
            uint32_t sreg[16];
            tmem_ld16(tS + 16 * half, sreg);
            tmem_wait_ld();
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
                const float4 l4 = *reinterpret_cast<const float4*>(sc + 4 * e4);       // broadcast
                const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(sreg[4 * e4]), __uint_as_float(sreg[4 * e4 + 1])), cc, make_float2(-l4.x, -l4.y));
                const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(sreg[4 * e4 + 2]), __uint_as_float(sreg[4 * e4 + 3])), cc, make_float2(-l4.z, -l4.w));
                float2 v0, v1;
                if ((e4 & 1) == 0) { v0 = ex2_emu2(x0); } else { v0.x = ex2(x0.x); v0.y = ex2(x0.y); }   // 1 pair in 4 on the FMA pipe
                v1.x = ex2(x1.x); v1.y = ex2(x1.y);
                pv[16 * half + 4 * e4] = v0.x; pv[16 * half + 4 * e4 + 1] = v0.y; pv[16 * half + 4 * e4 + 2] = v1.x; pv[16 * half + 4 * e4 + 3] = v1.y;
            }
            if (masked) {
#pragma unroll
                for (int e = 16 * half; e < 16 * half + 16; ++e) pv[e] = (alive & (1u << e)) ? pv[e] : 0.f;
            }
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) pk[e] = pack2<BF16>(pv[16 * half + 2 * e], pv[16 * half + 2 * e + 1]);
            tmem_st8(tS + 8 * half, pk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(half ? bar_pb : bar_p);
        };
        if (nsteps > 0) {
            p_begin(0);
            p_half(0, 0);
            p_half(0, 1);
        }
        for (uint32_t s = 0; s < nsteps; ++s) {
            // ---- (a) dS^T(s) = P^T o (dP^T - delta[query]) -> 16-bit -> the shared-memory tile
            tr.ev(20, s);
            mbar_wait(bar_dp, s & 1);
            tr.ev(21, s);
            tc_fence_after();
            {
                const float* sd = stat + (s & 1) * 256 + 128 + 32 * qt;       // this quarter's 32 deltas
                uint32_t dp[32], pk[16];
                tmem_ld32(tDP, dp);
                tmem_wait_ld();
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 d4 = *reinterpret_cast<const float4*>(sd + 4 * e4);  // broadcast
                    const float2 a0 = __fmul2_rn(make_float2(pv[4 * e4], pv[4 * e4 + 1]),
                                                 __fadd2_rn(make_float2(__uint_as_float(dp[4 * e4]), __uint_as_float(dp[4 * e4 + 1])), make_float2(-d4.x, -d4.y)));
                    const float2 a1 = __fmul2_rn(make_float2(pv[4 * e4 + 2], pv[4 * e4 + 3]),
                                                 __fadd2_rn(make_float2(__uint_as_float(dp[4 * e4 + 2]), __uint_as_float(dp[4 * e4 + 3])), make_float2(-d4.z, -d4.w)));
                    pk[2 * e4] = pack2<BF16>(a0.x, a0.y);
                    pk[2 * e4 + 1] = pack2<BF16>(a1.x, a1.y);
                }
                tr.ev(22, s);
                if (s > 0) mbar_wait(bar_dsfree, (s - 1) & 1);       // dQ^T(s-1) and dK(s-1) have read the tile
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t a = ds_row + (((((qt & 1) << 2) + u) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * u]), "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3]) : "memory");
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ds);
            }
            tr.ev(23, s);
            // ---- (b) first half of P^T(s+1)
            if (s + 1 < nsteps) {
                p_begin(s + 1);
                tr.ev(24, s);
                p_half(s + 1, 0);
            }
            tr.ev(25, s);
            // ---- (c) drain dQ^T(s): lane = head-dim index, 32 query columns -> 32 coalesced fp32 reductions
            mbar_wait(bar_dq, s & 1);
            tr.ev(26, s);
            tc_fence_after();
            {
                // (two 16-column halves: 32 more live registers would spill next to pv[] and the second half of S^T)
                uint32_t dq[16];
                const uint32_t q0 = i_d * 128 + 32 * qt;
                float* dst = p.dq_acc + (((size_t)b * p.Hq + hk * group + g_d) * p.Sq + q0) * D + r;
                const bool full = q0 + 32 <= p.Sq;
                tmem_ld16(tDP, dq);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (full || q0 + e < p.Sq) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + e * D), "f"(__uint_as_float(dq[e])) : "memory");
                tmem_ld16(tDP + 16, dq);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_dqfree);
                if (warp < 4 && s + 2 < nsteps) {                    // every warp is past the delta reads of step s: buffer s&1 is free
                    publish(s + 2);
                    if (s + 3 < nsteps) fetch_stats();
                }
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (full || q0 + 16 + e < p.Sq) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (16 + e) * D), "f"(__uint_as_float(dq[e])) : "memory");
                if (++i_d == nqb) { i_d = i_begin; ++g_d; }
            }
            tr.ev(27, s);
            // ---- (d) second half of P^T(s+1)
            if (s + 1 < nsteps) p_half(s + 1, 1);
        }
    }

    // ---- epilogue: dV, dK (x scale) -> 16-bit -> swizzled SMEM (Q stage 0 / dO stage 0, free now) -> TMA store
    __syncthreads();                                                 // every MMA is complete (the issuer waited on bar_done)
    tc_fence_after();
    if (warp < 16) {
        const uint32_t h = (warp >> 2) & 1, r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const int which = warp >> 3;                                 // warps 0-7: dV, warps 8-15: dK
        const uint32_t col = (which ? COL_DK : COL_DV) + (D / 2) * h;
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
            const uint32_t dcol = (D / 2) * h + c * 32;              // first output column of this chunk
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
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < C::CHUNKS; ++c) {
            tma_store_3d(tmdV, sQ0 + c * C::CHUNK_BYTES, c * 64, (int32_t)key0, (int32_t)bhk);
            tma_store_3d(tmdK, sdO0 + c * C::CHUNK_BYTES, c * 64, (int32_t)key0, (int32_t)bhk);
        }
        tma_store_commit();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc<512>(tmem);
    // The CTA only has to outlive the TMA engine's READ of its shared memory; the global writes complete on their own
    // (kernel completion orders them).  Waiting for the writes held every CTA ~1 us past its last useful cycle -- with 8.5
    // steps per CTA on config B that is paid 14 times per SM.
    if (threadIdx.x == 0) tma_store_wait_read<0>();
}

// fp32 dQ accumulator -> 16-bit (x scale): 8 elements per thread and iteration
template <bool BF16>
__device__ __forceinline__ void dq_convert_body(const float* acc, void* dq, uint64_t n8, float scale) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n8; t += stride) {
        const float4 a = reinterpret_cast<const float4*>(acc)[2 * t], c = reinterpret_cast<const float4*>(acc)[2 * t + 1];
        uint4 o;
        o.x = pack2<BF16>(a.x * scale, a.y * scale); o.y = pack2<BF16>(a.z * scale, a.w * scale);
        o.z = pack2<BF16>(c.x * scale, c.y * scale); o.w = pack2<BF16>(c.z * scale, c.w * scale);
        reinterpret_cast<uint4*>(dq)[t] = o;
    }
}

}  // namespace bwd100f

#define AULE_BWD100_FUSED(NAME, BF)                                                                      \
    extern "C" __global__ void __launch_bounds__(544, 1) NAME(const __grid_constant__ CUtensorMap tmQ,    \
                                                              const __grid_constant__ CUtensorMap tmK,    \
                                                              const __grid_constant__ CUtensorMap tmV,    \
                                                              const __grid_constant__ CUtensorMap tmdO,   \
                                                              const __grid_constant__ CUtensorMap tmdK,   \
                                                              const __grid_constant__ CUtensorMap tmdV,   \
                                                              const aule_kp::BwdParams p) {               \
        bwd100f::bwd_fused_body<BF>(&tmQ, &tmK, &tmV, &tmdO, &tmdK, &tmdV, p);                            \
    }
AULE_BWD100_FUSED(aule_bwd_fused_sm100_bf16_d128, true)
AULE_BWD100_FUSED(aule_bwd_fused_sm100_f16_d128, false)

extern "C" __global__ void aule_bwd_dq_convert_bf16(const float* acc, void* dq, uint64_t n8, float scale) {
    bwd100f::dq_convert_body<true>(acc, dq, n8, scale);
}
extern "C" __global__ void aule_bwd_dq_convert_f16(const float* acc, void* dq, uint64_t n8, float scale) {
    bwd100f::dq_convert_body<false>(acc, dq, n8, scale);
}
