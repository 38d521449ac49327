// Attention backward for sm_100a (B200): TMA -> SMEM -> tcgen05.mma -> TMEM.                    (kernel v3)
//
// Replaces (behaviour, not code): python/aule/triton_flash.py:242-379 (_flash_attn_bwd_kernel,
// _compute_delta_kernel) and shaders/attention_backward_f32.comp:1-234.  Same math (:321-347):
//     P = exp(S*scale - LSE)   dP = dO V^T   dS = P o (dP - Delta)
//     dV += P^T dO             dK += dS^T Q  dQ += dS K          (dK, dQ scaled by `scale` at the end)
// Differences by design: every contraction runs on the tensor cores and NOTHING is reduced through global
// atomics (the reference accumulates dQ, and GQA dK/dV, with atomics, :335-347), so the gradients are
// bit-reproducible:
//   * dK/dV kernel (bwd_dkv_body): one CTA per (KV block of 128 keys, kv head, batch); K_j, V_j resident in SMEM;
//     dV [256,384) and dK [384,512) accumulate in TMEM over every q-head of the GQA group and every query block;
//   * dQ kernel (bwd_dq_body): one CTA per (query block of 128 rows, q head, batch); Q_i, dO_i resident, K_j/V_j
//     stream through a 2-stage ring, dS stays in TMEM (TS MMA), dQ accumulates in TMEM.  S, P and dS are recomputed there (two extra GEMMs and
//     one extra exp pass) -- measured faster than reducing dQ partials through 4.4 GB of fp32 red.global.add per
//     config-C/2 launch (experiments/attn_bwd_sm100_v2_fused_atomics.cu.txt), and deterministic.
// Both are 544-thread CTAs (four compute warps per scheduler hide the TMEM-load / MUFU / barrier latencies):
//   warps 0-15       compute: thread == (query row, key quarter of 32 columns)
//   warp  16         issuer: one elected thread issues every TMA load and every tcgen05.mma
// and both split a step into two phases so that the tensor pipe always has work the compute warps are not waiting on:
//   P  phase  compute: S (TMEM) -> P = exp2(S*c - LSE*log2e) kept in registers (fp32) [+ 16-bit P -> SMEM]
//             tensor : dP(i) = dO V^T          (+ dK(i-1) | dQ(j-1), S(j+1) into the second S buffer)
//   dS phase  compute: dP (TMEM) -> dS = P o (dP - Delta) -> 16-bit -> SMEM
//             tensor : dV(i) = P^T dO, S(i+1)
// Every operand tile is the same [128 rows][64 elements] 128B-swizzled sub-tile, which serves as a K-major
// operand (Q, K, V, dO, dS) and, read through an MN-major descriptor, as its own transpose (P^T, dS^T, and Q, dO, K
// as [k][n] B operands) -- no data is ever transposed.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sm100_ptx.cuh"
#include "kernel_params.h"

namespace bwd100 {
using namespace sm100;
using aule_kp::BwdParams;
template <int D> using Cfg = aule_kp::BwdCfg<D>;

// Bring-up tracer (aule_set_trace_buffer): CTA 0 only, one region of 4096 entries per traced thread.  Compiled in only
// with -DAULE_BWD_TRACE=1 (make EXTRA_NVFLAGS=-DAULE_BWD_TRACE=1; tools/bwd_trace.py): even disabled at run time its
// checks cost ~16 instructions per thread and step in kernels whose compute warps are issue-bound.
#ifndef AULE_BWD_TRACE
#define AULE_BWD_TRACE 0
#endif
#ifndef AULE_BWD_XPASS
#define AULE_BWD_XPASS 0          // 1: scale/offset pass of its own before the exp2 loop of the P phase (A/B: tools/gpu/r2_exp11.sh)
#endif
struct Tracer {
#if AULE_BWD_TRACE
    unsigned long long* buf;
    uint32_t n;
    __device__ __forceinline__ Tracer(unsigned long long* base, uint32_t region, bool on)
        : buf((base && on && blockIdx.x == 0) ? base + region * 4096 : nullptr), n(0) {}
    __device__ __forceinline__ void ev(uint32_t code, uint32_t step) {
        if (buf && n < 4096) buf[n++] = ((unsigned long long)((code << 8) | (step & 255u)) << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
    }
#else
    __device__ __forceinline__ Tracer(unsigned long long*, uint32_t, bool) {}
    __device__ __forceinline__ void ev(uint32_t, uint32_t) {}
#endif
};

// P phase for 32 columns of one row: exp2(S*c - lse2), masked when the block touches the diagonal / the ragged ends.
// EMU of every 4 element pairs take the FMA-pipe polynomial exp2 (sm100::ex2_emu2) instead of MUFU.EX2: with four
// compute warps per scheduler the 32 MUFU per thread and step are 1024 issue cycles per scheduler, the same as the
// tensor time of the phase they run under.
template <bool MASK, int EMU = 0>
__device__ __forceinline__ void p_from_s(const uint32_t (&s)[32], float* pv, float c2, float lse2, uint32_t col0, uint32_t row,
                                         uint32_t Sk, bool row_ok, bool diag) {
    const float2 cc = make_float2(c2, c2), nl = make_float2(-lse2, -lse2);
    uint32_t alive = 0xffffffffu;                         // bit e: column col0+e of this row is visible
    if (MASK) {
        const int64_t last = min((int64_t)Sk - 1, diag ? (int64_t)row : (int64_t)0x7fffffff) - (int64_t)col0;   // last visible local column
        alive = !row_ok || last < 0 ? 0u : (last >= 31 ? 0xffffffffu : (0xffffffffu >> (31 - (int)last)));
    }
#if AULE_BWD_XPASS
    uint32_t xs[32];
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), cc, nl);
        xs[e] = __float_as_uint(x.x); xs[e + 1] = __float_as_uint(x.y);
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) asm volatile("" : "+r"(xs[e]));
#endif
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
#if AULE_BWD_XPASS
        const float2 x = make_float2(__uint_as_float(xs[e]), __uint_as_float(xs[e + 1]));
#else
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), cc, nl);
#endif
        float2 v;
        if (((e >> 1) & 3) < EMU) {
            v = ex2_emu2(x);
        } else {
            v.x = ex2(x.x);
            v.y = ex2(x.y);
        }
        if (MASK) {                                       // one alive-bit test + select per element (see attn_fwd_sm100.cu)
            v.x = (alive & (1u << e)) ? v.x : 0.f;
            v.y = (alive & (2u << e)) ? v.y : 0.f;
        }
        pv[e] = v.x;
        pv[e + 1] = v.y;
    }
}

// Same for a block on the edge of the visibility band row - win_left <= key <= row + win_right (causal: win_right = 0) or on
// a ragged end: visible local columns [max(row - win_left, 0), min(row + win_right, Sk - 1)] - col0.
__device__ __forceinline__ void p_from_s_band(const uint32_t (&s)[32], float* pv, float c2, float lse2, uint32_t col0, uint32_t row,
                                              uint32_t Sk, bool row_ok, uint32_t win_left, uint32_t win_right) {
    const int64_t first = (int64_t)row - (int64_t)win_left - (int64_t)col0;
    const int64_t last = min((int64_t)Sk - 1, (int64_t)row + (int64_t)win_right) - (int64_t)col0;
    const uint32_t lo_m = first <= 0 ? 0xffffffffu : (first > 31 ? 0u : (0xffffffffu << (int)first));
    const uint32_t hi_m = last >= 31 ? 0xffffffffu : (last < 0 ? 0u : (0xffffffffu >> (31 - (int)last)));
    const uint32_t alive = row_ok ? (lo_m & hi_m) : 0u;
    const float2 cc = make_float2(c2, c2), nl = make_float2(-lse2, -lse2);
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), cc, nl);
        const float vx = ex2(x.x), vy = ex2(x.y);
        pv[e] = (alive & (1u << e)) ? vx : 0.f;          // select, not multiply: a row without any visible key has LSE = -inf
        pv[e + 1] = (alive & (2u << e)) ? vy : 0.f;
    }
}

// 32 values of one row -> 16-bit -> swizzled [128 rows][64 elements] sub-tile (4 x 16-byte units, unit index (c*4+u) ^ (r&7))
template <bool BF16>
__device__ __forceinline__ void store_row32(uint32_t base, uint32_t r, int c, const float* v) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t a = base + r * 128 + (((c * 4 + u) ^ (r & 7)) << 4);
        const uint32_t w0 = pack2<BF16>(v[8 * u + 0], v[8 * u + 1]), w1 = pack2<BF16>(v[8 * u + 2], v[8 * u + 3]);
        const uint32_t w2 = pack2<BF16>(v[8 * u + 4], v[8 * u + 5]), w3 = pack2<BF16>(v[8 * u + 6], v[8 * u + 7]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
    }
}

// =====================================================================================================
// dK/dV kernel
template <int D, bool BF16>
__device__ __forceinline__ void bwd_dkv_body(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                             const CUtensorMap* tmdO, const CUtensorMap* tmdK, const CUtensorMap* tmdV,
                                             const BwdParams& p) {
    using C = Cfg<D>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barriers
    const uint32_t bar_kv = sb + C::OFF_BAR;        // K_j, V_j landed (once)
    const uint32_t bar_q0 = bar_kv + 8;             // Q buffer 0 / 1 landed (+8)
    const uint32_t bar_do0 = bar_kv + 24;           // dO buffer 0 landed (buffer 1: bar_do1)
    const uint32_t bar_s0 = bar_kv + 32;            // S(i) complete in buffer i&1 (commit) (+8 for odd i)
    const uint32_t bar_dp0 = bar_kv + 48;           // dP(i) complete in buffer i&1 (commit) (+8 for odd i)
    const uint32_t bar_p = bar_kv + 64;             // compute -> issuer: P(i) in SMEM (one arrival per compute warp)
    const uint32_t bar_ds = bar_kv + 72;            // compute -> issuer: dS(i) in SMEM, buffer i&1 consumed (16 arrivals)
    const uint32_t bar_dv0 = bar_kv + 80;           // dV(i) complete (commit), even i: P buffer and dO buffer 0 free (odd i: bar_dv1)
    const uint32_t bar_dk = bar_kv + 88;            // dK(i) complete (commit): dS buffer and Q buffer i&1 free
    const uint32_t bar_sfree = bar_kv + 96;         // compute -> issuer: S(i) copied to registers, its buffer may take dP(i) (16 arrivals)
    const uint32_t bar_do1 = bar_kv + 104, bar_dv1 = bar_kv + 112;
    // dV(i) commits to bar_dv[i&1] so that "dV(i-2) complete" (dO buffer i&1 reusable) can never be confused with
    // dV(i-1): each of the two barriers completes once per two steps and its next completion needs the very load
    // that is waiting on it.
    auto bar_do = [&](uint32_t i) { return (i & 1) ? bar_do1 : bar_do0; };
    auto bar_dv = [&](uint32_t i) { return (i & 1) ? bar_dv1 : bar_dv0; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_SLOT);
    const uint32_t sK = sb + C::OFF_K, sV = sb + C::OFF_V, sQ0 = sb + C::OFF_Q, sdO0 = sb + C::OFF_DO,
                   sP = sb + C::OFF_P, sdS = sb + C::OFF_DS;

    if (threadIdx.x == 0) {
        if (sb & 1023u) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }
        mbar_init(bar_kv, 1); mbar_init(bar_q0, 1); mbar_init(bar_q0 + 8, 1); mbar_init(bar_do0, 1); mbar_init(bar_do1, 1);
        mbar_init(bar_s0, 1); mbar_init(bar_s0 + 8, 1); mbar_init(bar_dp0, 1); mbar_init(bar_dp0 + 8, 1);
        mbar_init(bar_p, 16); mbar_init(bar_ds, 16); mbar_init(bar_sfree, 16);    // one arrival per compute warp
        mbar_init(bar_dv0, 1); mbar_init(bar_dv1, 1); mbar_init(bar_dk, 1);
        fence_mbar_init();
        tma_prefetch_desc(tmQ); tma_prefetch_desc(tmK); tma_prefetch_desc(tmV); tma_prefetch_desc(tmdO);
    }
    if (warp == 16) tmem_alloc<512>(sb + C::OFF_TMEM_SLOT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // S(i) and dP(i) share buffer i&1 (columns [0,128) / [128,256)): S(i) is dead once the compute warps hold it in
    // registers, so dP(i) lands in the same columns while S(i+1) is already waiting in the other buffer.
    constexpr uint32_t COL_DV = 256, COL_DK = 384;

    // ---- which KV block
    const uint32_t per = p.Hkv * p.B;
    // CTA order: runs of `units_per_run` (batch, kv-head) units; inside a run KV block 0 (heaviest under causal) of
    // every unit first.  All KV-block CTAs of a unit stream the same Q / dO tiles: keeping only a few units in
    // flight keeps those tiles L2-resident (all units at once re-read them from HBM 3.5x, ncu) and cuts the TMA
    // latency the single dO buffer exposes.
    const uint32_t nkb_grid = gridDim.x / per;
    const uint32_t upr = (p.units_per_run == 0 || p.units_per_run > per) ? per : p.units_per_run;
    const uint32_t run = blockIdx.x / (upr * nkb_grid), lw = blockIdx.x - run * upr * nkb_grid;
    const uint32_t uir = min(upr, per - run * upr);                  // units in this run (the last may be short)
    const uint32_t jb = lw / uir;                                    // KV block
    const uint32_t bhk = run * upr + (lw - jb * uir);                // b * Hkv + hk
    const uint32_t b = bhk / p.Hkv, hk = bhk - b * p.Hkv;
    const uint32_t group = p.Hq / p.Hkv;
    const uint32_t key0 = jb * 128;
    const uint32_t nqb = (p.Sq + 127) / 128;
    const uint32_t i_begin = p.causal ? jb : 0;                      // query blocks with rows >= key0 (top-left causal)
    const uint32_t steps_per_head = (i_begin < nqb) ? (nqb - i_begin) : 0;
    const uint32_t nsteps = steps_per_head * group;

    if (warp == 16) {
        // ===================================================== issuer
        if (elect_one()) {
            constexpr uint64_t HI_K = smem_desc_hi(16, 1024);                // K-major SW128
            constexpr uint64_t HI_MN = smem_desc_hi(C::CHUNK_BYTES, 1024);   // MN-major SW128 (64-wide chunks 16 KB apart)
            constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
            constexpr uint32_t HI_MN_HI = uint32_t(HI_MN >> 32), HI_MN_LO = uint32_t(HI_MN);
            auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
            constexpr uint32_t ID_KK = instr_desc_f16(BF16, 128, 128, false);                    // A K-major, B K-major, N=128
            constexpr uint32_t ID_MNMN = instr_desc_f16(BF16, 128, D, true) | (1u << 15);        // A MN-major, B MN-major, N=D
            auto coords = [&](uint32_t step, int32_t& row, int32_t& bh) {
                const uint32_t g = step / steps_per_head, i = i_begin + (step - g * steps_per_head);
                row = (int32_t)(i * 128);
                bh = (int32_t)(b * p.Hq + hk * group + g);
            };                                                               // (issuer thread only: one division per load)
            auto load_q = [&](uint32_t step) {
                int32_t row, bh; coords(step, row, bh);
                const uint32_t bar = bar_q0 + 8 * (step & 1), dst = sQ0 + (step & 1) * C::TILE_BYTES;
                mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                for (int c = 0; c < C::CHUNKS; ++c) tma_load_3d(dst + c * C::CHUNK_BYTES, tmQ, bar, c * 64, row, bh);
                if (step + 1 < nsteps) {                                     // the next Q tile is on the critical path
                    coords(step + 1, row, bh);                               // (dK -> Q load -> S -> P): warm it in L2
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c) tma_prefetch_3d(tmQ, c * 64, row, bh);
                }
            };
            auto load_do = [&](uint32_t step) {
                int32_t row, bh; coords(step, row, bh);
                const uint32_t bar = bar_do(step), dst = sdO0 + (step & 1) * C::TILE_BYTES;
                mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                for (int c = 0; c < C::CHUNKS; ++c) tma_load_3d(dst + c * C::CHUNK_BYTES, tmdO, bar, c * 64, row, bh);
            };
            auto issue_s = [&](uint32_t step) {                              // S = Q K^T
                const uint32_t sQ = sQ0 + (step & 1) * C::TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + 128 * (step & 1), mk(HI_K_HI, (HI_K_LO | (sQ >> 4)) + off), mk(HI_K_HI, (HI_K_LO | (sK >> 4)) + off), ID_KK, kk > 0);
                }
                mma_commit(bar_s0 + 8 * (step & 1));
            };
            auto issue_dk = [&](uint32_t step) {                             // dK += dS^T Q   (K = 128 query rows)
                const uint32_t sQ = sQ0 + (step & 1) * C::TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    mma_ss(tmem + COL_DK, mk(HI_MN_HI, (HI_MN_LO | (sdS >> 4)) + kk * 128), mk(HI_MN_HI, (HI_MN_LO | (sQ >> 4)) + kk * 128),
                           ID_MNMN, (step > 0 || kk > 0) ? 1u : 0u);
                mma_commit(bar_dk);
            };
            mbar_expect_tx(bar_kv, 2 * C::TILE_BYTES);
#pragma unroll
            for (int c = 0; c < C::CHUNKS; ++c) {
                tma_load_3d(sK + c * C::CHUNK_BYTES, tmK, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
                tma_load_3d(sV + c * C::CHUNK_BYTES, tmV, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
            }
            Tracer tr(p.trace, 0, true);
            // ---- loads, issued from the wait loops as soon as their buffer is released:
            //   Q(s)  -> buffer s&1 once dK(s-2) has read it;   dO(s) -> buffer s&1 once dV(s-2) has read it.
            uint32_t ql = 0, dl = 0;
            auto pump = [&]() {
                if (ql < nsteps && (ql < 2 || mbar_try_wait<0>(bar_dk, (ql - 2) & 1))) { load_q(ql); ++ql; }
                if (dl < nsteps && (dl < 2 || mbar_try_wait<0>(bar_dv(dl), ((dl - 2) >> 1) & 1))) { load_do(dl); ++dl; }
            };
            auto wait = [&](uint32_t bar, uint32_t parity) {
                while (!mbar_try_wait<0>(bar, parity)) pump();
            };
            pump(); pump();
            wait(bar_kv, 0);
            if (nsteps > 0) {
                wait(bar_q0, 0);
                tc_fence_after();
                issue_s(0);
            }
            // Steady state, per step i (tensor pipe order == readiness order):
            //   dK(i-1)  [dS(i-1) ready]          -> releases Q buffer (i+1)&1 -> TMA Q(i+1)
            //   dP(i)    [S(i) in registers]      -> buffer i&1
            //   S(i+1)   [Q(i+1) landed]          -> buffer (i+1)&1 (its dP(i-1) was consumed with dS(i-1))
            //   dV(i)    [P(i) ready]             -> releases the dO buffer -> TMA dO(i+1)
            for (uint32_t step = 0; step < nsteps; ++step) {
                const uint32_t sdO = sdO0 + (step & 1) * C::TILE_BYTES;
                if (step > 0) {
                    wait(bar_ds, (step - 1) & 1);                            // dS(step-1) written, buffer (step-1)&1 consumed
                    tc_fence_after();
                    tr.ev(13, step);
                    issue_dk(step - 1);
                }
                tr.ev(10, step);
                wait(bar_sfree, step & 1);                                   // S(step) is in registers
                tr.ev(11, step);
                wait(bar_do(step), (step >> 1) & 1);
                tr.ev(12, step);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {                        // dP = dO V^T
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + 128 * (step & 1), mk(HI_K_HI, (HI_K_LO | (sdO >> 4)) + off), mk(HI_K_HI, (HI_K_LO | (sV >> 4)) + off), ID_KK, kk > 0);
                }
                mma_commit(bar_dp0 + 8 * (step & 1));
                auto issue_dv = [&]() {
                    tr.ev(17, step);
                    wait(bar_p, step & 1);                                   // P(step) written
                    tr.ev(18, step);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)                           // dV += P^T dO   (K = 128 query rows)
                        mma_ss(tmem + COL_DV, mk(HI_MN_HI, (HI_MN_LO | (sP >> 4)) + kk * 128), mk(HI_MN_HI, (HI_MN_LO | (sdO >> 4)) + kk * 128),
                               ID_MNMN, (step > 0 || kk > 0) ? 1u : 0u);
                    mma_commit(bar_dv(step));
                };
                // With the shared P/dS buffer dV(step) gates dS(step) (and through it the next step): it goes first and
                // S(step+1), which may still be waiting for its Q tile, after it.
                if (C::SHARE_PDS) issue_dv();
                if (step + 1 < nsteps) {
                    wait(bar_q0 + 8 * ((step + 1) & 1), ((step + 1) >> 1) & 1);   // (Q(step+1) is requested by pump once dK(step-1) is done)
                    tc_fence_after();
                    tr.ev(14, step);
                    issue_s(step + 1);
                }
                if (!C::SHARE_PDS) issue_dv();
            }
            if (nsteps > 0) {
                wait(bar_ds, (nsteps - 1) & 1);
                tc_fence_after();
                issue_dk(nsteps - 1);
                wait(bar_dk, (nsteps - 1) & 1);                              // every MMA complete
            }
        }
    } else {
        // ===================================================== compute warps
        const uint32_t qt = warp >> 2;                               // key quarter: columns [32qt, 32qt+32)
        const uint32_t r = (warp & 3) * 32 + lane;                   // row of the tile == TMEM lane
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        // row statistics of the NEXT step are fetched one step ahead and only touched one step later;
        // (g, i) = (q-head of the group, query block) advance without a division
        float lse_n = 0.f, delta_n = 0.f;
        uint32_t g_n = 0, i_n = i_begin;                             // coordinates of the step being prefetched
        auto fetch_stats = [&]() {
            const uint32_t row = i_n * 128 + r;
            const size_t off = ((size_t)b * p.Hq + hk * group + g_n) * p.Sq;
            const bool ok = row < p.Sq;
            lse_n = ok ? p.lse[off + row] : 0.f;
            delta_n = ok ? p.delta[off + row] : 0.f;
            if (++i_n == nqb) { i_n = i_begin; ++g_n; }
        };
        if (nsteps > 0) fetch_stats();
        Tracer tr(p.trace, 1 + (qt & 1), (warp == 0 || warp == 4) && lane == 0);
        uint32_t i = i_begin;                                        // query block of the current step
        for (uint32_t step = 0; step < nsteps; ++step) {
            const uint32_t row = i * 128 + r;                        // global query row of this thread
            const bool row_ok = row < p.Sq;
            const float lse2 = lse_n * 1.4426950408889634f, delta = delta_n;
            if (step + 1 < nsteps) fetch_stats();
            const bool diag = p.causal && (i * 128 < key0 + 128);    // block touches the diagonal
            const bool masked = diag || key0 + 128 > p.Sk || i * 128 + 128 > p.Sq;
            if (++i == nqb) i = i_begin;

            // ---- P phase: P = exp2(S*scale_log2 - LSE*log2e) (fp32, kept in registers) -> 16-bit -> swizzled SMEM
            float pv[32];
            const uint32_t tbuf = tmem + lane_addr + 128 * (step & 1) + 32 * qt;
            tr.ev(20, step);
            mbar_wait(bar_s0 + 8 * (step & 1), (step >> 1) & 1);
            tr.ev(21, step);
            tc_fence_after();
            {
                uint32_t s[32];
                tmem_ld32(tbuf, s);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sfree);               // the buffer may take dP(step)
                tr.ev(22, step);
                if (masked) p_from_s<true>(s, pv, p.scale_log2, lse2, key0 + 32 * qt, row, p.Sk, row_ok, diag);
                else p_from_s<false, 1>(s, pv, p.scale_log2, lse2, 0, 0, 0, true, false);   // 1 pair in 4 on the FMA pipe: -3 % (A/B, s20)
            }
            tr.ev(28, step);
            if (step > 0) {
                if (C::SHARE_PDS) mbar_wait(bar_dk, (step - 1) & 1);  // dK(step-1) has read dS(step-1) out of the shared buffer
                else mbar_wait(bar_dv(step - 1), ((step - 1) >> 1) & 1);   // dV(step-1) has read the P buffer
            }
            store_row32<BF16>(sP + (qt >> 1) * C::CHUNK_BYTES, r, qt & 1, pv);
            fence_proxy_async_smem();                                // generic-proxy writes -> visible to the MMA (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p);
            tr.ev(23, step);

            // ---- dS phase: dS = P o (dP - Delta) -> 16-bit -> swizzled SMEM
            mbar_wait(bar_dp0 + 8 * (step & 1), (step >> 1) & 1);
            tr.ev(24, step);
            tc_fence_after();
            {
                uint32_t dp[32];
                tmem_ld32(tbuf, dp);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 32; ++e) pv[e] *= (__uint_as_float(dp[e]) - delta);
            }
            tr.ev(25, step);
            if (C::SHARE_PDS) mbar_wait(bar_dv(step), (step >> 1) & 1);   // dV(step) has read P(step) out of the shared buffer
            else if (step > 0) mbar_wait(bar_dk, (step - 1) & 1);    // dK(step-1) has read the dS buffer
            tr.ev(26, step);
            store_row32<BF16>(sdS + (qt >> 1) * C::CHUNK_BYTES, r, qt & 1, pv);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ds);
            tr.ev(27, step);
        }
    }

    // ---- epilogue: dV, dK (x scale) -> 16-bit -> swizzled SMEM (P buffer; dS buffer, or Q buffer 0 when P and dS
    //      share theirs) -> TMA store
    const uint32_t sKst = C::SHARE_PDS ? sQ0 : sdS;
    __syncthreads();                                                 // every MMA is complete (the issuer waited on the last commit)
    tc_fence_after();
    if (warp < 16) {
        const uint32_t h = (warp >> 2) & 1, r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        {
            const int which = warp >> 3;                             // warps 0-7: dV, warps 8-15: dK
            const uint32_t col = (which ? COL_DK : COL_DV) + (D / 2) * h;
            const uint32_t sbuf = which ? sKst : sP;
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
            tma_store_3d(tmdK, sKst + c * C::CHUNK_BYTES, c * 64, (int32_t)key0, (int32_t)bhk);
        }
        tma_store_commit();
        tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc<512>(tmem);
}

// =====================================================================================================
// dK/dV kernel, transposed form (v4).  Same CTA decomposition and the same dV / dK accumulators as bwd_dkv_body, but
// the score tiles are computed TRANSPOSED -- S^T = K_j Q_i^T and dP^T = V_j dO_i^T, rows (TMEM lanes) = keys, columns =
// queries -- so that P^T and dS^T, written back in place over their own fp32 columns, are directly the TMEM A operands
// of  dV += P^T dO  and  dK += dS^T Q  (TS MMAs).  Nothing is staged through shared memory: the two 32 KB P / dS
// buffers of bwd_dkv_body, their ~1200 cycles of st.shared + proxy fences per step and the single-buffered dO are
// gone; the freed space holds a 3-stage Q ring and a 2-stage dO ring, which takes the TMA latency off the chain.
//   TMEM: [0,128) S^T(i) -> P^T(i) | [128,256) dP^T(i) -> dS^T(i) | [256,384) dV | [384,512) dK
//   tensor order per step i:   dV(i)  S^T(i+1)  dK(i)  dP^T(i+1)     (in-order execution closes the MMA->MMA hazards:
//   S^T(i+1) overwrites P^T(i) after dV(i) has read it, dP^T(i+1) overwrites dS^T(i) after dK(i) has)
// Row statistics are per COLUMN here (LSE_i, Delta_i of the 128 queries): 128 threads publish them one step ahead in
// shared memory, every thread reads its 32 with broadcast LDS.128.
template <int D, bool BF16>
__device__ __forceinline__ void bwd_dkv_t_body(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                               const CUtensorMap* tmdO, const CUtensorMap* tmdK, const CUtensorMap* tmdV,
                                               const BwdParams& p) {
    using C = aule_kp::BwdTCfg<D>;
    constexpr int NQ = C::NQ, NDO = C::NDO;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_kv = sb + C::OFF_BAR;            // K_j, V_j landed (once)
    const uint32_t bar_qfull0 = bar_kv + 8;             // Q stage s landed (+8s)
    const uint32_t bar_qfree0 = bar_qfull0 + 8 * NQ;    // dK(i) complete (commit): Q stage i%NQ free (+8s)
    const uint32_t bar_dofull0 = bar_qfree0 + 8 * NQ;   // dO stage s landed (+8s)
    const uint32_t bar_dofree0 = bar_dofull0 + 8 * NDO; // dV(i) complete (commit): dO stage i%NDO free (+8s)
    const uint32_t bar_s = bar_dofree0 + 8 * NDO;       // S^T(i) complete (commit)
    const uint32_t bar_dp = bar_s + 8;                  // dP^T(i) complete (commit)
    const uint32_t bar_p = bar_dp + 8;                  // compute -> issuer: P^T(i) in TMEM (16 arrivals)
    const uint32_t bar_ds = bar_p + 8;                  // compute -> issuer: dS^T(i) in TMEM (16 arrivals)
    const uint32_t bar_done = bar_ds + 8;               // every MMA complete (commit)
    const uint32_t bar_pb = bar_done + 24;              // compute -> issuer: second half of P^T(i) (bar_p announces the first half)
    const uint32_t bar_s1 = bar_done + 32;              // S^T(i) complete, odd steps (D = 64: second S^T buffer, one barrier per buffer)
    const uint32_t bar_p1 = bar_done + 40;              // bar_p / bar_pb of the second buffer: with S^T(i+1) available early a fast
    const uint32_t bar_pb1 = bar_done + 48;             //   warp reaches step i+1 before a slow one has published P^T(i); on shared
                                                        //   barriers those arrivals would complete step i's phase without the slow
                                                        //   warp and dV(i) would read a P^T that is not there (tools/bwd_protocol_sim.py)
    const uint32_t bar_stat0 = bar_done + 8;            // publishers -> everyone: statistics of step s are in buffer s&1 (4 arrivals);
                                                        // one barrier per buffer, so a waiter can never be lapped (the next
                                                        // completion of ITS barrier needs its own dS^T arrival two steps on)
    static_assert(8 * (1 + 2 * NQ + 2 * NDO + 11) <= C::BAR_BYTES, "barrier area too small");
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_SLOT);
    float* stat = reinterpret_cast<float*>(smem + C::OFF_STAT);      // [3][lse2 x128 | delta x128]
    const uint32_t sK = sb + C::OFF_K, sV = sb + C::OFF_V, sQ0 = sb + C::OFF_Q, sdO0 = sb + C::OFF_DO;

    if (threadIdx.x == 0) {
        if (sb & 1023u) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }
        mbar_init(bar_kv, 1);
        for (int i = 0; i < NQ; ++i) { mbar_init(bar_qfull0 + 8 * i, 1); mbar_init(bar_qfree0 + 8 * i, 1); }
        for (int i = 0; i < NDO; ++i) { mbar_init(bar_dofull0 + 8 * i, 1); mbar_init(bar_dofree0 + 8 * i, 1); }
        mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_p, 16); mbar_init(bar_ds, 16); mbar_init(bar_done, 1);
        mbar_init(bar_stat0, 4); mbar_init(bar_stat0 + 8, 4); mbar_init(bar_pb, 16); mbar_init(bar_s1, 1); mbar_init(bar_p1, 16); mbar_init(bar_pb1, 16);
        fence_mbar_init();
        tma_prefetch_desc(tmQ); tma_prefetch_desc(tmK); tma_prefetch_desc(tmV); tma_prefetch_desc(tmdO);
    }
    if (warp == 16) tmem_alloc<512>(sb + C::OFF_TMEM_SLOT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM: S^T [0,128) | dP^T [128,256) | dV [256,256+D) | dK.  D = 128 fills the 512 columns; D = 64 leaves 128, which hold a
    // SECOND S^T buffer [384,512): S^T(i+1) then no longer has to wait behind dV(i) for the columns P^T(i) occupies, it is
    // issued at the top of step i and is complete long before the compute warps finish dS^T(i).  (Trace of config B before:
    // ~600 of the ~2950 cycles per step were the compute warps waiting for S^T(i+1), whose issue waited for every warp's
    // P^T(i); BwdParams::order bit 3 keeps the single buffer for A/B.)
    constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = (D == 64) ? 320 : 384, COL_S1 = 384;
    const bool two_s = (D == 64) && !(p.order & 8);

    // ---- which KV block
    const uint32_t per = p.Hkv * p.B;
    const uint32_t jb = blockIdx.x / per;                            // KV block (0 = heaviest under causal)
    const uint32_t bhk = blockIdx.x - jb * per;                      // b * Hkv + hk
    const uint32_t b = bhk / p.Hkv, hk = bhk - b * p.Hkv;
    const uint32_t group = p.Hq / p.Hkv;
    const uint32_t key0 = jb * 128;
    const uint32_t nqb = (p.Sq + 127) / 128;
    // query blocks whose rows can see a key of this block: key - win_right <= q <= key + win_left (causal: win_right = 0)
    const uint32_t i_begin = key0 > p.win_right ? (key0 - p.win_right) / 128 : 0;
    const uint32_t i_end = (uint32_t)min((uint64_t)nqb, ((uint64_t)key0 + 127 + p.win_left) / 128 + 1);
    const uint32_t steps_per_head = (i_begin < i_end) ? (i_end - i_begin) : 0;
    const uint32_t nsteps = steps_per_head * group;

    if (warp == 16) {
        // ===================================================== issuer
        if (elect_one() && nsteps > 0) {
            constexpr uint64_t HI_K = smem_desc_hi(16, 1024);                // K-major SW128
            constexpr uint64_t HI_MN = smem_desc_hi(C::CHUNK_BYTES, 1024);   // MN-major SW128 (64-wide chunks 16 KB apart)
            constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
            constexpr uint32_t HI_MN_HI = uint32_t(HI_MN >> 32), HI_MN_LO = uint32_t(HI_MN);
            auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
            constexpr uint32_t ID_KK = instr_desc_f16(BF16, 128, 128, false);    // A K-major, B K-major, N = 128 queries
            constexpr uint32_t ID_KMN = instr_desc_f16(BF16, 128, D, true);      // A in TMEM, B MN-major, N = D
            // (q-head of the group, query block) of the next Q / dO load, advanced without divisions
            uint32_t ql = 0, ql_st = 0, ql_use = 0, ql_g = 0, ql_i = i_begin;
            uint32_t dl = 0, dl_st = 0, dl_use = 0, dl_g = 0, dl_i = i_begin;
            // side polls never suspend (mbarrier.test_wait): a try_wait on a stage that is not free yet sleeps for the hardware's
            // time limit (~400 cycles measured) and delays the check of the barrier the issuer is actually waiting for
            // (p.order bit 2 = the old try_wait polls, for A/B)
            const bool legacy_poll = (p.order & 4) != 0;
            auto poll = [&](uint32_t bar, uint32_t parity) { return legacy_poll ? mbar_try_wait<0>(bar, parity) : mbar_test(bar, parity); };
            auto pump = [&]() {
                if (ql < nsteps && (ql_use == 0 || poll(bar_qfree0 + 8 * ql_st, (ql_use - 1) & 1))) {
                    const uint32_t bar = bar_qfull0 + 8 * ql_st, dst = sQ0 + ql_st * C::TILE_BYTES;
                    mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        tma_load_3d(dst + c * C::CHUNK_BYTES, tmQ, bar, c * 64, (int32_t)(ql_i * 128), (int32_t)(b * p.Hq + hk * group + ql_g));
                    ++ql;
                    if (++ql_i == i_end) { ql_i = i_begin; ++ql_g; }
                    if (++ql_st == NQ) { ql_st = 0; ++ql_use; }
                }
                if (dl < nsteps && (dl_use == 0 || poll(bar_dofree0 + 8 * dl_st, (dl_use - 1) & 1))) {
                    const uint32_t bar = bar_dofull0 + 8 * dl_st, dst = sdO0 + dl_st * C::TILE_BYTES;
                    mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        tma_load_3d(dst + c * C::CHUNK_BYTES, tmdO, bar, c * 64, (int32_t)(dl_i * 128), (int32_t)(b * p.Hq + hk * group + dl_g));
                    ++dl;
                    if (++dl_i == i_end) { dl_i = i_begin; ++dl_g; }
                    if (++dl_st == NDO) { dl_st = 0; ++dl_use; }
                }
            };
            auto wait = [&](uint32_t bar, uint32_t parity) {                 // blocking wait that keeps the loads flowing
                while (!mbar_try_wait<0>(bar, parity)) pump();
            };
            auto issue_s = [&](uint32_t qst, uint32_t buf) {                 // S^T = K_j Q^T  (A = K_j, B = Q as [n = query][k = d])
                const uint32_t sQ = sQ0 + qst * C::TILE_BYTES;
                const uint32_t col = buf ? COL_S1 : COL_S;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + col, mk(HI_K_HI, (HI_K_LO | (sK >> 4)) + off), mk(HI_K_HI, (HI_K_LO | (sQ >> 4)) + off), ID_KK, kk > 0);
                }
                mma_commit(buf ? bar_s1 : bar_s);
            };
            auto issue_dp = [&](uint32_t dst_) {                             // dP^T = V_j dO^T
                const uint32_t sdO = sdO0 + dst_ * C::TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ss(tmem + COL_DP, mk(HI_K_HI, (HI_K_LO | (sV >> 4)) + off), mk(HI_K_HI, (HI_K_LO | (sdO >> 4)) + off), ID_KK, kk > 0);
                }
                mma_commit(bar_dp);
            };
            mbar_expect_tx(bar_kv, 2 * C::TILE_BYTES);
#pragma unroll
            for (int c = 0; c < C::CHUNKS; ++c) {
                tma_load_3d(sK + c * C::CHUNK_BYTES, tmK, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
                tma_load_3d(sV + c * C::CHUNK_BYTES, tmV, bar_kv, c * 64, (int32_t)key0, (int32_t)bhk);
            }
            for (int t = 0; t < NQ; ++t) pump();                             // fill both rings
            wait(bar_kv, 0);
            wait(bar_qfull0, 0);
            tc_fence_after();
            issue_s(0, 0);
            wait(bar_dofull0, 0);
            tc_fence_after();
            issue_dp(0);
            Tracer tr(p.trace, 0, true);
            uint32_t qs = 0, qp = 0, ds_ = 0, dp_ = 0;                        // Q / dO stage of step i and the parity of its fill
            for (uint32_t step = 0; step < nsteps; ++step) {
                const uint32_t qs_next = (qs == NQ - 1) ? 0 : qs + 1, qp_next = (qs == NQ - 1) ? (qp ^ 1) : qp;
                const uint32_t ds_next = (ds_ == NDO - 1) ? 0 : ds_ + 1, dp_next = (ds_ == NDO - 1) ? (dp_ ^ 1) : dp_;
                // dV += P^T dO (K = 128 queries, A = P^T in TMEM), in two halves: every thread publishes the first 16 of its
                // 32 query columns (the even k-steps) before it computes the other 16, so half of dV runs under that math
                // instead of after it -- dV gates S^T(step+1), the head of the next step's chain.
                const uint32_t colS = (two_s && (step & 1)) ? COL_S1 : COL_S;      // where S^T(step) / P^T(step) live
                if (two_s && step + 1 < nsteps) {
                    // second buffer: its previous tenant P^T(step-1) was read by dV(step-1), issued one iteration ago and
                    // ahead of this in the in-order tensor pipe; every warp stored that P^T before bar_pb(step-1) completed
                    wait(bar_qfull0 + 8 * qs_next, qp_next);
                    tr.ev(14, step);
                    tc_fence_after();
                    issue_s(qs_next, (step + 1) & 1);
                }
                {
                    const uint32_t sdO = sdO0 + ds_ * C::TILE_BYTES;
                    tr.ev(17, step);
                    const bool odd_b = two_s && (step & 1);
                    const uint32_t p_par = two_s ? ((step >> 1) & 1) : (step & 1);
                    wait(odd_b ? bar_p1 : bar_p, p_par);                     // first halves of P^T(step) in TMEM
                    tr.ev(18, step);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 8; kk += 2)
                        mma_ts(tmem + COL_DV, tmem + colS + 32 * (kk >> 1) + 8 * (kk & 1),
                               mk(HI_MN_HI, (HI_MN_LO | (sdO >> 4)) + kk * 128), ID_KMN, (step > 0 || kk > 0) ? 1u : 0u);
                    wait(odd_b ? bar_pb1 : bar_pb, p_par);                   // second halves
                    tc_fence_after();
#pragma unroll
                    for (int kk = 1; kk < 8; kk += 2)
                        mma_ts(tmem + COL_DV, tmem + colS + 32 * (kk >> 1) + 8 * (kk & 1),
                               mk(HI_MN_HI, (HI_MN_LO | (sdO >> 4)) + kk * 128), ID_KMN, 1u);
                    mma_commit(bar_dofree0 + 8 * ds_);
                }
                if (!two_s && step + 1 < nsteps) {
                    wait(bar_qfull0 + 8 * qs_next, qp_next);
                    tr.ev(14, step);
                    tc_fence_after();
                    issue_s(qs_next, 0);                                     // overwrites P^T(step): after dV(step) in the pipe
                }
                wait(bar_ds, step & 1);                                      // dS^T(step) in TMEM
                tr.ev(13, step);
                tc_fence_after();
                {
                    const uint32_t sQ = sQ0 + qs * C::TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)                           // dK += dS^T Q   (A = dS^T in TMEM)
                        mma_ts(tmem + COL_DK, tmem + COL_DP + 32 * (kk >> 1) + 8 * (kk & 1),
                               mk(HI_MN_HI, (HI_MN_LO | (sQ >> 4)) + kk * 128), ID_KMN, (step > 0 || kk > 0) ? 1u : 0u);
                    mma_commit(bar_qfree0 + 8 * qs);
                }
                if (step + 1 < nsteps) {
                    wait(bar_dofull0 + 8 * ds_next, dp_next);
                    tr.ev(12, step);
                    tc_fence_after();
                    issue_dp(ds_next);                                       // overwrites dS^T(step): after dK(step) in the pipe
                }
                qs = qs_next; qp = qp_next; ds_ = ds_next; dp_ = dp_next;
            }
            mma_commit(bar_done);
            wait(bar_done, 0);
        }
    } else {
        // ===================================================== compute warps: thread == (key row, query quarter)
        const uint32_t qt = warp >> 2;                               // query quarter: columns [32qt, 32qt+32)
        const uint32_t r = (warp & 3) * 32 + lane;                   // key row of the tile == TMEM lane
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const uint32_t key = key0 + r;
        const bool key_ok = key < p.Sk;
        const uint32_t tS = tmem + lane_addr + COL_S + 32 * qt, tDP = tmem + lane_addr + COL_DP + 32 * qt;
        // Statistics publishers (warps 0-3: thread t publishes query t): the values of step s+1 are written to buffer
        // (s+1)&1 during the dS phase of step s -- once dP^T(s) is complete every warp has finished step s-1, the last
        // reader of that buffer -- from registers fetched a step earlier (nothing here waits on a global load), and
        // announced on bar_stat.
        const uint32_t t128 = threadIdx.x;                           // < 128 for the publishers
        float lse_n = 0.f, delta_n = 0.f;
        uint32_t g_n = 0, i_n = i_begin;
        auto fetch_stats = [&]() {
            const uint32_t row = i_n * 128 + t128;
            const size_t off = ((size_t)b * p.Hq + hk * group + g_n) * p.Sq;
            const bool ok = row < p.Sq;
            lse_n = ok ? p.lse[off + row] : 0.f;
            delta_n = ok ? p.delta[off + row] : 0.f;
            if (++i_n == i_end) { i_n = i_begin; ++g_n; }
        };
        auto publish = [&](uint32_t s_) {                            // statistics of step s_ -> buffer s_&1, one arrival per warp
            float* sn = stat + (s_ & 1) * 256;
            sn[t128] = lse_n * 1.4426950408889634f; sn[128 + t128] = delta_n;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_stat0 + 8 * (s_ & 1));
        };
        if (warp < 4 && nsteps > 0) {
            fetch_stats();
            publish(0);
            if (nsteps > 1) fetch_stats();                           // step 1, published during step 0
        }
        Tracer tr(p.trace, 1 + (qt & 1), (warp == 0 || warp == 4) && lane == 0);
        uint32_t i = i_begin;
        for (uint32_t step = 0; step < nsteps; ++step) {
            mbar_wait(bar_stat0 + 8 * (step & 1), (step >> 1) & 1);  // statistics of this step are visible
            const float* sc = stat + (step & 1) * 256 + 32 * qt;     // this quarter's 32 lse2, then (+128) its 32 deltas
            const uint32_t q0 = i * 128 + 32 * qt;                   // first query of this thread's columns
            // the block is inside the band when its last key sees its first query (key0+127 - win_right <= i*128) and its
            // first key sees its last query (i*128+127 <= key0 + win_left)
            const bool edge = (uint64_t)key0 + 127 > (uint64_t)i * 128 + p.win_right || (uint64_t)i * 128 + 127 > (uint64_t)key0 + p.win_left;
            const bool masked = edge || !key_ok || key0 + 128 > p.Sk || i * 128 + 128 > p.Sq;
            if (++i == i_end) i = i_begin;

            uint32_t alive = 0xffffffffu;
            if (masked) {                                            // visible queries of this key: [key - win_right, min(key + win_left, Sq - 1)]
                const int64_t first = (int64_t)key - (int64_t)p.win_right - (int64_t)q0;
                const int64_t last = min((int64_t)p.Sq - 1, (int64_t)key + (int64_t)p.win_left) - (int64_t)q0;
                const uint32_t lo_m = first <= 0 ? 0xffffffffu : (first > 31 ? 0u : (0xffffffffu << (int)first));
                const uint32_t hi_m = last >= 31 ? 0xffffffffu : (last < 0 ? 0u : (0xffffffffu >> (31 - (int)last)));
                alive = key_ok ? (lo_m & hi_m) : 0u;
            }
            // ---- P phase: P^T = exp2(S^T*scale_log2 - lse2[query]) -> 16-bit, in place over the first 16 columns
            float pv[32];
            tr.ev(20, step);
            const bool odd_buf = two_s && (step & 1);                // S^T(step) / P^T(step) in the second buffer
            const uint32_t tSs = tS + (odd_buf ? COL_S1 : 0u);
            mbar_wait(odd_buf ? bar_s1 : bar_s, two_s ? ((step >> 1) & 1) : (step & 1));
            tr.ev(21, step);
            tc_fence_after();
            {
                uint32_t s[32];
                tmem_ld32(tSs, s);
                tmem_wait_ld();
                tr.ev(22, step);
                const float2 cc = make_float2(p.scale_log2, p.scale_log2);
#if AULE_BWD_XPASS
                // x = S^T*scale_log2 - lse2 in a pass of its own (in place) before the exp2 loop: an FFMA2 in front of every
                // MUFU pair costs far more than its issue slots (profiles/r2_fwd_softmax_loop.md)
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 l4 = *reinterpret_cast<const float4*>(sc + 4 * e4);       // broadcast
                    const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(s[4 * e4]), __uint_as_float(s[4 * e4 + 1])), cc, make_float2(-l4.x, -l4.y));
                    const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(s[4 * e4 + 2]), __uint_as_float(s[4 * e4 + 3])), cc, make_float2(-l4.z, -l4.w));
                    s[4 * e4] = __float_as_uint(x0.x); s[4 * e4 + 1] = __float_as_uint(x0.y);
                    s[4 * e4 + 2] = __float_as_uint(x1.x); s[4 * e4 + 3] = __float_as_uint(x1.y);
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) asm volatile("" : "+r"(s[e]));
#endif
#pragma unroll
                for (int half = 0; half < 2; ++half) {               // 16 query columns each: publish, then the next 16
#pragma unroll
                    for (int e4 = 4 * half; e4 < 4 * half + 4; ++e4) {
#if AULE_BWD_XPASS
                        const float2 x0 = make_float2(__uint_as_float(s[4 * e4]), __uint_as_float(s[4 * e4 + 1]));
                        const float2 x1 = make_float2(__uint_as_float(s[4 * e4 + 2]), __uint_as_float(s[4 * e4 + 3]));
#else
                        const float4 l4 = *reinterpret_cast<const float4*>(sc + 4 * e4);       // broadcast
                        const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(s[4 * e4]), __uint_as_float(s[4 * e4 + 1])), cc, make_float2(-l4.x, -l4.y));
                        const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(s[4 * e4 + 2]), __uint_as_float(s[4 * e4 + 3])), cc, make_float2(-l4.z, -l4.w));
#endif
                        float2 v0, v1;
                        if ((e4 & 1) == 0) { v0 = ex2_emu2(x0); } else { v0.x = ex2(x0.x); v0.y = ex2(x0.y); }   // 1 pair in 4 on the FMA pipe
                        v1.x = ex2(x1.x); v1.y = ex2(x1.y);
                        pv[4 * e4] = v0.x; pv[4 * e4 + 1] = v0.y; pv[4 * e4 + 2] = v1.x; pv[4 * e4 + 3] = v1.y;
                    }
                    if (masked) {                                    // alive bits of this thread's 32 query columns: q >= key (diagonal), q < Sq
#pragma unroll
                        for (int e = 16 * half; e < 16 * half + 16; ++e) pv[e] = (alive & (1u << e)) ? pv[e] : 0.f;
                    }
                    uint32_t pk[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) pk[e] = pack2<BF16>(pv[16 * half + 2 * e], pv[16 * half + 2 * e + 1]);
                    tmem_st8(tSs + 8 * half, pk);
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(half ? (odd_buf ? bar_pb1 : bar_pb) : (odd_buf ? bar_p1 : bar_p));
                    if (half == 0) tr.ev(28, step);
                }
            }
            tr.ev(23, step);
            // ---- dS phase: dS^T = P^T o (dP^T - delta[query]) -> 16-bit, in place
            mbar_wait(bar_dp, step & 1);
            tr.ev(24, step);
            tc_fence_after();
            if (warp < 4 && step + 1 < nsteps) {
                publish(step + 1);
                if (step + 2 < nsteps) fetch_stats();
            }
            {
                uint32_t dp[32], pk[16];
                tmem_ld32(tDP, dp);
                tmem_wait_ld();
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 d4 = *reinterpret_cast<const float4*>(sc + 128 + 4 * e4);  // broadcast
                    const float2 a0 = __fmul2_rn(make_float2(pv[4 * e4], pv[4 * e4 + 1]),
                                                 __fadd2_rn(make_float2(__uint_as_float(dp[4 * e4]), __uint_as_float(dp[4 * e4 + 1])), make_float2(-d4.x, -d4.y)));
                    const float2 a1 = __fmul2_rn(make_float2(pv[4 * e4 + 2], pv[4 * e4 + 3]),
                                                 __fadd2_rn(make_float2(__uint_as_float(dp[4 * e4 + 2]), __uint_as_float(dp[4 * e4 + 3])), make_float2(-d4.z, -d4.w)));
                    pk[2 * e4] = pack2<BF16>(a0.x, a0.y);
                    pk[2 * e4 + 1] = pack2<BF16>(a1.x, a1.y);
                }
                tr.ev(25, step);
                tmem_st16(tDP, pk);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ds);
            }
            tr.ev(27, step);
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

// =====================================================================================================
// dQ kernel: one CTA per (query block of 128 rows, q head, batch), heaviest (last block) first.
// Every MMA is a TS MMA (A operand in TMEM), so a 128x128x16 step reads 4 KB of shared memory instead of the 8 KB
// (= the whole 128 B/clk) an SS step needs -- measured: the SS forms of S and dP took ~900 cycles per 128^3 GEMM
// instead of 512 (tools/bwd_trace.py):
//   Q_i, dO_i      : global -> registers -> TMEM once per CTA (A operands of S = Q K^T and dP = dO V^T)
//   dS(j)          : written in place over the dP(j) columns the same thread has just read (A operand of dQ += dS K)
// TMEM: Q [0,64) | dO [64,128) | S/dP/dS buffer 0 [128,256) / 1 [256,384) | dQ [384,384+D).
//   buffer j&1 holds S(j), then dP(j) (issued once the compute warps hold S(j) in registers), then dS(j) in columns
//   32q+[0,16) (q = key quarter); dQ(j) precedes S(j+2) in the tensor pipe, so the buffer is recycled without a wait.
// SMEM: K ring 4 stages (K_j lives from S(j) to dQ(j)) | V ring 3 stages (V_j is dead after dP(j)); the loads are
// issued from the issuer's wait loops as soon as a stage is released.
// Tensor-pipe order per step j:  dQ(j-1)  dP(j)  S(j+1).
template <int D, bool BF16>
__device__ __forceinline__ void bwd_dq_body(const CUtensorMap* tmK, const CUtensorMap* tmV, const BwdParams& p) {
    using C = aule_kp::BwdDqCfg<D>;
    constexpr int NK = C::NK, NV = C::NV;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_qdo = sb + C::OFF_BAR;       // compute -> issuer: Q_i, dO_i are in TMEM (one arrival per compute warp)
    const uint32_t bar_kfull0 = bar_qdo + 8;        // K stage s landed (+8s)
    const uint32_t bar_kfree0 = bar_kfull0 + 8 * NK;   // dQ(j) complete (commit): K stage j%NK free (+8s)
    const uint32_t bar_vfull0 = bar_kfree0 + 8 * NK;   // V stage s landed (+8s)
    const uint32_t bar_vfree0 = bar_vfull0 + 8 * NV;   // dP(j) complete (commit): V stage j%NV free (+8s)
    const uint32_t bar_s0 = bar_vfree0 + 8 * NV;    // S(j) complete in buffer j&1 (commit) (+8)
    const uint32_t bar_dp0 = bar_s0 + 16;           // dP(j) complete in buffer j&1 (commit) (+8)
    const uint32_t bar_ds = bar_dp0 + 16;           // compute -> issuer: dS(j) in TMEM (16 arrivals)
    const uint32_t bar_sfree = bar_ds + 8;          // compute -> issuer: S(j) in registers, its buffer may take dP(j) (16)
    const uint32_t bar_done = bar_sfree + 8;        // every MMA complete (commit)
    static_assert(8 * (1 + 2 * NK + 2 * NV + 4 + 3) <= C::BAR_BYTES, "barrier area too small");
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_SLOT);
    const uint32_t sK0 = sb + C::OFF_K, sV0 = sb + C::OFF_V;

    if (threadIdx.x == 0) {
        if (sb & 1023u) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }
        mbar_init(bar_qdo, 16);
        for (int i = 0; i < NK; ++i) { mbar_init(bar_kfull0 + 8 * i, 1); mbar_init(bar_kfree0 + 8 * i, 1); }
        for (int i = 0; i < NV; ++i) { mbar_init(bar_vfull0 + 8 * i, 1); mbar_init(bar_vfree0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_s0 + 8 * i, 1); mbar_init(bar_dp0 + 8 * i, 1); }
        mbar_init(bar_ds, 16); mbar_init(bar_sfree, 16); mbar_init(bar_done, 1);
        fence_mbar_init();
        tma_prefetch_desc(tmK); tma_prefetch_desc(tmV);
    }
    if (warp == 16) tmem_alloc<512>(sb + C::OFF_TMEM_SLOT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t COL_Q = 0, COL_DO = 64, COL_BUF = 128, COL_DQ = 384;

    // ---- which query block (last = heaviest under causal first)
    const uint32_t per = p.Hq * p.B;
    const uint32_t nqb = (p.Sq + 127) / 128, nkb = (p.Sk + 127) / 128;
    const uint32_t irev = blockIdx.x / per;
    const uint32_t bh = blockIdx.x - irev * per;                     // b * Hq + hq
    const uint32_t i = p.causal ? (nqb - 1 - irev) : irev;
    const uint32_t b = bh / p.Hq, hq = bh - b * p.Hq;
    const uint32_t bkv = b * p.Hkv + hq / (p.Hq / p.Hkv);
    const uint32_t n = p.causal ? min(nkb, i + 1) : nkb;             // key blocks 0..n-1 (top-left causal)

    if (warp == 16) {
        // ===================================================== issuer
        if (elect_one()) {
            constexpr uint64_t HI_K = smem_desc_hi(16, 1024);
            constexpr uint64_t HI_MN = smem_desc_hi(C::CHUNK_BYTES, 1024);
            constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
            constexpr uint32_t HI_MN_HI = uint32_t(HI_MN >> 32), HI_MN_LO = uint32_t(HI_MN);
            auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
            constexpr uint32_t ID_KK = instr_desc_f16(BF16, 128, 128, false);
            constexpr uint32_t ID_KMN = instr_desc_f16(BF16, 128, D, true);
            Tracer tr(p.trace, 0, true);
            // ---- loads: K(jj) -> stage jj%NK once dQ(jj-NK) has released it, V(jj) -> stage jj%NV once dP(jj-NV) has.
            uint32_t kl = 0, kl_st = 0, kl_use = 0;                           // next K load, its stage, earlier fills of that stage
            uint32_t vl = 0, vl_st = 0, vl_use = 0;
            auto pump = [&]() {
                if (kl < n && (kl_use == 0 || mbar_try_wait<0>(bar_kfree0 + 8 * kl_st, (kl_use - 1) & 1))) {
                    const uint32_t bar = bar_kfull0 + 8 * kl_st;
                    mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        tma_load_3d(sK0 + kl_st * C::TILE_BYTES + c * C::CHUNK_BYTES, tmK, bar, c * 64, (int32_t)(kl * 128), (int32_t)bkv);
                    ++kl;
                    if (++kl_st == NK) { kl_st = 0; ++kl_use; }
                }
                if (vl < n && (vl_use == 0 || mbar_try_wait<0>(bar_vfree0 + 8 * vl_st, (vl_use - 1) & 1))) {
                    const uint32_t bar = bar_vfull0 + 8 * vl_st;
                    mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        tma_load_3d(sV0 + vl_st * C::TILE_BYTES + c * C::CHUNK_BYTES, tmV, bar, c * 64, (int32_t)(vl * 128), (int32_t)bkv);
                    ++vl;
                    if (++vl_st == NV) { vl_st = 0; ++vl_use; }
                }
            };
            auto wait = [&](uint32_t bar, uint32_t parity) {                 // blocking wait that keeps the loads flowing
                while (!mbar_try_wait<0>(bar, parity)) pump();
            };
            auto issue_s = [&](uint32_t j, uint32_t kst) {                   // S = Q K_j^T into buffer j&1 (A = Q in TMEM)
                const uint32_t sK = sK0 + kst * C::TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                    mma_ts(tmem + COL_BUF + 128 * (j & 1), tmem + COL_Q + kk * 8, mk(HI_K_HI, (HI_K_LO | (sK >> 4)) + off), ID_KK, kk > 0);
                }
                mma_commit(bar_s0 + 8 * (j & 1));
            };
            auto issue_dq = [&](uint32_t j, uint32_t kst) {                  // dQ += dS K_j (K = 128 keys, A = dS(j) inside buffer j&1)
                const uint32_t sK = sK0 + kst * C::TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    mma_ts(tmem + COL_DQ, tmem + COL_BUF + 128 * (j & 1) + 32 * (kk >> 1) + 8 * (kk & 1),
                           mk(HI_MN_HI, (HI_MN_LO | (sK >> 4)) + kk * 128), ID_KMN, (j > 0 || kk > 0) ? 1u : 0u);
                mma_commit(bar_kfree0 + 8 * kst);
            };
            for (int t = 0; t < NK; ++t) pump();                             // fill both rings
            wait(bar_qdo, 0);
            wait(bar_kfull0, 0);
            tc_fence_after();
            issue_s(0, 0);
            uint32_t ks = 0, kp = 0, vs = 0, vp = 0;                          // K / V stage of step j and the parity of its fill
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t ks_prev = (ks == 0) ? NK - 1 : ks - 1;                                         // of step j-1
                const uint32_t ks_next = (ks == NK - 1) ? 0 : ks + 1, kp_next = (ks == NK - 1) ? (kp ^ 1) : kp;   // of step j+1
                if (j > 0) {
                    wait(bar_ds, (j - 1) & 1);                               // dS(j-1) in TMEM
                    tc_fence_after();
                    tr.ev(13, j);
                    issue_dq(j - 1, ks_prev);
                    if (p.order & 1) { wait(bar_kfree0 + 8 * ks_prev, (ks == 0) ? (kp ^ 1) : kp); tr.ev(30, j); }
                }
                tr.ev(10, j);
                wait(bar_sfree, j & 1);                                      // S(j) is in registers
                tr.ev(11, j);
                wait(bar_vfull0 + 8 * vs, vp);
                tr.ev(12, j);
                tc_fence_after();
                {
                    const uint32_t sV = sV0 + vs * C::TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {                    // dP = dO V_j^T -> buffer j&1 (A = dO in TMEM)
                        const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                        mma_ts(tmem + COL_BUF + 128 * (j & 1), tmem + COL_DO + kk * 8, mk(HI_K_HI, (HI_K_LO | (sV >> 4)) + off), ID_KK, kk > 0);
                    }
                    mma_commit(bar_dp0 + 8 * (j & 1));
                    mma_commit(bar_vfree0 + 8 * vs);
                    if (p.order & 1) { tr.ev(31, j); wait(bar_dp0 + 8 * (j & 1), (j >> 1) & 1); tr.ev(32, j); }
                }
                if (j + 1 < n) {
                    wait(bar_kfull0 + 8 * ks_next, kp_next);
                    tc_fence_after();
                    tr.ev(14, j);
                    issue_s(j + 1, ks_next);                                 // buffer (j+1)&1: follows dQ(j-1) in the tensor pipe
                    if (p.order & 1) { tr.ev(33, j); wait(bar_s0 + 8 * ((j + 1) & 1), ((j + 1) >> 1) & 1); tr.ev(34, j); }
                }
                tr.ev(15, j);
                ks = ks_next; kp = kp_next;
                if (++vs == NV) { vs = 0; vp ^= 1; }
            }
            wait(bar_ds, (n - 1) & 1);
            tc_fence_after();
            issue_dq(n - 1, (n - 1) % NK);
            mma_commit(bar_done);
        }
    } else {
        // ===================================================== compute warps
        const uint32_t qt = warp >> 2;                               // key quarter: columns [32qt, 32qt+32)
        const uint32_t h = (warp >> 2) & 1;                          // epilogue (warps 0-7): D half
        const uint32_t r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const uint32_t row = i * 128 + r;
        const size_t stat_off = (size_t)bh * p.Sq;
        const bool row_ok = row < p.Sq;
        // ---- Q_i, dO_i: this thread's D/4 elements of its row -> TMEM (A operands), zero beyond Sq
        {
            constexpr int NR = D / 8;                                // 32-bit registers per thread and tensor
            uint32_t a[NR], g[NR];
            const size_t eoff = ((stat_off + (row_ok ? row : 0)) * D + (D / 4) * qt) * 2;
            const uint4* gq = reinterpret_cast<const uint4*>(static_cast<const char*>(p.q) + eoff);
            const uint4* gd = reinterpret_cast<const uint4*>(static_cast<const char*>(p.d_o) + eoff);
#pragma unroll
            for (int u = 0; u < NR / 4; ++u) {
                const uint4 x = row_ok ? gq[u] : make_uint4(0, 0, 0, 0);
                const uint4 y = row_ok ? gd[u] : make_uint4(0, 0, 0, 0);
                a[4 * u] = x.x; a[4 * u + 1] = x.y; a[4 * u + 2] = x.z; a[4 * u + 3] = x.w;
                g[4 * u] = y.x; g[4 * u + 1] = y.y; g[4 * u + 2] = y.z; g[4 * u + 3] = y.w;
            }
            if constexpr (NR == 16) {
                tmem_st16(tmem + lane_addr + COL_Q + NR * qt, a);
                tmem_st16(tmem + lane_addr + COL_DO + NR * qt, g);
            } else {
                tmem_st8(tmem + lane_addr + COL_Q + NR * qt, a);
                tmem_st8(tmem + lane_addr + COL_DO + NR * qt, g);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_qdo);
        }
        const float lse2 = row_ok ? p.lse[stat_off + row] * 1.4426950408889634f : 0.f;
        const float delta = row_ok ? p.delta[stat_off + row] : 0.f;
        Tracer tr(p.trace, 1 + (qt & 1), (warp == 0 || warp == 4) && lane == 0);
        for (uint32_t j = 0; j < n; ++j) {
            const uint32_t key0 = j * 128;
            const bool diag = p.causal && (i * 128 < key0 + 128);
            const bool masked = diag || key0 + 128 > p.Sk || i * 128 + 128 > p.Sq;
            const uint32_t tbuf = tmem + lane_addr + COL_BUF + 128 * (j & 1) + 32 * qt;
            // ---- P phase
            float pv[32];
            tr.ev(20, j);
            mbar_wait(bar_s0 + 8 * (j & 1), (j >> 1) & 1);
            tr.ev(21, j);
            tc_fence_after();
            {
                uint32_t s[32];
                tmem_ld32(tbuf, s);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sfree);               // one arrival per warp: the buffer may take dP(j)
                tr.ev(22, j);
                if (masked) p_from_s<true>(s, pv, p.scale_log2, lse2, key0 + 32 * qt, row, p.Sk, row_ok, diag);
                else p_from_s<false, 1>(s, pv, p.scale_log2, lse2, 0, 0, 0, true, false);   // 1 pair in 4 on the FMA pipe: -3 % (A/B, s20)
            }
            // ---- dS phase: dS = P o (dP - Delta) -> 16-bit, in place over the first 16 of this thread's 32 dP columns
            tr.ev(23, j);
            mbar_wait(bar_dp0 + 8 * (j & 1), (j >> 1) & 1);
            tr.ev(24, j);
            tc_fence_after();
            uint32_t pk[16];
            {
                uint32_t dp[32];
                tmem_ld32(tbuf, dp);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float2 a = __fmul2_rn(make_float2(pv[2 * e], pv[2 * e + 1]),
                                                __fadd2_rn(make_float2(__uint_as_float(dp[2 * e]), __uint_as_float(dp[2 * e + 1])), make_float2(-delta, -delta)));
                    pk[e] = pack2<BF16>(a.x, a.y);
                }
            }
            tr.ev(25, j);
            tmem_st16(tbuf, pk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ds);
            tr.ev(27, j);
        }
        // ---- epilogue: dQ (x scale) -> 16-bit -> global (this half's D/2 columns of the row: 64 or 128 contiguous bytes)
        mbar_wait(bar_done, 0);
        tc_fence_after();
        if (warp < 8) {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(p.dq_out) + ((stat_off + (row_ok ? row : 0)) * D + (D / 2) * h) * 2);
#pragma unroll 1
            for (int c = 0; c < D / 64; ++c) {
                uint32_t o[32];
                tmem_ld32(tmem + lane_addr + COL_DQ + (D / 2) * h + c * 32, o);
                tmem_wait_ld();
                if (row_ok) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint4 v;
                        v.x = pack2<BF16>(__uint_as_float(o[8 * u + 0]) * p.scale, __uint_as_float(o[8 * u + 1]) * p.scale);
                        v.y = pack2<BF16>(__uint_as_float(o[8 * u + 2]) * p.scale, __uint_as_float(o[8 * u + 3]) * p.scale);
                        v.z = pack2<BF16>(__uint_as_float(o[8 * u + 4]) * p.scale, __uint_as_float(o[8 * u + 5]) * p.scale);
                        v.w = pack2<BF16>(__uint_as_float(o[8 * u + 6]) * p.scale, __uint_as_float(o[8 * u + 7]) * p.scale);
                        dst[c * 4 + u] = v;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc<512>(tmem);
}

// =====================================================================================================
// dQ kernel v2 (round 2): same CTA decomposition as bwd_dq_body, but no tile ever waits for a buffer another tile still
// occupies.  v1 kept S(j), dP(j) and dS(j) in ONE buffer (j&1), so dP(j) could only be issued once S(j) was in registers
// and the compute warps then sat ~480 cycles per step waiting for it (trace: profiles/r1_bwd_v4_trace_dq.txt); here
//   S  always lands in [128,256), dP in [256,384), dS (16-bit) in its own 64 columns [64,128)
// -- the columns v1 spent on dO, which is now a shared-memory A operand (one TMA load per CTA; dP = dO V^T is an SS MMA).
// So  S(j+1) is issued as soon as S(j) is in registers, dP(j+1) as soon as dP(j) is, dQ(j) as soon as dS(j) is stored:
// all three are in flight or done before the compute warps need them, and the step is bound by the warps' own work.
//   TMEM: Q [0,64) | dS [64,128) | S [128,256) | dP [256,384) | dQ [384,384+D)
//   SMEM: K ring 4 (K_j lives from S(j) to dQ(j)) | V ring 2 | dO_i
//   tensor order per step j:  S(j+1)  dP(j+1)  dQ(j)
template <int D, bool BF16>
__device__ __forceinline__ void bwd_dq2_body(const CUtensorMap* tmK, const CUtensorMap* tmV, const CUtensorMap* tmdO, const BwdParams& p) {
    using C = aule_kp::BwdDq2Cfg<D>;
    constexpr int NK = C::NK, NV = C::NV;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_q = sb + C::OFF_BAR;            // compute -> issuer: Q_i is in TMEM (16 arrivals)
    const uint32_t bar_do = bar_q + 8;                 // dO_i landed
    const uint32_t bar_kfull0 = bar_do + 8;            // K stage s landed (+8s)
    const uint32_t bar_kfree0 = bar_kfull0 + 8 * NK;   // dQ(j) complete (commit): K stage j%NK free (+8s)
    const uint32_t bar_vfull0 = bar_kfree0 + 8 * NK;   // V stage s landed (+8s)
    const uint32_t bar_vfree0 = bar_vfull0 + 8 * NV;   // dP(j) complete (commit): V stage j%NV free (+8s)
    const uint32_t bar_s = bar_vfree0 + 8 * NV;        // S(j) complete (commit)
    const uint32_t bar_dp = bar_s + 8;                 // dP(j) complete (commit)
    const uint32_t bar_dsfree = bar_dp + 8;            // dQ(j) complete (commit): the dS columns may be rewritten
    const uint32_t bar_sfree = bar_dsfree + 8;         // compute -> issuer: S(j) in registers (16)
    const uint32_t bar_dpfree = bar_sfree + 8;         // compute -> issuer: dP(j) in registers (16)
    const uint32_t bar_ds = bar_dpfree + 8;            // compute -> issuer: dS(j) in TMEM (16)
    const uint32_t bar_done = bar_ds + 8;              // every MMA complete (commit)
    static_assert(8 * (2 + 2 * NK + 2 * NV + 7) <= C::BAR_BYTES, "barrier area too small");
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_SLOT);
    const uint32_t sK0 = sb + C::OFF_K, sV0 = sb + C::OFF_V, sdO = sb + C::OFF_DO;

    if (threadIdx.x == 0) {
        if (sb & 1023u) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }
        mbar_init(bar_q, 16); mbar_init(bar_do, 1);
        for (int i = 0; i < NK; ++i) { mbar_init(bar_kfull0 + 8 * i, 1); mbar_init(bar_kfree0 + 8 * i, 1); }
        for (int i = 0; i < NV; ++i) { mbar_init(bar_vfull0 + 8 * i, 1); mbar_init(bar_vfree0 + 8 * i, 1); }
        mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_dsfree, 1); mbar_init(bar_done, 1);
        mbar_init(bar_sfree, 16); mbar_init(bar_dpfree, 16); mbar_init(bar_ds, 16);
        fence_mbar_init();
        tma_prefetch_desc(tmK); tma_prefetch_desc(tmV); tma_prefetch_desc(tmdO);
    }
    if (warp == 16) tmem_alloc<512>(sb + C::OFF_TMEM_SLOT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t COL_Q = 0, COL_DS = 64, COL_S = 128, COL_DP = 256, COL_DQ = 384;

    // ---- which query block (last = heaviest under causal first)
    const uint32_t per = p.Hq * p.B;
    const uint32_t nqb = (p.Sq + 127) / 128, nkb = (p.Sk + 127) / 128;
    const uint32_t irev = blockIdx.x / per;
    const uint32_t bh = blockIdx.x - irev * per;                     // b * Hq + hq
    const uint32_t i = p.causal ? (nqb - 1 - irev) : irev;
    const uint32_t b = bh / p.Hq, hq = bh - b * p.Hq;
    const uint32_t bkv = b * p.Hkv + hq / (p.Hq / p.Hkv);
    // key blocks a row of this query block can see: row - win_left <= key <= row + win_right (causal: win_right = 0);
    // the kernel numbers them j = 0..n-1 from j0
    const uint32_t j0 = i * 128 > p.win_left ? (i * 128 - p.win_left) / 128 : 0;
    const uint32_t j_end = (uint32_t)min((uint64_t)nkb, ((uint64_t)i * 128 + 127 + p.win_right) / 128 + 1);
    const uint32_t n = j_end > j0 ? j_end - j0 : 0;

    if (warp >= 16) {
        // ===================================================== three issuer warps, one per MMA stream (one elected thread each).
        // The three streams only meet through the compute warps (S(j+1) needs "S(j) in registers", dP(j+1) "dP(j) in
        // registers", dQ(j) "dS(j) stored"), so one thread issuing all of them in a fixed order made every stream wait for the
        // slowest event (head-of-line blocking), and that thread shares its scheduler with four issue-bound compute warps
        // (trace: ~400 cycles per wait).  Warps 16 / 17 / 18 sit on three different schedulers.
        if (elect_one() && n > 0) {                                   // (n == 0: a query block beyond every key's window -- dQ = 0)
            constexpr uint64_t HI_K = smem_desc_hi(16, 1024);
            constexpr uint64_t HI_MN = smem_desc_hi(C::CHUNK_BYTES, 1024);
            constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
            constexpr uint32_t HI_MN_HI = uint32_t(HI_MN >> 32), HI_MN_LO = uint32_t(HI_MN);
            auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
            constexpr uint32_t ID_KK = instr_desc_f16(BF16, 128, 128, false);
            constexpr uint32_t ID_KMN = instr_desc_f16(BF16, 128, D, true);
            if (warp == 16) {
                // ---- S stream + K loads: K(m) -> stage m%NK once dQ(m-NK) has released it
                Tracer tr(p.trace, 0, true);
                uint32_t kl = 0, kl_st = 0, kl_use = 0;                       // next K load, its stage, earlier fills of that stage
                auto pump = [&]() {
                    if (kl < n && (kl_use == 0 || mbar_test(bar_kfree0 + 8 * kl_st, (kl_use - 1) & 1))) {
                        const uint32_t bar = bar_kfull0 + 8 * kl_st;
                        mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                        for (int c = 0; c < C::CHUNKS; ++c)
                            tma_load_3d(sK0 + kl_st * C::TILE_BYTES + c * C::CHUNK_BYTES, tmK, bar, c * 64, (int32_t)((j0 + kl) * 128), (int32_t)bkv);
                        ++kl;
                        if (++kl_st == NK) { kl_st = 0; ++kl_use; }
                    }
                };
                auto issue_s = [&](uint32_t j) {                             // S(j) = Q K_j^T (A = Q in TMEM)
                    const uint32_t sK = sK0 + (j % NK) * C::TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                        mma_ts(tmem + COL_S, tmem + COL_Q + kk * 8, mk(HI_K_HI, (HI_K_LO | (sK >> 4)) + off), ID_KK, kk > 0);
                    }
                    mma_commit(bar_s);
                };
                for (int t = 0; t < NK; ++t) pump();
                mbar_wait(bar_q, 0);
                mbar_wait(bar_kfull0, 0);
                tc_fence_after();
                issue_s(0);
                for (uint32_t j = 0; j + 1 < n; ++j) {
                    pump();
                    tr.ev(10, j);
                    mbar_wait(bar_sfree, j & 1);                             // S(j) is in registers
                    tr.ev(16, j);
                    while (!mbar_try_wait<0>(bar_kfull0 + 8 * ((j + 1) % NK), ((j + 1) / NK) & 1)) pump();
                    tc_fence_after();
                    issue_s(j + 1);
                    tr.ev(18, j);
                }
                while (kl < n) pump();                                       // (every load is requested before its consumer waits; nothing left here)
            } else if (warp == 17) {
                // ---- dP stream + dO / V loads: V(m) -> stage m%NV once dP(m-NV) has released it
                uint32_t vl = 0, vl_st = 0, vl_use = 0;
                auto pump = [&]() {
                    if (vl < n && (vl_use == 0 || mbar_test(bar_vfree0 + 8 * vl_st, (vl_use - 1) & 1))) {
                        const uint32_t bar = bar_vfull0 + 8 * vl_st;
                        mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
                        for (int c = 0; c < C::CHUNKS; ++c)
                            tma_load_3d(sV0 + vl_st * C::TILE_BYTES + c * C::CHUNK_BYTES, tmV, bar, c * 64, (int32_t)((j0 + vl) * 128), (int32_t)bkv);
                        ++vl;
                        if (++vl_st == NV) { vl_st = 0; ++vl_use; }
                    }
                };
                auto issue_dp = [&](uint32_t j) {                            // dP(j) = dO V_j^T (A = dO in shared memory)
                    const uint32_t sV = sV0 + (j % NV) * C::TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                        mma_ss(tmem + COL_DP, mk(HI_K_HI, (HI_K_LO | (sdO >> 4)) + off), mk(HI_K_HI, (HI_K_LO | (sV >> 4)) + off), ID_KK, kk > 0);
                    }
                    mma_commit(bar_dp);
                    mma_commit(bar_vfree0 + 8 * (j % NV));
                };
                mbar_expect_tx(bar_do, C::TILE_BYTES);
#pragma unroll
                for (int c = 0; c < C::CHUNKS; ++c) tma_load_3d(sdO + c * C::CHUNK_BYTES, tmdO, bar_do, c * 64, (int32_t)(i * 128), (int32_t)bh);
                for (int t = 0; t < NV; ++t) pump();
                mbar_wait(bar_do, 0);
                mbar_wait(bar_vfull0, 0);
                tc_fence_after();
                issue_dp(0);
                for (uint32_t j = 0; j + 1 < n; ++j) {
                    pump();
                    mbar_wait(bar_dpfree, j & 1);                            // dP(j) is in registers
                    while (!mbar_try_wait<0>(bar_vfull0 + 8 * ((j + 1) % NV), ((j + 1) / NV) & 1)) pump();
                    tc_fence_after();
                    issue_dp(j + 1);
                }
            } else if (warp == 18) {
                // ---- dQ stream
                for (uint32_t j = 0; j < n; ++j) {
                    const uint32_t sK = sK0 + (j % NK) * C::TILE_BYTES;
                    mbar_wait(bar_ds, j & 1);                                // dS(j) in TMEM
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)                           // dQ += dS(j) K_j (K = 128 keys, A = dS in TMEM)
                        mma_ts(tmem + COL_DQ, tmem + COL_DS + 8 * kk, mk(HI_MN_HI, (HI_MN_LO | (sK >> 4)) + kk * 128), ID_KMN, (j > 0 || kk > 0) ? 1u : 0u);
                    mma_commit(bar_kfree0 + 8 * (j % NK));
                    mma_commit(bar_dsfree);
                }
                mma_commit(bar_done);                                        // dQ(n-1) is the last MMA: its dS needed every S and dP
            }
        }
    } else {
        // ===================================================== compute warps
        const uint32_t qt = warp >> 2;                               // key quarter: columns [32qt, 32qt+32)
        const uint32_t h = (warp >> 2) & 1;                          // epilogue (warps 0-7): D half
        const uint32_t r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const uint32_t row = i * 128 + r;
        const size_t stat_off = (size_t)bh * p.Sq;
        const bool row_ok = row < p.Sq;
        // ---- Q_i: this thread's D/4 elements of its row -> TMEM (A operand of S = Q K^T), zero beyond Sq
        {
            constexpr int NR = D / 8;                                // 32-bit registers per thread
            uint32_t a[NR];
            // (padded head dims: rows are D_real elements apart in memory, columns >= D_real are zero -- D_real % 8 == 0, so a
            //  16-byte vector is either entirely real or entirely padding)
            const size_t eoff = ((stat_off + (row_ok ? row : 0)) * p.D_real + (D / 4) * qt) * 2;
            const uint4* gq = reinterpret_cast<const uint4*>(static_cast<const char*>(p.q) + eoff);
#pragma unroll
            for (int u = 0; u < NR / 4; ++u) {
                const uint4 x = (row_ok && (D / 4) * qt + 8 * u < p.D_real) ? gq[u] : make_uint4(0, 0, 0, 0);
                a[4 * u] = x.x; a[4 * u + 1] = x.y; a[4 * u + 2] = x.z; a[4 * u + 3] = x.w;
            }
            if constexpr (NR == 16) tmem_st16(tmem + lane_addr + COL_Q + NR * qt, a);
            else tmem_st8(tmem + lane_addr + COL_Q + NR * qt, a);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_q);
        }
        const float lse2 = row_ok ? p.lse[stat_off + row] * 1.4426950408889634f : 0.f;
        const float delta = row_ok ? p.delta[stat_off + row] : 0.f;
        const uint32_t tS = tmem + lane_addr + COL_S + 32 * qt, tDP = tmem + lane_addr + COL_DP + 32 * qt;
        const uint32_t tDS = tmem + lane_addr + COL_DS + 16 * qt;
        Tracer tr(p.trace, 1 + (qt & 1), (warp == 0 || warp == 4) && lane == 0);
        for (uint32_t j = 0; j < n; ++j) {
            const uint32_t key0 = (j0 + j) * 128;
            const bool edge = (uint64_t)key0 + 127 > (uint64_t)i * 128 + p.win_right || (uint64_t)i * 128 + 127 > (uint64_t)key0 + p.win_left;
            const bool masked = edge || key0 + 128 > p.Sk || i * 128 + 128 > p.Sq;
            // ---- P phase
            float pv[32];
            tr.ev(20, j);
            mbar_wait(bar_s, j & 1);
            tr.ev(21, j);
            tc_fence_after();
            {
                uint32_t s[32];
                tmem_ld32(tS, s);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sfree);               // the S columns may take S(j+1)
                tr.ev(22, j);
                if (masked) p_from_s_band(s, pv, p.scale_log2, lse2, key0 + 32 * qt, row, p.Sk, row_ok, p.win_left, p.win_right);
                else p_from_s<false, 1>(s, pv, p.scale_log2, lse2, 0, 0, 0, true, false);
            }
            // ---- dS phase: dS = P o (dP - Delta) -> 16-bit -> this quarter's 16 of the 64 dS columns
            tr.ev(23, j);
            mbar_wait(bar_dp, j & 1);
            tr.ev(24, j);
            tc_fence_after();
            uint32_t pk[16];
            {
                uint32_t dp[32];
                tmem_ld32(tDP, dp);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_dpfree);              // the dP columns may take dP(j+1)
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float2 a = __fmul2_rn(make_float2(pv[2 * e], pv[2 * e + 1]),
                                                __fadd2_rn(make_float2(__uint_as_float(dp[2 * e]), __uint_as_float(dp[2 * e + 1])), make_float2(-delta, -delta)));
                    pk[e] = pack2<BF16>(a.x, a.y);
                }
            }
            tr.ev(25, j);
            if (j > 0) mbar_wait(bar_dsfree, (j - 1) & 1);           // dQ(j-1) has read dS(j-1)
            tc_fence_after();
            tmem_st16(tDS, pk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ds);
            tr.ev(27, j);
        }
        // ---- epilogue: dQ (x scale) -> 16-bit -> global (this half's D/2 columns of the row: 64 or 128 contiguous bytes)
        if (n > 0) mbar_wait(bar_done, 0);
        tc_fence_after();
        if (warp < 8) {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(p.dq_out) + ((stat_off + (row_ok ? row : 0)) * p.D_real + (D / 2) * h) * 2);
#pragma unroll 1
            for (int c = 0; c < D / 64; ++c) {
                uint32_t o[32];
                if (n > 0) {
                    tmem_ld32(tmem + lane_addr + COL_DQ + (D / 2) * h + c * 32, o);
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = 0u;
                }
                if (row_ok) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if ((D / 2) * h + c * 32 + 8 * u >= p.D_real) break;       // padding columns
                        uint4 v;
                        v.x = pack2<BF16>(__uint_as_float(o[8 * u + 0]) * p.scale, __uint_as_float(o[8 * u + 1]) * p.scale);
                        v.y = pack2<BF16>(__uint_as_float(o[8 * u + 2]) * p.scale, __uint_as_float(o[8 * u + 3]) * p.scale);
                        v.z = pack2<BF16>(__uint_as_float(o[8 * u + 4]) * p.scale, __uint_as_float(o[8 * u + 5]) * p.scale);
                        v.w = pack2<BF16>(__uint_as_float(o[8 * u + 6]) * p.scale, __uint_as_float(o[8 * u + 7]) * p.scale);
                        dst[c * 4 + u] = v;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc<512>(tmem);
}

}  // namespace bwd100

#define AULE_BWD100(NAME, DD, BF)                                                                        \
    extern "C" __global__ void __launch_bounds__(544, 1) NAME(const __grid_constant__ CUtensorMap tmQ,    \
                                                              const __grid_constant__ CUtensorMap tmK,    \
                                                              const __grid_constant__ CUtensorMap tmV,    \
                                                              const __grid_constant__ CUtensorMap tmdO,   \
                                                              const __grid_constant__ CUtensorMap tmdK,   \
                                                              const __grid_constant__ CUtensorMap tmdV,   \
                                                              const aule_kp::BwdParams p) {               \
        bwd100::bwd_dkv_body<DD, BF>(&tmQ, &tmK, &tmV, &tmdO, &tmdK, &tmdV, p);                           \
    }
#define AULE_BWD100_DQ(NAME, DD, BF)                                                                     \
    extern "C" __global__ void __launch_bounds__(544, 1) NAME(const __grid_constant__ CUtensorMap tmK,    \
                                                              const __grid_constant__ CUtensorMap tmV,    \
                                                              const aule_kp::BwdParams p) {               \
        bwd100::bwd_dq_body<DD, BF>(&tmK, &tmV, p);                                                       \
    }

#define AULE_BWD100_T(NAME, DD, BF)                                                                      \
    extern "C" __global__ void __launch_bounds__(544, 1) NAME(const __grid_constant__ CUtensorMap tmQ,    \
                                                              const __grid_constant__ CUtensorMap tmK,    \
                                                              const __grid_constant__ CUtensorMap tmV,    \
                                                              const __grid_constant__ CUtensorMap tmdO,   \
                                                              const __grid_constant__ CUtensorMap tmdK,   \
                                                              const __grid_constant__ CUtensorMap tmdV,   \
                                                              const aule_kp::BwdParams p) {               \
        bwd100::bwd_dkv_t_body<DD, BF>(&tmQ, &tmK, &tmV, &tmdO, &tmdK, &tmdV, p);                         \
    }
AULE_BWD100_T(aule_bwd_dkvt_sm100_bf16_d128, 128, true)
AULE_BWD100_T(aule_bwd_dkvt_sm100_bf16_d64, 64, true)
AULE_BWD100_T(aule_bwd_dkvt_sm100_f16_d128, 128, false)
AULE_BWD100_T(aule_bwd_dkvt_sm100_f16_d64, 64, false)
#ifdef AULE_TUNING_VARIANTS
// the superseded v3 dK/dV kernel (P, dS staged through shared memory): tuning builds only, for A/B runs and as an independent
// data path in tests (aule_set_kernel_path bit 13)
AULE_BWD100(aule_bwd_sm100_bf16_d128, 128, true)
AULE_BWD100(aule_bwd_sm100_bf16_d64, 64, true)
AULE_BWD100(aule_bwd_sm100_f16_d128, 128, false)
AULE_BWD100(aule_bwd_sm100_f16_d64, 64, false)
#endif
#define AULE_BWD100_DQ2(NAME, DD, BF)                                                                    \
    extern "C" __global__ void __launch_bounds__(608, 1) NAME(const __grid_constant__ CUtensorMap tmK,    \
                                                              const __grid_constant__ CUtensorMap tmV,    \
                                                              const __grid_constant__ CUtensorMap tmdO,   \
                                                              const aule_kp::BwdParams p) {               \
        bwd100::bwd_dq2_body<DD, BF>(&tmK, &tmV, &tmdO, p);                                               \
    }
AULE_BWD100_DQ2(aule_bwd_dq_sm100_bf16_d128, 128, true)
AULE_BWD100_DQ2(aule_bwd_dq_sm100_bf16_d64, 64, true)
AULE_BWD100_DQ2(aule_bwd_dq_sm100_f16_d128, 128, false)
AULE_BWD100_DQ2(aule_bwd_dq_sm100_f16_d64, 64, false)
#ifdef AULE_TUNING_VARIANTS
// the superseded v1 dQ kernel (S, dP and dS of a step share one TMEM buffer; Q and dO both in TMEM): tuning builds only
// (aule_set_kernel_path bit 23)
AULE_BWD100_DQ(aule_bwd_dq1_sm100_bf16_d128, 128, true)
AULE_BWD100_DQ(aule_bwd_dq1_sm100_bf16_d64, 64, true)
AULE_BWD100_DQ(aule_bwd_dq1_sm100_f16_d128, 128, false)
AULE_BWD100_DQ(aule_bwd_dq1_sm100_f16_d64, 64, false)
#endif

// Delta_i = sum_d O_id dO_id (triton_flash.py:353-379).  HBM-bound (reads O and dO once: 4*D bytes per row at 16 bits):
// a thread owns 8 consecutive elements (one 16-byte load per tensor) of a row, L = 4 / 8 / 16 lanes share a row (D <= 32 /
// 64 / 128; D % 8 == 0 and 16-byte aligned pointers on this path), and every thread keeps 4 rows in flight (8 independent
// 16-byte loads) before it reduces -- the round-1 kernel (a warp per row, 4-byte loads) ran at 2-3.8 TB/s.
template <bool BF16>
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 x, y;
        if constexpr (BF16) {
            x = make_float2(__uint_as_float(aw[i] << 16), __uint_as_float(aw[i] & 0xffff0000u));
            y = make_float2(__uint_as_float(bw[i] << 16), __uint_as_float(bw[i] & 0xffff0000u));
        } else {
            x = __half22float2(*reinterpret_cast<const __half2*>(&aw[i]));
            y = __half22float2(*reinterpret_cast<const __half2*>(&bw[i]));
        }
        acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
    }
    return acc;
}
template <typename T, bool BF16>
__device__ __forceinline__ void delta_body(const T* o, const T* d_o, float* delta, uint64_t rows, uint32_t D) {
    constexpr int R = 4;                                            // rows in flight per thread
    const uint32_t L = D <= 32 ? 4u : (D <= 64 ? 8u : 16u);          // lanes per row
    const uint32_t rpw = 32u / L;                                    // rows per warp and pass
    const uint32_t lane = threadIdx.x & 31, sub = lane % L, rl = lane / L;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t row0 = warp * (uint64_t)(rpw * R) + rl;           // this thread's rows: row0 + i * rpw
    const bool col_ok = sub * 8 < D;
    uint4 a[R], b[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const uint64_t row = row0 + (uint64_t)i * rpw;
        a[i] = b[i] = make_uint4(0u, 0u, 0u, 0u);
        if (row < rows && col_ok) {
            a[i] = __ldg(reinterpret_cast<const uint4*>(o + row * D + sub * 8));
            b[i] = __ldg(reinterpret_cast<const uint4*>(d_o + row * D + sub * 8));
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
        float acc = dot8<BF16>(a[i], b[i]);
#pragma unroll
        for (int s = 8; s > 0; s >>= 1)
            if ((uint32_t)s < L) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        const uint64_t row = row0 + (uint64_t)i * rpw;
        if (sub == 0 && row < rows) delta[row] = acc;
    }
}
extern "C" __global__ void __launch_bounds__(256) aule_bwd_delta_bf16(const __nv_bfloat16* o, const __nv_bfloat16* d_o, float* delta, uint64_t rows, uint32_t D) {
    delta_body<__nv_bfloat16, true>(o, d_o, delta, rows, D);
}
extern "C" __global__ void __launch_bounds__(256) aule_bwd_delta_f16(const __half* o, const __half* d_o, float* delta, uint64_t rows, uint32_t D) {
    delta_body<__half, false>(o, d_o, delta, rows, D);
}
