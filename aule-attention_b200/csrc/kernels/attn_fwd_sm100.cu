// Fused FlashAttention-2 forward for sm_100a (B200): TMA -> SMEM -> tcgen05.mma -> TMEM.   (kernel v5, "stream")
//
// Replaces (behaviour, not code) the reference's forward kernels:
//   python/aule/triton_flash.py:62-235 (_flash_attn_fwd_kernel),
//   shaders/attention_f32.comp / attention_f32_fast.comp / attention_forward_f32.comp.
// Semantics kept: O = softmax(scale*QK^T + mask) V; top-left causal mask (j <= i,
// triton_flash.py:187); GQA kv_head = q_head / (Hq/Hkv) (:95-96); LSE = m + ln(l) (:232).
//
// One persistent CTA per SM (512 threads).  A work item is two 128-row query tiles that share their K/V blocks
// (GQA: the two q-heads 2h, 2h+1 of a KV group at the same 128 rows; otherwise 256 rows of one head).
//
//   warps 0-3   softmax, tile 0   (thread == query row: S TMEM -> registers, exp2, P -> TMEM)
//   warps 4-7   softmax, tile 1
//   warps 8-11  epilogue          (O: TMEM -> regs -> 1/l -> 256-bit global stores; LSE)
//   warp  12    MMA issuer        (one elected thread issues every tcgen05.mma)
//   warp  13    scheduler + TMA producer (one elected thread claims work items and issues every bulk tensor load)
//   warp  14    TMEM allocator
//
// What v5 changes over v4 (attn_fwd_sm100_v4.cu): the K/V blocks of ALL work items a CTA processes form one
// continuous stream of "slots".  The MMA thread, the TMA thread and the softmax warps walk that stream with
// cursors that cross work-item boundaries, so the steady-state issue order
//     PV_0(i)  QK_1(i+1)  PV_1(i)  QK_0(i+2)
// never drains at an item boundary: while the last blocks of item k are still in the softmax / PV stages, the first
// Q K^T of item k+1 (both tiles) are already issued.  v4 lost ~1.5 block times of the ping-pong per item
// (16.5 blocks per item on config C).  That needs (a) Q tiles of two items resident at once: a 3-slot Q ring
// (tile x of item k lives in slot (2k + x) mod 3), paid for by dropping the O staging tile -- the epilogue now
// stores O straight from registers with 256-bit global stores (a thread owns a row: 32-byte sectors are written
// whole); (b) the scheduler claiming items ahead of use: an 8-slot ring of DECODED work descriptors in shared
// memory (no role re-derives the item geometry; the decode's integer divisions are paid once, by the producer).
//
// TMEM (512 columns): S [0,128) shared by both tiles | P0 [128,192) | P1 [192,256) |
//                     O0 [256,256+D) | O1 [256+D,256+2D).
// S only lives from the end of Q K^T until the softmax warps have copied it to registers, so ONE S buffer serves
// both tiles and P gets columns of its own; the next Q K^T of a tile is issued as soon as the OTHER tile's softmax
// has drained S.  Hazards: P_t is rewritten for the next block only after pv_done[t]; the rare in-place O_t rescale
// waits for the same barrier; S is handed over through s_free; O_t is handed to the epilogue at the last PV of an
// item (o_full) and back before the first PV of the next (o_empty).
//
// SMEM (D=128): Q ring 3x32 KB | K/V ring 4x32 KB (load order K0 K1 V0 K2 V1 K3 ..., continuous across items) |
// row statistics 2 KB | work descriptors | mbarriers.  All operand tiles are [128 rows][64 elements] 128B-swizzled
// sub-tiles (TMA box 64x128): the K-major canonical UMMA layout for Q/K and the MN-major one for V.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sm100_ptx.cuh"
#include "kernel_params.h"

namespace fwd100 {
using namespace sm100;
using aule_kp::FwdParams;
template <int D> using Cfg = aule_kp::FwdCfg<D>;

// barrier indices
enum : int {
    B_QFULL = 0,     // [3] TMA -> MMA: Q ring slot landed
    B_QEMPTY = 3,    // [3] MMA -> TMA: last Q K^T that reads the slot is done (commit)
    B_SFULL = 6,     // [2] MMA -> softmax t: S holds Q_t K^T (commit)
    B_PFULL = 8,     // [2] softmax t -> MMA: P_t columns [0,48) written (keys 0..95)
    B_PFULLB = 10,   // [2] softmax t -> MMA: P_t columns [48,64) written (keys 96..127)
    B_PVDONE = 12,   // [2] MMA -> softmax t: PV_t complete (commit): P_t / O_t may be touched
    B_OFULL = 14,    // [2] MMA -> epilogue: last PV_t of the work item done (commit)
    B_OEMPTY = 16,   // [2] epilogue -> MMA: O_t drained from TMEM
    B_STFULL = 18,   // [2] softmax t -> epilogue: row statistics written
    B_STEMPTY = 20,  // [2] epilogue -> softmax t: row statistics consumed
    B_SFREE = 22,    // softmax (either tile) -> MMA: S copied to registers
    B_WKFULL = 23,   // [8] scheduler -> all roles: work descriptor slot published
    B_WKEMPTY = 31,  // [8] all roles -> scheduler: slot consumed (1 MMA + 256 softmax + 128 epilogue arrivals)
    B_KVFULL = 39    // [NS] full, then [NS] empty
};
constexpr int WK_SLOTS = 8;

struct Work {
    uint32_t bh, bkv, row0, n0;
    uint32_t n1;                 // 0 = stop sentinel
    uint32_t j0;                 // first K/V block (left edge of a sliding window; 0 otherwise)
    uint32_t dbh, drow;          // tile t covers q-head (bh + t*dbh), rows [row0 + t*drow, +128)
};

// Work items.  A "unit" is one (batch, kv-head) with the Hq/Hkv query heads that share it.  Units are scheduled in
// runs of `units_per_run` (chosen by the host so that a run's K/V stays L2-resident: without it every unit is in
// flight at once and K/V is re-fetched from HBM for every query block -- measured 1.65 GB of DRAM reads per
// config-C launch against 0.40 GB compulsory).  Inside a run items go heaviest-first (causal).
//  pair_heads: item = 128 query rows x the two q-heads (2h, 2h+1) of the unit -> both tiles walk the same K/V
//              blocks and have EQUAL trip counts.
//  otherwise : item = 256 query rows of one q-head (tile 1 needs one more block than tile 0 under causal).
__device__ __forceinline__ Work decode(const FwdParams& p, uint32_t w) {
    Work t;
    const uint32_t nkb = (p.Sk + 127) / 128;
    const uint32_t G = p.Hq / p.Hkv, NU = p.B * p.Hkv;
    const uint32_t ipu = p.pair_heads ? G / 2 : G;                   // items per (query level, unit)
    const uint32_t ipr = p.num_q_super * p.units_per_run * ipu;      // items per full run
    const uint32_t run = w / ipr, lw = w - run * ipr;
    const uint32_t uir = min(p.units_per_run, NU - run * p.units_per_run);   // units in this run (last may be short)
    const uint32_t per = uir * ipu;
    const uint32_t qrev = lw / per, rem = lw - qrev * per;
    const uint32_t unit = run * p.units_per_run + rem / ipu, hh = rem % ipu;
    const uint32_t b = unit / p.Hkv, hk = unit - b * p.Hkv;
    const bool heavy_last = p.win_right != aule_kp::kWinInf && p.win_left == aule_kp::kWinInf;   // plain causal: later rows are heavier
    const uint32_t ql = heavy_last ? (p.num_q_super - 1 - qrev) : qrev;    // heaviest first
    t.bkv = unit;
    // K/V block range of a 128-row tile starting at row r0: blocks holding a key in [r0 - win_left, r0 + 127 + win_right]
    auto first_blk = [&](uint32_t r0) -> uint32_t { return r0 > p.win_left ? min((r0 - p.win_left) / 128, nkb - 1) : 0u; };
    auto last_blk = [&](uint32_t r0) -> uint32_t {
        return p.win_right == aule_kp::kWinInf ? nkb - 1 : min(nkb - 1, (r0 + 127 + p.win_right) / 128);
    };
    if (p.pair_heads) {
        t.bh = b * p.Hq + hk * G + 2 * hh;
        t.row0 = ql * 128;
        t.j0 = first_blk(t.row0);
        t.n0 = t.n1 = max(last_blk(t.row0), t.j0) - t.j0 + 1;
        t.dbh = 1; t.drow = 0;
    } else {
        t.bh = b * p.Hq + hk * G + hh;
        t.row0 = ql * 256;
        t.j0 = first_blk(t.row0);                               // common start: tile 0's first row is the leftmost
        t.n0 = max(last_blk(t.row0), t.j0) - t.j0 + 1;          // tile 0: rows [row0, row0+128)
        t.n1 = max(last_blk(t.row0 + 128), t.j0) - t.j0 + 1;    // tile 1: rows [row0+128, row0+256); n0 <= n1 always
        t.dbh = 0; t.drow = 128;
    }
    return t;
}

// Work descriptors travel through an 8-slot ring in shared memory (32 bytes each), written by the scheduler thread.
__device__ __forceinline__ void ring_write(uint32_t ring_smem, uint32_t slot, const Work& w) {
    const uint32_t a = ring_smem + slot * 32;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w.bh), "r"(w.bkv), "r"(w.row0), "r"(w.n0) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 16), "r"(w.n1), "r"(w.j0), "r"(w.dbh), "r"(w.drow) : "memory");
}
__device__ __forceinline__ Work ring_read(uint32_t ring_smem, uint32_t slot) {
    Work w;
    const uint32_t a = ring_smem + slot * 32;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.bh), "=r"(w.bkv), "=r"(w.row0), "=r"(w.n0) : "r"(a) : "memory");
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.n1), "=r"(w.j0), "=r"(w.dbh), "=r"(w.drow) : "r"(a + 16) : "memory");
    return w;
}
// Consumer side: wait until item `it` of this CTA has been published, copy it.  false = stop sentinel.
__device__ __forceinline__ bool fetch_work(uint32_t bar_full0, uint32_t ring_smem, uint32_t it, Work& w) {
    const uint32_t slot = it & (WK_SLOTS - 1);
    mbar_wait(bar_full0 + 8 * slot, (it / WK_SLOTS) & 1);
    w = ring_read(ring_smem, slot);
    return w.n1 != 0;
}

struct Ring {
    uint32_t stage = 0, phase = 0;
    template <int NS> __device__ __forceinline__ void advance() {
        if (++stage == NS) { stage = 0; phase ^= 1; }
    }
};

// Bring-up tracer (aule_set_trace_buffer + kernel variant "_tr"): CTA 0 records (tag << 48 | clock64) events,
// one region of 4096 entries per traced thread (0 MMA issuer, 1 / 2 softmax tile 0 / 1 row 0, 3 TMA producer).
template <bool ON>
struct Tracer {
    unsigned long long* buf = nullptr;
    uint32_t n = 0;
    __device__ __forceinline__ Tracer(unsigned long long* base, uint32_t region, bool on) {
        if constexpr (ON) buf = (base && on && blockIdx.x == 0) ? base + region * 4096 : nullptr;
    }
    __device__ __forceinline__ void ev(uint32_t code, uint32_t step) {
        if constexpr (ON) {
            if (buf && n < 4096) buf[n++] = ((unsigned long long)((code << 8) | (step & 255u)) << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
        }
    }
};

// VAR (tuning bits of the softmax loop; A/B results in profiles/r2_fwd_softmax_loop.md):
//   1       wait for pv_done only before the third P chunk is stored (P[0:96] is held in registers meanwhile)   [shipped]
//   2       publish P once per block (all four chunks, one wait::st)
//   4       S is loaded in two halves, the first half's row maximum is computed under the second half's TMEM load
//   65536   scalar FFMA instead of FFMA2 for x = s*scale - m
//   131072  x = s*scale - m computed in a pass of its own (in place) before the exp2 loop instead of interleaved  [shipped]
//   256     softmax-only microbenchmark (tools/softmax_only.py): no MMA / TMA / epilogue, no barriers;
//           + 512 tile 0 only, + 1024 tile 1 half a block late; diagnostics that BREAK the result: 2048 no row maximum,
//           4096 no P stores, 8192 no S loads, 16384 no scale/offset, 32768 no row sum
template <int D, bool BF16, int EMUN, int EMUD, bool TRUNC_PACK, bool TRACE, int VAR>
__device__ __forceinline__ void fwd_body(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                         const FwdParams& p) {
    using C = Cfg<D>;
    constexpr int NS = C::NS;
    constexpr uint32_t HOT_HINT = 1000000u;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar = [&](int i) -> uint32_t { return sb + C::OFF_BAR + 8u * i; };
    float* sStat = reinterpret_cast<float*>(smem + C::OFF_STAT);          // [l0 | l1 | m0 | m1] x 128
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_SLOT);
    const uint32_t wring = sb + C::OFF_WORK;

    if (threadIdx.x == 0 && (sb & 1023u)) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }
    // VAR bit 524288: one mbarrier arrival per WARP (after __syncwarp) instead of one per thread.  An arrive from 32 lanes on
    // one address is 32 serialised shared-memory atomics; ~24 such warp-arrives per block pair compete with the operand
    // reads of the SS-form Q K^T MMAs, which need the full 128 B/clk of shared-memory bandwidth on their own.
    constexpr bool WARP_ARR = (VAR & 524288) != 0;
    constexpr uint32_t NARR = WARP_ARR ? 4 : 128;                   // arrivals of one 128-thread group

    if (warp == 13 && lane == 0) {
        for (int i = 0; i < 3; ++i) {
            mbar_init(bar(B_QFULL + i), 1);     // TMA tx
            mbar_init(bar(B_QEMPTY + i), 1);    // tcgen05.commit
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar(B_SFULL + t), 1);     // tcgen05.commit
            mbar_init(bar(B_PFULL + t), NARR);  // softmax threads
            mbar_init(bar(B_PFULLB + t), NARR); // softmax threads
            mbar_init(bar(B_PVDONE + t), 1);    // tcgen05.commit
            mbar_init(bar(B_OFULL + t), 1);     // tcgen05.commit
            mbar_init(bar(B_OEMPTY + t), NARR); // epilogue threads
            mbar_init(bar(B_STFULL + t), NARR); // softmax threads
            mbar_init(bar(B_STEMPTY + t), NARR);// epilogue threads
        }
        mbar_init(bar(B_SFREE), NARR);          // softmax threads of whichever tile owns S
        for (int i = 0; i < WK_SLOTS; ++i) {
            mbar_init(bar(B_WKFULL + i), 1);     // scheduler (TMA thread)
            mbar_init(bar(B_WKEMPTY + i), 1 + 3 * NARR);  // 1 MMA + 256 softmax + 128 epilogue threads
        }
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar(B_KVFULL + s), 1);
            mbar_init(bar(B_KVFULL + NS + s), 1);
        }
        fence_mbar_init();
        tma_prefetch_desc(tmQ); tma_prefetch_desc(tmK); tma_prefetch_desc(tmV);
    }
    if (warp == 14) tmem_alloc<512>(sb + C::OFF_TMEM_SLOT);
    if constexpr (C::ROWSUM_MMA) {
        // constant tile [128 keys][64 cols], MN-major 128B-swizzled like a V chunk: column 0 = 1.0, the rest 0
        uint4* ones = reinterpret_cast<uint4*>(smem + C::OFF_ONES);
        for (uint32_t i = threadIdx.x; i < C::CHUNK_BYTES / 16; i += blockDim.x) {
            const uint32_t krow = i >> 3, unit = i & 7;                 // 8 16-byte units per key row; column 0 lives in unit (0 ^ (krow & 7))
            const uint32_t one = BF16 ? 0x3F80u : 0x3C00u;
            ones[i] = make_uint4(unit == (krow & 7) ? one : 0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    constexpr uint64_t HI_K = smem_desc_hi(16, 1024);                   // K-major SW128 (Q, K)
    constexpr uint64_t HI_V = smem_desc_hi(C::CHUNK_BYTES, 1024);       // MN-major SW128 (V)
    constexpr uint32_t IDESC_QK = instr_desc_f16(BF16, 128, 128, false);
    constexpr uint32_t IDESC_PV = instr_desc_f16(BF16, 128, C::PV_N, true);

    constexpr bool SO = (VAR & 256) != 0;   // softmax-only microbenchmark: no MMA / TMA / epilogue, no barriers (tools/softmax_only.py)
    if (warp < 8) {
        // ===================================================== softmax warps
        reg_inc<C::REGS_SOFTMAX>();
        const uint32_t t = warp >> 2;                               // tile 0 / 1
        const uint32_t r = (warp & 3) * 32 + lane;                  // row within the tile == TMEM lane
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const uint32_t tS = tmem + lane_addr + C::COL_S;
        const uint32_t tP = tmem + lane_addr + (t ? C::COL_P1 : C::COL_P0);
        const uint32_t tO = tmem + lane_addr + (t ? C::COL_O1 : C::COL_O0);
        Tracer<TRACE> tr(p.trace, 1 + t, (warp & 3) == 0 && lane == 0);
        uint32_t g = 0, it = 0;                                     // g: blocks processed so far by this tile
        Work wk;
        auto WAIT = [&](uint32_t b, uint32_t parity) { if constexpr (!SO) mbar_wait<HOT_HINT>(b, parity); };
        auto ARRIVE = [&](uint32_t b) {
            if constexpr (!SO) {
                if constexpr (WARP_ARR) { __syncwarp(); if (lane == 0) mbar_arrive(b); }
                else mbar_arrive(b);
            }
        };
        if constexpr (SO) {                                         // benign scores: S = 0 everywhere
            uint32_t z[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_st32(tS + c * 32, z);
            tmem_wait_st();
            named_bar_sync(2, 256);
        }
        if constexpr ((VAR & 1024) != 0) {                          // anti-phase: tile 1 starts half a block late
            if (t == 1) { const long long w0 = clock64(); while (clock64() - w0 < 1400) {} }
        }
        const long long so_t0 = SO ? clock64() : 0;
        for (; SO ? (it < 1 && !((VAR & 512) && t == 1)) : fetch_work(bar(B_WKFULL), wring, it, wk); ++it) {
            if constexpr (SO) { wk.bh = wk.bkv = wk.j0 = wk.dbh = wk.drow = 0; wk.row0 = 1u << 24; wk.n0 = wk.n1 = 400; }
            ARRIVE(bar(B_WKEMPTY + (it & (WK_SLOTS - 1))));    // descriptor copied to registers
            const uint32_t n = t ? wk.n1 : wk.n0;
            const uint32_t trow0 = wk.row0 + t * wk.drow;
            const uint32_t grow = trow0 + r;                        // global query row
            float m_used = -INFINITY, l = 0.f;
            for (uint32_t j = 0; j < n; ++j, ++g) {
                tr.ev(10, g);
                WAIT(bar(B_SFULL + t), g & 1);
                tr.ev(11, g);
                tc_fence_after();
                const uint32_t jg = wk.j0 + j;                      // global K/V block index
                // a block needs masking when some row of the tile loses a column of it: right limit (causal diagonal / window),
                // ragged key tail, left limit (window)
                const bool need_mask = !SO && ((p.win_right != aule_kp::kWinInf && jg * 128 + 127 > trow0 + p.win_right) ||
                                               ((jg + 1) * 128 > p.Sk) ||
                                               (p.win_left != aule_kp::kWinInf && jg * 128 + p.win_left < trow0 + 127));
                uint32_t s[4][32];
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                bool half_done = false;
                if ((VAR & 4) && !need_mask) {                      // first half's maximum under the second half's TMEM load
                    tmem_ld32(tS, s[0]);
                    tmem_ld32(tS + 32, s[1]);
                    tmem_wait_ld();
                    tmem_ld32(tS + 64, s[2]);
                    tmem_ld32(tS + 96, s[3]);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            mx0 = fmaxf(mx0, __uint_as_float(s[c][i]));
                            mx1 = fmaxf(mx1, __uint_as_float(s[c][i + 1]));
                            mx2 = fmaxf(mx2, __uint_as_float(s[c][i + 2]));
                            mx3 = fmaxf(mx3, __uint_as_float(s[c][i + 3]));
                        }
                    half_done = true;
                } else if (!(VAR & 8192) || g == 0) {               // (diagnostic bit 8192: S stays in registers after block 0)
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, s[c]);
                }
                if (VAR & 8192) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(s[c][i]));   // keep the values opaque
                }
                tmem_wait_ld();
                tc_fence_before();
                ARRIVE(bar(B_SFREE));                               // S may be overwritten by the next Q K^T
                tr.ev(12, g);
                if (need_mask) {                                    // diagonal / ragged-tail / window-edge blocks only
                    const uint32_t lim = p.win_right != aule_kp::kWinInf ? min(grow + p.win_right, p.Sk - 1) : p.Sk - 1;   // last visible key
                    const int32_t thr = (int32_t)lim - (int32_t)(jg * 128);           // local columns > thr are masked (a suffix)
                    const int32_t lo = p.win_left != aule_kp::kWinInf ? (int32_t)grow - (int32_t)p.win_left - (int32_t)(jg * 128) : -1;   // local columns < lo are masked (a prefix)
                    // Per 32-column chunk: one bit per column ("alive"), then one bit test + select per element, and only in the
                    // chunks where some row of the warp loses a column (on a causal diagonal block that is ONE chunk per warp
                    // plus the fully masked ones).  The first version looped over the chunks with a run-time index and paid four
                    // selects per element: a masked block cost 2.2x a plain one, 6 % (config C) to 11 % (config B) of all blocks
                    // (config B 604 -> 674 TFLOP/s, config C +3 %).  Skipping the exp2 of fully masked chunks with a warp-uniform
                    // branch was measured too: it splits the unrolled exp loop into basic blocks and LOSES 6 % (B 635, C 1118).
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int32_t t_ = thr - c * 32, l_ = lo - c * 32;
                        const uint32_t hi_m = t_ >= 31 ? 0xffffffffu : (t_ < 0 ? 0u : (0xffffffffu >> (31 - t_)));    // bits 0..t_
                        const uint32_t lo_m = l_ <= 0 ? 0xffffffffu : (l_ > 31 ? 0u : (0xffffffffu << l_));          // bits l_..31
                        const uint32_t alive = hi_m & lo_m;
                        if (__any_sync(0xffffffffu, alive != 0xffffffffu)) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) s[c][i] = (alive & (1u << i)) ? s[c][i] : 0xff800000u;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < 2 && half_done) continue;
                    if ((VAR & 2048) && g > 0) continue;            // diagnostic (softmax-only runs): no row maximum
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        mx0 = fmaxf(mx0, __uint_as_float(s[c][i]));
                        mx1 = fmaxf(mx1, __uint_as_float(s[c][i + 1]));
                        mx2 = fmaxf(mx2, __uint_as_float(s[c][i + 2]));
                        mx3 = fmaxf(mx3, __uint_as_float(s[c][i + 3]));
                    }
                }
                const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_used);
                // Lazy rescale: adopt the new maximum only when it grew by more than 2^8 in the
                // exp2 domain; otherwise P stays <= 256, harmless in bf16/fp16 and fp32 sums.
                const bool grow_max = (m_new - m_used) * p.scale_log2 > 8.f;   // m_used = -inf -> true (unless m_new = -inf too: NaN -> false)
                bool pv_waited = false;
                if (__any_sync(0xffffffffu, grow_max)) {
                    const float alpha = grow_max ? ex2((m_used - m_new) * p.scale_log2) : 1.f;
                    l *= alpha;
                    if (grow_max) m_used = m_new;
                    if (j > 0) {
                        WAIT(bar(B_PVDONE + t), (g - 1) & 1);              // PV_t(j-1) complete: O_t is stable
                        pv_waited = true;
                        tc_fence_after();
#pragma unroll 1
                        for (int c = 0; c < D / 32; ++c) {
                            uint32_t o[32];
                            tmem_ld32(tO + c * 32, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st32(tO + c * 32, o);
                        }
                        if constexpr (C::ROWSUM_MMA) {                    // the row-sum column rides with O
                            uint32_t o[16];
                            tmem_ld16(tO + D, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st16(tO + D, o);
                        }
                    }
                }
                tr.ev(13, g);
                // bf16: P is packed by TRUNCATION (one PRMT instead of the quarter-rate F2FP); the exponent carries
                // +log2(1+2^-9) so that the truncated values are unbiased, and l is corrected by the same factor.
                const float neg_ms = ((m_used == -INFINITY) ? 0.f : -m_used * p.scale_log2) + ((BF16 && TRUNC_PACK) ? 0.0028150156f : 0.f);
                // P = exp2(s*scale_log2 - m*scale_log2), two values per instruction (f32x2). EMU4 of every 4
                // pairs may take a polynomial path (FMA/ALU pipes) instead of MUFU.EX2 (EMUN of every EMUD pairs).
                const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(neg_ms, neg_ms);
                float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
                if constexpr ((VAR & 131072) != 0) {                // scale/offset pass of its own, in place
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1])), sc2, nm2);
                            s[c][2 * i] = __float_as_uint(x.x); s[c][2 * i + 1] = __float_as_uint(x.y);
                        }
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(s[c][i]));
                }
                uint32_t pk[4][16];
                constexpr bool NOST = (VAR & 4096) != 0;            // diagnostic (softmax-only runs): P is not stored
                auto ST16 = [&](uint32_t a, const uint32_t* r) {
                    if constexpr (NOST) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) asm volatile("" ::"r"(r[i]));
                    } else {
                        tmem_st16(a, r);
                    }
                };
                auto WAIT_ST = [&]() { if constexpr (!NOST) tmem_wait_st(); };
                auto wait_pv = [&]() {                              // P_t is still being read by PV_t of the previous block until pv_done
                    tr.ev(14, g);
                    if (g > 0 && !pv_waited) WAIT(bar(B_PVDONE + t), (g - 1) & 1);
                    tr.ev(15, g);
                    tc_fence_after();
                };
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float2 x;
                        if constexpr ((VAR & 16384) || (VAR & 131072)) {          // (131072: x was computed in a pass of its own)
                            x = make_float2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1]));
                        } else if constexpr ((VAR & 65536) != 0) {                 // scalar FFMA instead of the packed form
                            x.x = fmaf(__uint_as_float(s[c][2 * i]), p.scale_log2, neg_ms);
                            x.y = fmaf(__uint_as_float(s[c][2 * i + 1]), p.scale_log2, neg_ms);
                        } else {
                            x = __ffma2_rn(make_float2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1])), sc2, nm2);
                        }
                        float2 e;
                        if ((i % EMUD) < EMUN) {
                            e = ex2_emu2(x);
                        } else {
                            e.x = ex2(x.x);
                            e.y = ex2(x.y);
                        }
                        if (!(VAR & 32768) && !C::ROWSUM_MMA) { if (i & 1) acc1 = __fadd2_rn(acc1, e); else acc0 = __fadd2_rn(acc0, e); }
                        pk[c][i] = (BF16 && TRUNC_PACK) ? __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632) : pack2<BF16>(e.x, e.y);
                    }
                    if constexpr (VAR & 2) {                        // one publish per block
                        if (c == 3) {
                            wait_pv();
                            ST16(tP, pk[0]); ST16(tP + 16, pk[1]); ST16(tP + 32, pk[2]); ST16(tP + 48, pk[3]);
                            WAIT_ST();
                            tc_fence_before();
                            ARRIVE(bar(B_PFULL + t));
                            ARRIVE(bar(B_PFULLB + t));
                            tr.ev(17, g);
                        }
                    } else {
                        if (c == 1 && !(VAR & 1)) {
                            // the first two chunks are computed under PV_t of the previous block and stored once it has finished
                            wait_pv();
                            ST16(tP, pk[0]);
                            ST16(tP + 16, pk[1]);
                        } else if (c == 2) {
                            if (VAR & 1) {
                                wait_pv();
                                ST16(tP, pk[0]);
                                ST16(tP + 16, pk[1]);
                            }
                            ST16(tP + 32, pk[2]);               // keys 0..95 ready: PV k-steps 0..5 may start
                            WAIT_ST();
                            tc_fence_before();
                            ARRIVE(bar(B_PFULL + t));
                            tr.ev(16, g);
                        } else if (c == 3) {
                            ST16(tP + 48, pk[3]);
                            WAIT_ST();
                            tc_fence_before();
                            ARRIVE(bar(B_PFULLB + t));
                            tr.ev(17, g);
                        }
                    }
                }
                const float2 acc = __fadd2_rn(acc0, acc1);
                l += acc.x + acc.y;
            }
            // hand the row statistics to the epilogue warps; a row that saw no visible key reports l = 0
            if constexpr (!SO) mbar_wait(bar(B_STEMPTY + t), (it & 1) ^ 1);
            sStat[t * 128 + r] = (m_used == -INFINITY) ? 0.f : ((BF16 && TRUNC_PACK) ? l * (1.f / 1.001953125f) : l);   // undo the 1+2^-9 bias carried by the exponents
            sStat[256 + t * 128 + r] = m_used;
            ARRIVE(bar(B_STFULL + t));
        }
        if constexpr (SO) {
            if (blockIdx.x == 0 && (warp & 3) == 0 && lane == 0 && p.trace) p.trace[t] = (unsigned long long)(clock64() - so_t0);
        }
    } else if (SO) {
        reg_dec<C::REGS_EPILOGUE>();
    } else if (warp < 12) {
        // ===================================================== epilogue warps
        reg_dec<C::REGS_EPILOGUE>();
        const uint32_t r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        uint32_t it = 0;
        Work wk;
        for (; fetch_work(bar(B_WKFULL), wring, it, wk); ++it) {
            auto ARRIVE = [&](uint32_t b) {
                if constexpr (WARP_ARR) { __syncwarp(); if (lane == 0) mbar_arrive(b); }
                else mbar_arrive(b);
            };
            ARRIVE(bar(B_WKEMPTY + (it & (WK_SLOTS - 1))));
            for (uint32_t t = 0; t < 2; ++t) {
                const uint32_t tO = tmem + lane_addr + (t ? C::COL_O1 : C::COL_O0);
                mbar_wait(bar(B_STFULL + t), it & 1);
                float l = sStat[t * 128 + r];
                const float m = sStat[256 + t * 128 + r];
                ARRIVE(bar(B_STEMPTY + t));
                const uint32_t grow = wk.row0 + t * wk.drow + r;
                const bool row_ok = grow < p.Sq;
                const size_t orow = (size_t)(wk.bh + t * wk.dbh) * p.Sq + grow;
                uint8_t* optr = reinterpret_cast<uint8_t*>(p.o) + orow * (size_t)(p.D_real * 2);
                mbar_wait(bar(B_OFULL + t), it & 1);
                tc_fence_after();
                if constexpr (C::ROWSUM_MMA) {                        // row sum = column D of O (sum of the P values the MMA actually used)
                    uint32_t ls[16];
                    tmem_ld16(tO + D, ls);
                    tmem_wait_ld();
                    l = (m == -INFINITY) ? 0.f : __uint_as_float(ls[0]);
                }
                const float inv = (l > 0.f) ? 1.f / l : 0.f;        // rows without a visible key: O = 0, LSE = -inf
#pragma unroll
                for (int c = 0; c < D / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld32(tO + c * 32, o);
                    tmem_wait_ld();
                    if (c == D / 32 - 1) {                          // O_t fully read: MMA may overwrite it
                        tc_fence_before();
                        ARRIVE(bar(B_OEMPTY + t));
                    }
                    // 32 fp32 -> 32 x 16-bit = 64 B of this row
                    uint32_t v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) v[u] = pack2<BF16>(__uint_as_float(o[2 * u]) * inv, __uint_as_float(o[2 * u + 1]) * inv);
                    if (row_ok) {
                        if (p.D_real == (uint32_t)D) {              // two 256-bit stores (whole 32-byte sectors)
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(optr + c * 64),
                                         "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(optr + c * 64 + 32),
                                         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
                        } else {                                    // padded head_dim (D_real % 8 == 0): 128-bit pieces inside the row
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if ((uint32_t)(c * 32 + u * 8 + 8) <= p.D_real)
                                    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(optr + c * 64 + u * 16),
                                                 "r"(v[4 * u]), "r"(v[4 * u + 1]), "r"(v[4 * u + 2]), "r"(v[4 * u + 3]) : "memory");
                        }
                    }
                }
                if (p.lse != nullptr && row_ok)
                    p.lse[orow] = (l > 0.f) ? m * p.scale + __logf(l) : -INFINITY;   // LSE = m + ln(l)
            }
        }
    } else {
        reg_dec<C::REGS_OTHER>();
        if (warp == 12) {
            // ================================================= MMA issuer
            if (elect_one()) {
                Tracer<TRACE> tr(p.trace, 0, true);
                Ring ring;                                          // next K/V ring slot to acquire
                uint32_t gpv0 = 0, gpv1 = 0;                        // PV_t issued so far (parity of p_full / p_fullb)
                uint32_t nqk = 0;                                   // Q K^T issued so far (parity of s_free)
                // Descriptors are (constant high word, low word = const | addr>>4); stepping along K
                // is an immediate add on the low word (the 14-bit address field never carries).
                constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
                constexpr uint32_t HI_V_HI = uint32_t(HI_V >> 32), HI_V_LO = uint32_t(HI_V);
                auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
                // A cursor walks the CTA's stream of K/V block slots: (item ordinal, block within the item).
                struct Cur { uint32_t it, j, n0, n1; bool valid; };
                auto cur_load = [&](Cur& c) {                       // c.it set: fetch the item's trip counts
                    Work w;
                    c.valid = fetch_work(bar(B_WKFULL), wring, c.it, w);
                    c.n0 = w.n0; c.n1 = w.n1; c.j = 0;
                };
                auto cur_next = [&](Cur c) -> Cur {
                    if (c.j + 1 < c.n1) { ++c.j; return c; }
                    ++c.it;
                    cur_load(c);
                    return c;
                };
                auto acquire = [&]() -> uint32_t {                  // wait for the next tile of the load order
                    const uint32_t st = ring.stage;
                    tr.ev(6, st);
                    mbar_wait<HOT_HINT>(bar(B_KVFULL + st), ring.phase);
                    tr.ev(7, st);
                    ring.advance<NS>();
                    return st;
                };
                auto issue_qk = [&](uint32_t t, const Cur& c, uint32_t kstage) {  // S = Q_t K^T (waits until S has been drained)
                    const uint32_t n = t ? c.n1 : c.n0;
                    const uint32_t qi = 2 * c.it + t, qslot = qi % 3;
                    if (c.j == 0) mbar_wait(bar(B_QFULL + qslot), (qi / 3) & 1);
                    tr.ev(4, nqk);
                    if (nqk > 0) mbar_wait<HOT_HINT>(bar(B_SFREE), (nqk - 1) & 1);
                    tr.ev(5, nqk);
                    ++nqk;
                    tc_fence_after();
                    const uint32_t a_lo = HI_K_LO | ((sb + C::OFF_Q + qslot * C::TILE_BYTES) >> 4);
                    const uint32_t b_lo = HI_K_LO | ((sb + C::OFF_KV + kstage * C::TILE_BYTES) >> 4);
                    const uint32_t d = tmem + C::COL_S;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                        mma_ss(d, mk(HI_K_HI, a_lo + off), mk(HI_K_HI, b_lo + off), IDESC_QK, kk > 0);
                    }
                    mma_commit(bar(B_SFULL + t));
                    if (c.j == n - 1) mma_commit(bar(B_QEMPTY + qslot));   // last Q K^T of the item for this tile
                };
                auto issue_pv = [&](uint32_t t, const Cur& c, uint32_t vstage) {   // O_t (+)= P_t V
                    uint32_t& gpv = t ? gpv1 : gpv0;
                    const uint32_t n = t ? c.n1 : c.n0;
                    const bool first = c.j == 0, last = c.j == n - 1;
                    uint32_t b_lo = HI_V_LO | ((sb + C::OFF_KV + vstage * C::TILE_BYTES) >> 4);
                    if constexpr (C::ROWSUM_MMA)                         // columns [64,80) of "V" come from the ones tile: LBO = its distance from this stage
                        b_lo = (b_lo & ~(0x3FFFu << 16)) | ((((C::OFF_ONES - C::OFF_KV) - vstage * C::TILE_BYTES) >> 4) << 16);
                    const uint32_t a = tmem + (t ? C::COL_P1 : C::COL_P0);
                    const uint32_t d = tmem + (t ? C::COL_O1 : C::COL_O0);
                    tr.ev(1 + 16 * t, gpv);
                    mbar_wait<HOT_HINT>(bar(B_PFULL + t), gpv & 1);
                    if (first) mbar_wait(bar(B_OEMPTY + t), (c.it & 1) ^ 1);   // epilogue drained the previous O_t
                    tr.ev(2 + 16 * t, gpv);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 6; ++kk)
                        mma_ts(d, a + kk * 8, mk(HI_V_HI, b_lo + kk * (2048 >> 4)), IDESC_PV, (!first || kk > 0) ? 1u : 0u);
                    mbar_wait<HOT_HINT>(bar(B_PFULLB + t), gpv & 1);
                    tr.ev(3 + 16 * t, gpv);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 6; kk < 8; ++kk)
                        mma_ts(d, a + kk * 8, mk(HI_V_HI, b_lo + kk * (2048 >> 4)), IDESC_PV, 1u);
                    ++gpv;
                    mma_commit(bar(B_PVDONE + t));
                    if (last) mma_commit(bar(B_OFULL + t));
                };
                // Slots i (PV), i+1 (QK of tile 1), i+2 (QK of tile 0) are in flight; in steady state the issue order is
                //     PV_0(i)  QK_1(i+1)  PV_1(i)  QK_0(i+2)
                // and the cursors cross work-item boundaries without draining the pipe.
                Cur cC; cC.it = 0;
                cur_load(cC);
                if (cC.valid) {
                    uint32_t kB = 0, kA = 0;
                    {   // prologue: QK_0(s0) QK_1(s0) QK_0(s1)
                        const uint32_t k0 = acquire();
                        issue_qk(0, cC, k0);
                        issue_qk(1, cC, k0);
                        mma_commit(bar(B_KVFULL + NS + k0));
                    }
                    Cur cB = cur_next(cC), cA = cB;
                    if (cB.valid) {
                        kB = acquire();
                        if (cB.j < cB.n0) issue_qk(0, cB, kB);
                        cA = cur_next(cB);
                    }
                    while (cC.valid) {
                        const uint32_t v = acquire();               // V(s_i)
                        if (cC.j < cC.n0) issue_pv(0, cC, v);
                        if (cB.valid) {
                            issue_qk(1, cB, kB);                    // last user of K(s_{i+1})
                            mma_commit(bar(B_KVFULL + NS + kB));
                        }
                        issue_pv(1, cC, v);
                        mma_commit(bar(B_KVFULL + NS + v));         // V(s_i) released
                        if (cB.valid && cA.valid) {
                            kA = acquire();                         // K(s_{i+2})
                            if (cA.j < cA.n0) issue_qk(0, cA, kA);
                        }
                        if (cC.j == cC.n1 - 1) mbar_arrive(bar(B_WKEMPTY + (cC.it & (WK_SLOTS - 1))));   // item fully issued
                        cC = cB; cB = cA; kB = kA;
                        if (cA.valid) cA = cur_next(cA);
                    }
                }
            }
        } else if (warp == 13) {
            // ================================================= scheduler + TMA producer
            if (elect_one()) {
                Tracer<TRACE> tr(p.trace, 3, true);
                Ring ring;
                // Work items are claimed just in time: the atomic for item k+1 is issued when the K cursor is two slots from
                // the end of item k (its latency hides under those slots) and its result is published when the cursor
                // crosses the boundary.  Claiming earlier (v5 first claimed 2-3 items ahead) turns the heaviest-first dynamic
                // schedule into a static one when a CTA only gets a handful of items: config D/8, 3.5 items per CTA, ran
                // 25 % slower with the CTAs up to 30 % out of balance.
                uint32_t published = 0;
                uint32_t w_pending = 0;
                bool requested = false;
                auto request = [&]() {
                    if (!requested) { w_pending = atomicAdd(p.sched_counter, 1u); requested = true; }
                };
                auto publish = [&]() {                              // item `published` <- decode(w_pending)
                    request();
                    requested = false;
                    const uint32_t slot = published & (WK_SLOTS - 1);
                    mbar_wait(bar(B_WKEMPTY + slot), ((published / WK_SLOTS) & 1) ^ 1);
                    const uint32_t w = w_pending;
                    Work wk;
                    if (w < p.num_tiles) {
                        wk = decode(p, w);
                    } else {
                        wk.bh = wk.bkv = wk.row0 = wk.n0 = wk.n1 = wk.j0 = wk.dbh = wk.drow = 0;
                    }
                    ring_write(wring, slot, wk);
                    mbar_arrive(bar(B_WKFULL + slot));
                    ++published;
                };
                struct Cur { uint32_t it, j, n1, bkv, j0, bh, row0, dbh, drow; bool valid; };
                auto cur_load = [&](Cur& c) {                       // c.it < published
                    const Work w = ring_read(wring, c.it & (WK_SLOTS - 1));
                    c.j = 0; c.n1 = w.n1; c.bkv = w.bkv; c.j0 = w.j0; c.bh = w.bh; c.row0 = w.row0; c.dbh = w.dbh; c.drow = w.drow;
                    c.valid = w.n1 != 0;
                };
                auto load_kv = [&](const CUtensorMap* map, uint32_t j, uint32_t bkv) {
                    tr.ev(8, ring.stage);
                    mbar_wait(bar(B_KVFULL + NS + ring.stage), ring.phase ^ 1);
                    tr.ev(9, ring.stage);
                    const uint32_t full = bar(B_KVFULL + ring.stage);
                    const uint32_t dst = sb + C::OFF_KV + ring.stage * C::TILE_BYTES;
                    mbar_expect_tx(full, C::TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        tma_load_3d(dst + c * C::CHUNK_BYTES, map, full, c * 64, (int32_t)(j * 128), (int32_t)bkv);
                    ring.advance<NS>();
                };
                auto load_q = [&](const Cur& c, uint32_t t) {
                    const uint32_t qi = 2 * c.it + t, qslot = qi % 3;
                    mbar_wait(bar(B_QEMPTY + qslot), ((qi / 3) & 1) ^ 1);
                    const uint32_t full = bar(B_QFULL + qslot);
                    mbar_expect_tx(full, C::TILE_BYTES);
#pragma unroll
                    for (int ch = 0; ch < C::CHUNKS; ++ch)
                        tma_load_3d(sb + C::OFF_Q + qslot * C::TILE_BYTES + ch * C::CHUNK_BYTES, tmQ, full, ch * 64,
                                    (int32_t)(c.row0 + t * c.drow), (int32_t)(c.bh + t * c.dbh));
                };
                auto load_k = [&](const Cur& c) {                   // K of slot c; an item's Q tiles ride with its first K
                    if (c.j == 0) load_q(c, 0);
                    load_kv(tmK, c.j0 + c.j, c.bkv);
                    if (c.j == 0) load_q(c, 1);
                };
                auto advance = [&](Cur& c, bool leading) {
                    ++c.j;
                    if (leading && c.j + 2 >= c.n1) request();      // claim the next item two slots before this one ends
                    if (c.j < c.n1) return;
                    ++c.it;
                    if (leading) publish();
                    cur_load(c);
                };
                publish();
                Cur kc; kc.it = 0; cur_load(kc);
                Cur vc = kc;
                if (kc.valid) {
                    load_k(kc); advance(kc, true);
                    if (kc.valid) { load_k(kc); advance(kc, true); }
                    while (vc.valid) {                              // load order K0 K1 V0 K2 V1 K3 ... over the whole stream
                        load_kv(tmV, vc.j0 + vc.j, vc.bkv); advance(vc, false);
                        if (kc.valid) { load_k(kc); advance(kc, true); }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 14) tmem_dealloc<512>(tmem);
}

}  // namespace fwd100

#define AULE_FWD100(NAME, DD, BF, EMUN, EMUD, TP, TR, VAR)                                                         \
    extern "C" __global__ void __launch_bounds__(512, 1) NAME(const __grid_constant__ CUtensorMap tmQ,   \
                                                              const __grid_constant__ CUtensorMap tmK,   \
                                                              const __grid_constant__ CUtensorMap tmV,   \
                                                              const aule_kp::FwdParams p) {              \
        fwd100::fwd_body<DD, BF, EMUN, EMUD, TP, TR, VAR>(&tmQ, &tmK, &tmV, p);                                      \
    }

#ifndef AULE_FWD_EMUN
#define AULE_FWD_EMUN 1          // polynomial-exp2 pairs: EMUN of every EMUD pairs in the shipped kernels
#define AULE_FWD_EMUD 4
#endif
#ifndef AULE_FWD_VAR
#define AULE_FWD_VAR (1 + 131072)   // shipped softmax loop: late pv_done wait + scale/offset pass of its own (see VAR above)
#endif
AULE_FWD100(aule_fwd_sm100_bf16_d128, 128, true, AULE_FWD_EMUN, AULE_FWD_EMUD, true, false, AULE_FWD_VAR)
AULE_FWD100(aule_fwd_sm100_bf16_d64, 64, true, AULE_FWD_EMUN, AULE_FWD_EMUD, true, false, AULE_FWD_VAR)
AULE_FWD100(aule_fwd_sm100_f16_d128, 128, false, AULE_FWD_EMUN, AULE_FWD_EMUD, false, false, AULE_FWD_VAR)
AULE_FWD100(aule_fwd_sm100_f16_d64, 64, false, AULE_FWD_EMUN, AULE_FWD_EMUD, false, false, AULE_FWD_VAR)
#ifdef AULE_TUNING_VARIANTS
// bring-up / tuning builds only (make EXTRA_NVFLAGS=-DAULE_TUNING_VARIANTS): selected with aule_set_kernel_path(16 + v)
#define AULE_FWD_VARIANT(V, EMUN, EMUD, TR, VAR)                                                   \
    AULE_FWD100(aule_fwd_sm100_bf16_d128_e##V, 128, true, EMUN, EMUD, true, TR, VAR)               \
    AULE_FWD100(aule_fwd_sm100_bf16_d64_e##V, 64, true, EMUN, EMUD, true, TR, VAR)
AULE_FWD_VARIANT(0, 1, 4, true, AULE_FWD_VAR)  // pipeline tracer compiled in (tools/fwd_trace.py)
AULE_FWD_VARIANT(1, 1, 4, false, 1)            // interleaved scale/offset (round-2 first version)
AULE_FWD_VARIANT(2, 1, 8, false, AULE_FWD_VAR) // 12.5 % polynomial exp2
AULE_FWD_VARIANT(3, 0, 4, false, AULE_FWD_VAR) // MUFU only
// softmax-only microbenchmarks (tools/softmax_only.py)
AULE_FWD_VARIANT(4, 1, 4, false, 256 + AULE_FWD_VAR)
AULE_FWD_VARIANT(5, 1, 4, false, 256 + AULE_FWD_VAR + 512)
AULE_FWD_VARIANT(6, 1, 4, false, 256 + 1)
AULE_FWD_VARIANT(7, 1, 4, false, 256 + 1 + 2048)
AULE_FWD_VARIANT(8, 1, 4, false, 256 + 1 + 16384)
AULE_FWD_VARIANT(9, 1, 4, false, 256 + 1 + 32768)
AULE_FWD_VARIANT(10, 1, 4, false, 256 + 1 + 4096)
AULE_FWD_VARIANT(11, 0, 4, false, 256 + 1)
AULE_FWD_VARIANT(12, 1, 4, false, 256 + 1 + 65536)
AULE_FWD_VARIANT(13, 1, 4, false, AULE_FWD_VAR + 524288)   // one mbarrier arrival per warp
AULE_FWD_VARIANT(14, 1, 4, true, AULE_FWD_VAR + 524288)
AULE_FWD_VARIANT(15, 1, 8, false, AULE_FWD_VAR + 524288)
#endif
