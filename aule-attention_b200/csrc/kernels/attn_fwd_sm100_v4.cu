// Fused FlashAttention-2 forward for sm_100a (B200), kernel v4: superseded by v5 (attn_fwd_sm100.cu); compiled only
// into tuning builds (-DAULE_TUNING_VARIANTS) as aule_fwd4_sm100_* for A/B runs (aule_set_kernel_path bit 14).
//
// Replaces (behaviour, not code) the reference's forward kernels:
//   python/aule/triton_flash.py:62-235 (_flash_attn_fwd_kernel),
//   shaders/attention_f32.comp / attention_f32_fast.comp / attention_forward_f32.comp.
// Semantics kept: O = softmax(scale*QK^T + mask) V; top-left causal mask (j <= i,
// triton_flash.py:187); GQA kv_head = q_head / (Hq/Hkv) (:95-96); LSE = m + ln(l) (:232).
//
// One persistent CTA per SM (512 threads) looping over work items (256 query rows = two
// 128-row tiles of one (batch, q-head)), heaviest first, dealt in snake order:
//
//   warps 0-3   softmax, tile 0   (thread == query row: S TMEM -> registers, exp2, P -> TMEM)
//   warps 4-7   softmax, tile 1
//   warps 8-11  epilogue          (O: TMEM -> regs -> 1/l -> SMEM -> TMA store; LSE)
//   warp  12    MMA issuer        (one elected thread issues every tcgen05.mma)
//   warp  13    TMA producer      (one elected thread issues every bulk tensor load)
//   warp  14    TMEM allocator
//
// TMEM (512 columns): S [0,128) shared by both tiles | P0 [128,192) | P1 [192,256) |
//                     O0 [256,256+D) | O1 [256+D,256+2D).
// S only lives from the end of Q K^T until the softmax warps have copied it to registers
// (~150 cycles), so ONE S buffer serves both tiles and P gets columns of its own.  That removes the
// S/P aliasing of v3 and with it the serial chain  softmax(j) -> P V -> Q K^T(j+1) -> softmax(j+1):
// the next Q K^T of a tile is issued as soon as the OTHER tile's softmax has drained S, and runs
// under this tile's exp phase, so the softmax warps (the MUFU-bound stage) never wait for it.
// Steady-state issue order of the MMA thread (all waits blocking, order == readiness order):
//     PV_0(j)  QK_1(j+1)  PV_1(j)  QK_0(j+2)  PV_0(j+1)  QK_1(j+2)  ...
// Hazards: P_t is rewritten for block j+1 only after pv_done[t] (commit after PV_t(j)); the rare
// in-place O_t rescale waits for the same barrier; S is handed over through s_free (128 arrivals
// after the tcgen05.ld of S completes).
//
// SMEM (D=128): Q 2x32 KB | K/V ring 4x32 KB (load order K0 K1 V0 K2 V1 K3 ...) | O staging 32 KB |
// row statistics 2 KB | mbarriers.  All operand tiles are [128 rows][64 elements] 128B-swizzled
// sub-tiles (TMA box 64x128): the K-major canonical UMMA layout for Q/K and the MN-major one for V.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sm100_ptx.cuh"
#include "kernel_params.h"

namespace fwd100v4 {
using namespace sm100;
using aule_kp::FwdParams;
template <int D> using Cfg = aule_kp::FwdCfg4<D>;

// barrier indices (pairs are indexed by tile)
enum : int {
    B_QFULL = 0,     // TMA -> MMA: Q_t landed
    B_QEMPTY = 2,    // MMA -> TMA: last QK_t of the work item done (commit)
    B_SFULL = 4,     // MMA -> softmax t: S holds Q_t K^T (commit)
    B_PFULL = 6,     // softmax t -> MMA: P_t columns [0,48) written (keys 0..95)
    B_PFULLB = 8,    // softmax t -> MMA: P_t columns [48,64) written (keys 96..127)
    B_PVDONE = 10,   // MMA -> softmax t: PV_t(j) complete (commit): P_t / O_t may be touched
    B_OFULL = 12,    // MMA -> epilogue: last PV_t of the work item done (commit)
    B_OEMPTY = 14,   // epilogue -> MMA: O_t drained from TMEM
    B_STFULL = 16,   // softmax t -> epilogue: row statistics written
    B_STEMPTY = 18,  // epilogue -> softmax t: row statistics consumed
    B_SFREE = 20,    // softmax (either tile) -> MMA: S copied to registers
    B_WKFULL = 21,   // scheduler -> all roles: work_ring[slot] holds the next work item (4 slots)
    B_WKEMPTY = 25,  // all roles -> scheduler: slot consumed (1 MMA + 256 softmax + 128 epilogue arrivals)
    B_KVFULL = 29    // + NS: kv_empty
};

struct Work {
    uint32_t bh, bkv, row0, n0, n1;
    uint32_t j0;                 // first K/V block (left edge of a sliding window; 0 otherwise)
    uint32_t dbh, drow;          // tile t covers q-head (bh + t*dbh), rows [row0 + t*drow, +128)
};

// Work items.  A "unit" is one (batch, kv-head) with the Hq/Hkv query heads that share it.  Units are scheduled in
// runs of `units_per_run` (chosen by the host so that a run's K/V stays L2-resident: without it every unit is in
// flight at once and K/V is re-fetched from HBM for every query block -- measured 1.65 GB of DRAM reads per
// config-C launch against 0.40 GB compulsory).  Inside a run items go heaviest-first (causal), so the snake
// deal below balances every run on its own.
//  pair_heads: item = 128 query rows x the two q-heads (2h, 2h+1) of the unit -> both tiles walk the same K/V
//              blocks and have EQUAL trip counts (no idle slot for tile 0).
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
    const uint32_t ql = p.causal ? (p.num_q_super - 1 - qrev) : qrev; // heaviest first under causal
    t.bkv = unit;
    t.j0 = 0;
    if (p.pair_heads) {
        t.bh = b * p.Hq + hk * G + 2 * hh;
        t.row0 = ql * 128;
        t.n0 = t.n1 = p.causal ? min(nkb, ql + 1) : nkb;
        t.dbh = 1; t.drow = 0;
    } else {
        t.bh = b * p.Hq + hk * G + hh;
        t.row0 = ql * 256;
        t.n0 = p.causal ? min(nkb, t.row0 / 128 + 1) : nkb;    // KV blocks tile 0 needs (diagonal included)
        t.n1 = p.causal ? min(nkb, t.row0 / 128 + 2) : nkb;    // n0 <= n1 always
        t.dbh = 0; t.drow = 128;
    }
    if (p.window > 0 && t.row0 + 1 > (uint32_t)p.window) {           // blocks entirely left of the window are skipped
        t.j0 = min((t.row0 + 1 - (uint32_t)p.window) / 128, t.n0 - 1);  // common start (tile 0's first row is the leftmost)
        t.n0 -= t.j0; t.n1 -= t.j0;
    }
    return t;
}

// Dynamic persistent schedule: the TMA-producer thread claims work items with an atomic counter (items are sorted
// run by run, heaviest first inside a run, see decode) and publishes them through a 4-slot ring in shared memory;
// every other role consumes the ring in order.  A claimed index >= num_tiles is the stop sentinel.
__device__ __forceinline__ bool get_work(const FwdParams& p, uint32_t bar_full0, uint32_t bar_empty0,
                                         const volatile uint32_t* ring, uint32_t it, uint32_t& w) {
    const uint32_t slot = it & 3;
    mbar_wait(bar_full0 + 8 * slot, (it >> 2) & 1);
    w = ring[slot];
    mbar_arrive(bar_empty0 + 8 * slot);
    return w < p.num_tiles;
}

struct Ring {
    uint32_t stage = 0, phase = 0;
    template <int NS> __device__ __forceinline__ void advance() {
        if (++stage == NS) { stage = 0; phase ^= 1; }
    }
};

template <int D, bool BF16, int EMU4, bool TRUNC_PACK, uint32_t HOT_HINT>
__device__ __forceinline__ void fwd_body(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                         const CUtensorMap* tmO, const FwdParams& p) {
    using C = Cfg<D>;
    constexpr int NS = C::NS;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar = [&](int i) -> uint32_t { return sb + C::OFF_BAR + 8u * i; };
    float* sStat = reinterpret_cast<float*>(smem + C::OFF_STAT);          // [l0 | l1 | m0 | m1] x 128
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_SLOT);
    volatile uint32_t* work_ring = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_WORK);

    if (threadIdx.x == 0 && (sb & 1023u)) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }

    if (warp == 13 && lane == 0) {
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar(B_QFULL + t), 1);     // TMA tx
            mbar_init(bar(B_QEMPTY + t), 1);    // tcgen05.commit
            mbar_init(bar(B_SFULL + t), 1);     // tcgen05.commit
            mbar_init(bar(B_PFULL + t), 128);   // softmax threads
            mbar_init(bar(B_PFULLB + t), 128);  // softmax threads
            mbar_init(bar(B_PVDONE + t), 1);    // tcgen05.commit
            mbar_init(bar(B_OFULL + t), 1);     // tcgen05.commit
            mbar_init(bar(B_OEMPTY + t), 128);  // epilogue threads
            mbar_init(bar(B_STFULL + t), 128);  // softmax threads
            mbar_init(bar(B_STEMPTY + t), 128); // epilogue threads
        }
        mbar_init(bar(B_SFREE), 128);           // softmax threads of whichever tile owns S
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar(B_WKFULL + i), 1);     // scheduler (TMA thread)
            mbar_init(bar(B_WKEMPTY + i), 385);  // 1 MMA + 256 softmax + 128 epilogue threads
        }
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar(B_KVFULL + s), 1);
            mbar_init(bar(B_KVFULL + NS + s), 1);
        }
        fence_mbar_init();
        tma_prefetch_desc(tmQ); tma_prefetch_desc(tmK); tma_prefetch_desc(tmV); tma_prefetch_desc(tmO);
    }
    if (warp == 14) tmem_alloc<512>(sb + C::OFF_TMEM_SLOT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    constexpr uint64_t HI_K = smem_desc_hi(16, 1024);                   // K-major SW128 (Q, K)
    constexpr uint64_t HI_V = smem_desc_hi(C::CHUNK_BYTES, 1024);       // MN-major SW128 (V)
    constexpr uint32_t IDESC_QK = instr_desc_f16(BF16, 128, 128, false);
    constexpr uint32_t IDESC_PV = instr_desc_f16(BF16, 128, D, true);

    if (warp < 8) {
        // ===================================================== softmax warps
        reg_inc<192>();
        const uint32_t t = warp >> 2;                               // tile 0 / 1
        const uint32_t r = (warp & 3) * 32 + lane;                  // row within the tile == TMEM lane
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const uint32_t tS = tmem + lane_addr + C::COL_S;
        const uint32_t tP = tmem + lane_addr + (t ? C::COL_P1 : C::COL_P0);
        const uint32_t tO = tmem + lane_addr + (t ? C::COL_O1 : C::COL_O0);
        uint32_t g = 0, it = 0;                                     // g: blocks processed so far by this tile
        for (uint32_t w; get_work(p, bar(B_WKFULL), bar(B_WKEMPTY), work_ring, it, w); ++it) {
            const Work wk = decode(p, w);
            const uint32_t n = t ? wk.n1 : wk.n0;
            const uint32_t trow0 = wk.row0 + t * wk.drow;
            const uint32_t grow = trow0 + r;                        // global query row
            float m_used = -INFINITY, l = 0.f;
            for (uint32_t j = 0; j < n; ++j, ++g) {
                mbar_wait<HOT_HINT>(bar(B_SFULL + t), g & 1);
                tc_fence_after();
                uint32_t s[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, s[c]);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(bar(B_SFREE));                          // S may be overwritten by the next Q K^T
                const uint32_t jg = wk.j0 + j;                      // global K/V block index
                const bool win_edge = p.window > 0 && jg * 128 + (uint32_t)p.window < trow0 + 128;   // some key left of a row's window
                const bool need_mask = (p.causal && jg * 128 + 127 > trow0) || ((jg + 1) * 128 > p.Sk) || win_edge;
                if (need_mask) {                                    // diagonal / ragged-tail / window-edge blocks only: kept rolled (I-cache)
                    const uint32_t lim = p.causal ? min(grow, p.Sk - 1) : p.Sk - 1;   // last visible key
                    const int32_t thr = (int32_t)lim - (int32_t)(jg * 128);           // local columns > thr are masked (a suffix)
                    const int32_t lo = p.window > 0 ? (int32_t)grow - p.window + 1 - (int32_t)(jg * 128) : -1;   // local columns < lo are masked (a prefix)
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        if (thr >= c * 32 + 31 && lo <= c * 32) continue;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const bool dead = (c * 32 + i > thr) || (c * 32 + i < lo);
                            // dynamic c: select the chunk without dynamic register indexing
                            if (c == 0) s[0][i] = dead ? 0xff800000u : s[0][i];
                            else if (c == 1) s[1][i] = dead ? 0xff800000u : s[1][i];
                            else if (c == 2) s[2][i] = dead ? 0xff800000u : s[2][i];
                            else s[3][i] = dead ? 0xff800000u : s[3][i];
                        }
                    }
                }
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        mx0 = fmaxf(mx0, __uint_as_float(s[c][i]));
                        mx1 = fmaxf(mx1, __uint_as_float(s[c][i + 1]));
                        mx2 = fmaxf(mx2, __uint_as_float(s[c][i + 2]));
                        mx3 = fmaxf(mx3, __uint_as_float(s[c][i + 3]));
                    }
                const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_used);
                // Lazy rescale: adopt the new maximum only when it grew by more than 2^8 in the
                // exp2 domain; otherwise P stays <= 256, harmless in bf16/fp16 and fp32 sums.
                const bool grow_max = (m_new - m_used) * p.scale_log2 > 8.f;   // m_used = -inf -> true
                bool pv_waited = false;
                if (__any_sync(0xffffffffu, grow_max)) {
                    const float alpha = grow_max ? ex2((m_used - m_new) * p.scale_log2) : 1.f;
                    l *= alpha;
                    if (grow_max) m_used = m_new;
                    if (j > 0) {
                        mbar_wait<HOT_HINT>(bar(B_PVDONE + t), (g - 1) & 1);   // PV_t(j-1) complete: O_t is stable
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
                    }
                }
                // bf16: P is packed by TRUNCATION (one PRMT instead of the quarter-rate F2FP); the exponent carries
                // +log2(1+2^-9) so that the truncated values are unbiased, and l is corrected by the same factor.
                const float neg_ms = ((m_used == -INFINITY) ? 0.f : -m_used * p.scale_log2) + ((BF16 && TRUNC_PACK) ? 0.0028150156f : 0.f);
                // P = exp2(s*scale_log2 - m*scale_log2), two values per instruction (f32x2). EMU4 of every 4
                // pairs may take a polynomial path (FMA/ALU pipes) instead of MUFU.EX2.
                const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(neg_ms, neg_ms);
                float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
                uint32_t pk[2][16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1])), sc2, nm2);
                        float2 e;
                        if ((i & 3) < EMU4) {
                            e = ex2_emu2(x);
                        } else {
                            e.x = ex2(x.x);
                            e.y = ex2(x.y);
                        }
                        if (i & 1) acc1 = __fadd2_rn(acc1, e); else acc0 = __fadd2_rn(acc0, e);
                        pk[c & 1][i] = (BF16 && TRUNC_PACK) ? __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632) : pack2<BF16>(e.x, e.y);
                    }
                    if (c == 1) {
                        // P_t is still being read by PV_t of the previous block until pv_done: the first
                        // two chunks are computed under that MMA and stored once it has finished.
                        if (g > 0 && !pv_waited) mbar_wait<HOT_HINT>(bar(B_PVDONE + t), (g - 1) & 1);
                        tc_fence_after();
                        tmem_st16(tP, pk[0]);
                        tmem_st16(tP + 16, pk[1]);
                    } else if (c == 2) {
                        tmem_st16(tP + 32, pk[0]);                   // keys 0..95 ready: PV k-steps 0..5 may start
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive(bar(B_PFULL + t));
                    } else if (c == 3) {
                        tmem_st16(tP + 48, pk[1]);
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive(bar(B_PFULLB + t));
                    }
                }
                const float2 acc = __fadd2_rn(acc0, acc1);
                l += acc.x + acc.y;
            }
            // hand the row statistics to the epilogue warps
            mbar_wait(bar(B_STEMPTY + t), (it & 1) ^ 1);
            sStat[t * 128 + r] = (BF16 && TRUNC_PACK) ? l * (1.f / 1.001953125f) : l;   // undo the 1+2^-9 bias carried by the exponents
            sStat[256 + t * 128 + r] = m_used;
            mbar_arrive(bar(B_STFULL + t));
        }
    } else if (warp < 12) {
        // ===================================================== epilogue warps
        reg_dec<64>();
        const uint32_t r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const bool issuer = (warp == 8 && lane == 0);
        const uint32_t sO = sb + C::OFF_O;
        uint32_t it = 0;
        for (uint32_t w; get_work(p, bar(B_WKFULL), bar(B_WKEMPTY), work_ring, it, w); ++it) {
            const Work wk = decode(p, w);
            for (uint32_t t = 0; t < 2; ++t) {
                const uint32_t tO = tmem + lane_addr + (t ? C::COL_O1 : C::COL_O0);
                mbar_wait(bar(B_STFULL + t), it & 1);
                const float l = sStat[t * 128 + r];
                const float m = sStat[256 + t * 128 + r];
                mbar_arrive(bar(B_STEMPTY + t));
                mbar_wait(bar(B_OFULL + t), it & 1);
                tc_fence_after();
                if (issuer) tma_store_wait_read<0>();               // the previous store has finished reading sO
                named_bar_sync(1, 128);
                const float inv = 1.f / l;
#pragma unroll
                for (int c = 0; c < D / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld32(tO + c * 32, o);
                    tmem_wait_ld();
                    if (c == D / 32 - 1) {                          // O_t fully read: MMA may overwrite it
                        tc_fence_before();
                        mbar_arrive(bar(B_OEMPTY + t));
                    }
                    // 32 fp32 -> 32 x 16-bit = 64 B = 4 swizzled 16-byte units of this row
                    const uint32_t chunk = (c * 32) / 64, unit0 = ((c * 32) % 64) / 8;
                    const uint32_t rowbase = sO + chunk * C::CHUNK_BYTES + r * 128;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t v0 = pack2<BF16>(__uint_as_float(o[8 * u + 0]) * inv, __uint_as_float(o[8 * u + 1]) * inv);
                        const uint32_t v1 = pack2<BF16>(__uint_as_float(o[8 * u + 2]) * inv, __uint_as_float(o[8 * u + 3]) * inv);
                        const uint32_t v2 = pack2<BF16>(__uint_as_float(o[8 * u + 4]) * inv, __uint_as_float(o[8 * u + 5]) * inv);
                        const uint32_t v3 = pack2<BF16>(__uint_as_float(o[8 * u + 6]) * inv, __uint_as_float(o[8 * u + 7]) * inv);
                        const uint32_t addr = rowbase + (((unit0 + u) ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (issuer) {
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        tma_store_3d(tmO, sO + c * C::CHUNK_BYTES, c * 64, (int32_t)(wk.row0 + t * wk.drow), (int32_t)(wk.bh + t * wk.dbh));
                    tma_store_commit();
                }
                const uint32_t grow = wk.row0 + t * wk.drow + r;
                if (p.lse != nullptr && grow < p.Sq)
                    p.lse[(size_t)(wk.bh + t * wk.dbh) * p.Sq + grow] = m * p.scale + __logf(l);   // LSE = m + ln(l)
            }
        }
        if (issuer) tma_store_wait_all<0>();
    } else {
        reg_dec<64>();
        if (warp == 12) {
            // ================================================= MMA issuer
            if (elect_one()) {
                Ring ring;                                          // next K/V ring slot to acquire
                uint32_t gpv0 = 0, gpv1 = 0;                        // PV_t issued so far (parity of p_full / p_fullb)
                uint32_t nqk = 0;                                   // Q K^T issued so far (parity of s_free)
                uint32_t it = 0;
                // Descriptors are (constant high word, low word = const | addr>>4); stepping along K
                // is an immediate add on the low word (the 14-bit address field never carries).
                constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
                constexpr uint32_t HI_V_HI = uint32_t(HI_V >> 32), HI_V_LO = uint32_t(HI_V);
                auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
                auto acquire = [&]() -> uint32_t {                  // wait for the next tile of the load order
                    const uint32_t st = ring.stage;
                    mbar_wait<HOT_HINT>(bar(B_KVFULL + st), ring.phase);
                    ring.advance<NS>();
                    return st;
                };
                auto issue_qk = [&](uint32_t t, uint32_t kstage) {  // S = Q_t K^T (waits until S has been drained)
                    if (nqk > 0) mbar_wait<HOT_HINT>(bar(B_SFREE), (nqk - 1) & 1);
                    ++nqk;
                    tc_fence_after();
                    const uint32_t a_lo = HI_K_LO | ((sb + C::OFF_Q + t * C::TILE_BYTES) >> 4);
                    const uint32_t b_lo = HI_K_LO | ((sb + C::OFF_KV + kstage * C::TILE_BYTES) >> 4);
                    const uint32_t d = tmem + C::COL_S;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                        mma_ss(d, mk(HI_K_HI, a_lo + off), mk(HI_K_HI, b_lo + off), IDESC_QK, kk > 0);
                    }
                    mma_commit(bar(B_SFULL + t));
                };
                auto issue_pv = [&](uint32_t t, uint32_t vstage, bool first, bool last) {   // O_t (+)= P_t V
                    uint32_t& gpv = t ? gpv1 : gpv0;
                    const uint32_t b_lo = HI_V_LO | ((sb + C::OFF_KV + vstage * C::TILE_BYTES) >> 4);
                    const uint32_t a = tmem + (t ? C::COL_P1 : C::COL_P0);
                    const uint32_t d = tmem + (t ? C::COL_O1 : C::COL_O0);
                    mbar_wait<HOT_HINT>(bar(B_PFULL + t), gpv & 1);
                    if (first) mbar_wait(bar(B_OEMPTY + t), (it & 1) ^ 1);   // epilogue drained the previous O_t
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 6; ++kk)
                        mma_ts(d, a + kk * 8, mk(HI_V_HI, b_lo + kk * (2048 >> 4)), IDESC_PV, (!first || kk > 0) ? 1u : 0u);
                    mbar_wait<HOT_HINT>(bar(B_PFULLB + t), gpv & 1);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 6; kk < 8; ++kk)
                        mma_ts(d, a + kk * 8, mk(HI_V_HI, b_lo + kk * (2048 >> 4)), IDESC_PV, 1u);
                    ++gpv;
                    mma_commit(bar(B_PVDONE + t));
                    if (last) mma_commit(bar(B_OFULL + t));
                };
                // Cross-item pipelining: tile 0's softmax warps would otherwise sit idle from their last P of an item until
                // the whole prologue of the next one has gone through the pipe.  In the LAST iteration of an item, right
                // after PV_0, the next item is fetched and its QK_0(0) issued (its K_0 is the next tile of the load order,
                // its Q_0 was requested when the last QK_0 of this item released the buffer), so S_0(0) of the next item
                // is ready about one MMA group after tile 0 finished.
                bool have_next = false, next_ok = false, qk0_done = false;
                uint32_t w_next = 0, k_pre = 0;
                for (uint32_t w;; ++it) {
                    if (have_next) { w = w_next; if (!next_ok) break; }
                    else if (!get_work(p, bar(B_WKFULL), bar(B_WKEMPTY), work_ring, it, w)) break;
                    have_next = false;
                    const Work wk = decode(p, w);
                    const uint32_t n0 = wk.n0, n1 = wk.n1;          // n0 <= n1, n1 >= 1
                    // ---- prologue: QK_0(0) QK_1(0) QK_0(1)
                    uint32_t k_cur;
                    if (qk0_done) {
                        k_cur = k_pre;                              // K_0 acquired and QK_0(0) issued under the previous item
                        qk0_done = false;
                    } else {
                        k_cur = acquire();                          // K_0
                        mbar_wait(bar(B_QFULL + 0), it & 1);
                        issue_qk(0, k_cur);
                        if (n0 == 1) mma_commit(bar(B_QEMPTY + 0));
                    }
                    mbar_wait(bar(B_QFULL + 1), it & 1);
                    issue_qk(1, k_cur);
                    if (n1 == 1) mma_commit(bar(B_QEMPTY + 1));
                    mma_commit(bar(B_KVFULL + NS + k_cur));         // K_0 released
                    uint32_t k_next = 0, k_next2 = 0;               // stages of K_{j+1}, K_{j+2}
                    if (n1 > 1) {
                        k_next = acquire();                         // K_1
                        if (1 < n0) {
                            issue_qk(0, k_next);
                            if (n0 == 2) mma_commit(bar(B_QEMPTY + 0));
                        }
                    }
                    // ---- main loop
                    for (uint32_t j = 0; j < n1; ++j) {
                        const uint32_t v = acquire();               // V_j
                        if (j < n0) issue_pv(0, v, j == 0, j == n0 - 1);
                        if (j == n1 - 1 && p.cross_item) {          // last block: start the next item's tile 0
                            next_ok = get_work(p, bar(B_WKFULL), bar(B_WKEMPTY), work_ring, it + 1, w_next);
                            have_next = true;
                            if (next_ok) {
                                const Work wn = decode(p, w_next);
                                k_pre = acquire();                  // K_0 of the next item: next tile of the load order
                                mbar_wait(bar(B_QFULL + 0), (it + 1) & 1);
                                issue_qk(0, k_pre);
                                if (wn.n0 == 1) mma_commit(bar(B_QEMPTY + 0));
                                qk0_done = true;
                            }
                        }
                        if (j + 1 < n1) {
                            issue_qk(1, k_next);                    // QK_1(j+1): last user of K_{j+1}
                            if (j + 1 == n1 - 1) mma_commit(bar(B_QEMPTY + 1));
                            mma_commit(bar(B_KVFULL + NS + k_next));
                        }
                        issue_pv(1, v, j == 0, j == n1 - 1);
                        mma_commit(bar(B_KVFULL + NS + v));         // V_j released
                        if (j + 2 < n1) {
                            k_next2 = acquire();                    // K_{j+2}
                            if (j + 2 < n0) {
                                issue_qk(0, k_next2);
                                if (j + 2 == n0 - 1) mma_commit(bar(B_QEMPTY + 0));
                            }
                        }
                        k_next = k_next2;
                    }
                }
            }
        } else if (warp == 13) {
            // ================================================= TMA producer
            if (elect_one()) {
                Ring ring;
                uint32_t it = 0;
                auto load_kv = [&](const CUtensorMap* map, uint32_t j, uint32_t bkv) {
                    mbar_wait(bar(B_KVFULL + NS + ring.stage), ring.phase ^ 1);
                    const uint32_t full = bar(B_KVFULL + ring.stage);
                    const uint32_t dst = sb + C::OFF_KV + ring.stage * C::TILE_BYTES;
                    mbar_expect_tx(full, C::TILE_BYTES);
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        tma_load_3d(dst + c * C::CHUNK_BYTES, map, full, c * 64, (int32_t)(j * 128), (int32_t)bkv);
                    ring.advance<NS>();
                };
                for (;; ++it) {
                    uint32_t w;
                    {   // claim and publish the next work item
                        const uint32_t slot = it & 3;
                        mbar_wait(bar(B_WKEMPTY + slot), ((it >> 2) & 1) ^ 1);
                        w = atomicAdd(p.sched_counter, 1u);
                        work_ring[slot] = w;
                        mbar_arrive(bar(B_WKFULL + slot));
                        if (w >= p.num_tiles) break;
                    }
                    const Work wk = decode(p, w);
                    load_kv(tmK, wk.j0, wk.bkv);
                    for (uint32_t t = 0; t < 2; ++t) {
                        mbar_wait(bar(B_QEMPTY + t), (it & 1) ^ 1);
                        const uint32_t full = bar(B_QFULL + t);
                        mbar_expect_tx(full, C::TILE_BYTES);
#pragma unroll
                        for (int c = 0; c < C::CHUNKS; ++c)
                            tma_load_3d(sb + C::OFF_Q + t * C::TILE_BYTES + c * C::CHUNK_BYTES, tmQ, full, c * 64,
                                        (int32_t)(wk.row0 + t * wk.drow), (int32_t)(wk.bh + t * wk.dbh));
                    }
                    if (wk.n1 > 1) load_kv(tmK, wk.j0 + 1, wk.bkv);   // same order as the MMA thread acquires
                    for (uint32_t j = 0; j < wk.n1; ++j) {
                        load_kv(tmV, wk.j0 + j, wk.bkv);
                        if (j + 2 < wk.n1) load_kv(tmK, wk.j0 + j + 2, wk.bkv);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 14) tmem_dealloc<512>(tmem);
}

}  // namespace fwd100v4

#define AULE_FWD100V4(NAME, DD, BF, EMU, TP, HINT)                                                        \
    extern "C" __global__ void __launch_bounds__(512, 1) NAME(const __grid_constant__ CUtensorMap tmQ,   \
                                                              const __grid_constant__ CUtensorMap tmK,   \
                                                              const __grid_constant__ CUtensorMap tmV,   \
                                                              const __grid_constant__ CUtensorMap tmO,   \
                                                              const aule_kp::FwdParams p) {              \
        fwd100v4::fwd_body<DD, BF, EMU, TP, HINT>(&tmQ, &tmK, &tmV, &tmO, p);                              \
    }

#ifndef AULE_FWD_TRUNC
#define AULE_FWD_TRUNC true      // bf16 P packed by bias-compensated truncation (PRMT) instead of F2FP
#endif
#ifndef AULE_FWD_HINT
#define AULE_FWD_HINT 1000000u   // suspend hint (ns) of the waits on the softmax <-> MMA critical path
#endif
AULE_FWD100V4(aule_fwd4_sm100_bf16_d128, 128, true, 1, AULE_FWD_TRUNC, AULE_FWD_HINT)
AULE_FWD100V4(aule_fwd4_sm100_bf16_d64, 64, true, 1, AULE_FWD_TRUNC, AULE_FWD_HINT)
