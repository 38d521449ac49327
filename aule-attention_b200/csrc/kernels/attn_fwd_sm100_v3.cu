// Fused FlashAttention-2 forward for sm_100a (B200): TMA -> SMEM -> tcgen05.mma -> TMEM.
//
// Replaces (behaviour, not code) the reference's forward kernels:
//   python/aule/triton_flash.py:62-235 (_flash_attn_fwd_kernel),
//   shaders/attention_f32.comp / attention_f32_fast.comp / attention_forward_f32.comp.
// Semantics kept: O = softmax(scale*QK^T + mask) V; top-left causal mask (j <= i,
// triton_flash.py:187); GQA kv_head = q_head / (Hq/Hkv) (:95-96); LSE = m + ln(l) (:232).
//
// One persistent CTA per SM (512 threads), each looping over work items
// (256 query rows = two 128-row tiles of one (batch, q-head)), heaviest first:
//
//   warps 0-3   softmax for tile 0   (thread == query row; S read from TMEM, P written
//   warps 4-7   softmax for tile 1    back to TMEM as bf16/fp16; lazy O rescale in place)
//   warps 8-11  epilogue              (O: TMEM -> regs -> 1/l -> SMEM -> TMA store; LSE)
//   warp  12    MMA issuer            (one elected thread issues every tcgen05.mma)
//   warp  13    TMA producer          (one elected thread issues every bulk tensor load)
//   warp  14    TMEM allocator
//
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D).
// P_t aliases the first 64 columns of S_t (two 16-bit values per 32-bit column).
// The two tiles ping-pong: while the softmax warps of one tile run exp2 on S_t(j), the
// tensor core runs P V and the next Q K^T of the other tile
//   (issue order: QK0(0) QK1(0) | PV0(0) QK0(1) | PV1(0) QK1(1) | PV0(1) QK0(2) | ...).
// tcgen05.mma's issued by one thread execute in order, which is what protects the
// S_t/P_t aliasing (PV_t(j) is always issued before QK_t(j+1)).
//
// SMEM (D=128): Q 2x32 KB, K/V ring 3x32 KB (K and V tiles share the ring, load order
// K0 V0 K1 V1 ...), O staging 2x32 KB, row statistics 2 KB, mbarriers.  All operand tiles are
// [128 rows][64 elements] 128B-swizzled sub-tiles (TMA box 64x128), which is at once the
// K-major canonical layout for Q/K and the MN-major canonical layout for V.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sm100_ptx.cuh"
#include "kernel_params.h"

namespace fwd100v3 {
using namespace sm100;
using aule_kp::FwdParams;
template <int D> using Cfg = aule_kp::FwdCfgV3<D>;

// barrier indices
enum : int { B_QFULL = 0, B_QEMPTY = 2, B_SFULL = 4, B_PFULL = 6 /* P columns [0,48): keys 0..95 */, B_OFULL = 8,
             B_OEMPTY = 10, B_STFULL = 12, B_STEMPTY = 14, B_PFULLB = 16 /* P columns [48,64): keys 96..127 */,
             B_KVFULL = 18 /* + NS: kv_empty */ };

struct Work {
    uint32_t bh, bkv, row0, n0, n1;
};

__device__ __forceinline__ Work decode(const FwdParams& p, uint32_t w) {
    Work t;
    const uint32_t per = p.Hq * p.B;
    const uint32_t qrev = w / per;
    t.bh = w - qrev * per;
    const uint32_t qs = p.causal ? (p.num_q_super - 1 - qrev) : qrev;   // heaviest first under causal
    const uint32_t hq = t.bh % p.Hq, b = t.bh / p.Hq;
    t.bkv = b * p.Hkv + hq / (p.Hq / p.Hkv);
    t.row0 = qs * 256;
    const uint32_t nkb = (p.Sk + 127) / 128;
    t.n0 = p.causal ? min(nkb, t.row0 / 128 + 1) : nkb;    // KV blocks tile 0 needs (diagonal included)
    t.n1 = p.causal ? min(nkb, t.row0 / 128 + 2) : nkb;    // n0 <= n1 always
    return t;
}

// Static persistent schedule: work items are sorted heaviest-first (decode) and dealt to the CTAs in
// boustrophedon ("snake") order, round r going left-to-right when even and right-to-left when odd, which
// balances the monotonically decreasing causal weights to ~0.3% (plain round-robin: 3.4% on config C).
__device__ __forceinline__ bool next_work(const FwdParams& p, uint32_t it, uint32_t& w) {
    const uint32_t base = it * gridDim.x;
    if (base >= p.num_tiles) return false;
    w = base + ((it & 1) ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x);
    return w < p.num_tiles;          // only the last round can be partial
}

struct Ring {
    uint32_t stage = 0, phase = 0;
    template <int NS> __device__ __forceinline__ void advance() {
        if (++stage == NS) { stage = 0; phase ^= 1; }
    }
};

template <int D, bool BF16, int EMU4>
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

    if (threadIdx.x == 0 && (sb & 1023u)) { printf("[aule] dynamic smem not 1024-aligned\n"); __trap(); }

    if (warp == 13 && lane == 0) {
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar(B_QFULL + t), 1);     // TMA tx
            mbar_init(bar(B_QEMPTY + t), 1);    // tcgen05.commit
            mbar_init(bar(B_SFULL + t), 1);     // tcgen05.commit
            mbar_init(bar(B_PFULL + t), 128);   // softmax threads
            mbar_init(bar(B_PFULLB + t), 128);  // softmax threads
            mbar_init(bar(B_OFULL + t), 1);     // tcgen05.commit
            mbar_init(bar(B_OEMPTY + t), 128);  // epilogue threads
            mbar_init(bar(B_STFULL + t), 128);  // softmax threads
            mbar_init(bar(B_STEMPTY + t), 128); // epilogue threads
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
        const uint32_t tS = tmem + lane_addr + (t ? C::COL_S1 : C::COL_S0);
        const uint32_t tO = tmem + lane_addr + (t ? C::COL_O1 : C::COL_O0);
        uint32_t cs = 0, it = 0;
        for (uint32_t w; next_work(p, it, w); ++it) {
            const Work wk = decode(p, w);
            const uint32_t n = t ? wk.n1 : wk.n0;
            const uint32_t trow0 = wk.row0 + t * 128;
            const uint32_t grow = trow0 + r;                        // global query row
            float m_used = -INFINITY, l = 0.f;
            for (uint32_t j = 0; j < n; ++j) {
                mbar_wait(bar(B_SFULL + t), cs & 1); ++cs;
                tc_fence_after();
                uint32_t s[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, s[c]);
                tmem_wait_ld();
                const bool need_mask = (p.causal && j * 128 + 127 > trow0) || ((j + 1) * 128 > p.Sk);
                if (need_mask) {
                    const uint32_t lim = p.causal ? min(grow, p.Sk - 1) : p.Sk - 1;   // last visible key
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (j * 128 + c * 32 + i > lim) s[c][i] = 0xff800000u;        // -inf
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
                if (__any_sync(0xffffffffu, grow_max)) {
                    const float alpha = grow_max ? ex2((m_used - m_new) * p.scale_log2) : 1.f;
                    l *= alpha;
                    if (grow_max) m_used = m_new;
                    if (j > 0) {                                    // O_t(j-1) is complete (s_full tracks all prior MMAs)
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
                const float neg_ms = (m_used == -INFINITY) ? 0.f : -m_used * p.scale_log2;
                // P = exp2(s*scale_log2 - m*scale_log2), two values per instruction (f32x2). EMU4 of every 4
                // pairs take the polynomial path (FMA/ALU pipes) instead of MUFU.EX2, which is the pipe that
                // otherwise paces this kernel (16 ex2/clk/SM against 8192 MMA flop/clk/SM).
                const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(neg_ms, neg_ms);
                float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
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
                        pk[i] = pack2<BF16>(e.x, e.y);
                    }
                    tmem_st16(tS + c * 16, pk);                     // P_t: 32 values -> 16 columns
                    if (c == 2) {                                   // keys 0..95 ready: PV k-steps 0..5 may start
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive(bar(B_PFULL + t));
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(bar(B_PFULLB + t));
                const float2 acc = __fadd2_rn(acc0, acc1);
                l += acc.x + acc.y;
            }
            // hand the row statistics to the epilogue warps
            mbar_wait(bar(B_STEMPTY + t), (it & 1) ^ 1);
            sStat[t * 128 + r] = l;
            sStat[256 + t * 128 + r] = m_used;
            mbar_arrive(bar(B_STFULL + t));
        }
    } else if (warp < 12) {
        // ===================================================== epilogue warps
        reg_dec<64>();
        const uint32_t r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = ((warp & 3) * 32) << 16;
        const bool issuer = (warp == 8 && lane == 0);
        uint32_t it = 0;
        for (uint32_t w; next_work(p, it, w); ++it) {
            const Work wk = decode(p, w);
            for (uint32_t t = 0; t < 2; ++t) {
                const uint32_t tO = tmem + lane_addr + (t ? C::COL_O1 : C::COL_O0);
                const uint32_t sO = sb + C::OFF_O + t * C::TILE_BYTES;
                mbar_wait(bar(B_STFULL + t), it & 1);
                const float l = sStat[t * 128 + r];
                const float m = sStat[256 + t * 128 + r];
                mbar_arrive(bar(B_STEMPTY + t));
                mbar_wait(bar(B_OFULL + t), it & 1);
                tc_fence_after();
                if (issuer) tma_store_wait_read<1>();               // the store that last read sO[t] is done
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
                        tma_store_3d(tmO, sO + c * C::CHUNK_BYTES, c * 64, (int32_t)(wk.row0 + t * 128), (int32_t)wk.bh);
                    tma_store_commit();
                }
                const uint32_t grow = wk.row0 + t * 128 + r;
                if (p.lse != nullptr && grow < p.Sq)
                    p.lse[(size_t)wk.bh * p.Sq + grow] = m * p.scale + __logf(l);   // LSE = m + ln(l)
            }
        }
        if (issuer) tma_store_wait_all<0>();
    } else {
        reg_dec<64>();
        if (warp == 12) {
            // ================================================= MMA issuer
            if (elect_one()) {
                Ring ring;
                uint32_t cp0 = 0, cp1 = 0, it = 0;
                // Descriptors are (constant high word, low word = const | addr>>4); stepping along K
                // is an immediate add on the low word (the 14-bit address field never carries).
                constexpr uint32_t HI_K_HI = uint32_t(HI_K >> 32), HI_K_LO = uint32_t(HI_K);
                constexpr uint32_t HI_V_HI = uint32_t(HI_V >> 32), HI_V_LO = uint32_t(HI_V);
                auto mk = [](uint32_t hi, uint32_t lo) -> uint64_t { return (uint64_t(hi) << 32) | lo; };
                auto issue_qk = [&](uint32_t t, uint32_t kstage) {
                    const uint32_t a_lo = HI_K_LO | ((sb + C::OFF_Q + t * C::TILE_BYTES) >> 4);
                    const uint32_t b_lo = HI_K_LO | ((sb + C::OFF_KV + kstage * C::TILE_BYTES) >> 4);
                    const uint32_t d = tmem + (t ? C::COL_S1 : C::COL_S0);
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                        mma_ss(d, mk(HI_K_HI, a_lo + off), mk(HI_K_HI, b_lo + off), IDESC_QK, kk > 0);
                    }
                };
                auto issue_pv = [&](uint32_t t, uint32_t vstage, bool acc, int k0, int k1) {
                    const uint32_t b_lo = HI_V_LO | ((sb + C::OFF_KV + vstage * C::TILE_BYTES) >> 4);
                    const uint32_t a = tmem + (t ? C::COL_S1 : C::COL_S0);
                    const uint32_t d = tmem + (t ? C::COL_O1 : C::COL_O0);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        if (kk >= k0 && kk < k1)
                            mma_ts(d, a + kk * 8, mk(HI_V_HI, b_lo + kk * (2048 >> 4)), IDESC_PV, (acc || kk > 0) ? 1u : 0u);
                };
                for (uint32_t w; next_work(p, it, w); ++it) {
                    const Work wk = decode(p, w);
                    const uint32_t nmax = wk.n1;                    // n0 <= n1
                    // ---- prologue: S_t(0) = Q_t K_0^T
                    uint32_t kstage = ring.stage;
                    mbar_wait(bar(B_KVFULL + kstage), ring.phase);
                    ring.advance<NS>();
                    for (uint32_t t = 0; t < 2; ++t) {
                        mbar_wait(bar(B_QFULL + t), it & 1);
                        tc_fence_after();
                        issue_qk(t, kstage);
                        mma_commit(bar(B_SFULL + t));
                        if ((t ? wk.n1 : wk.n0) == 1) mma_commit(bar(B_QEMPTY + t));
                    }
                    mma_commit(bar(B_KVFULL + NS + kstage));
                    // ---- main loop
                    for (uint32_t j = 0; j < nmax; ++j) {
                        const uint32_t vstage = ring.stage;
                        mbar_wait(bar(B_KVFULL + vstage), ring.phase);
                        ring.advance<NS>();
                        const bool has_k = (j + 1 < nmax);
                        const uint32_t kst = ring.stage, kph = ring.phase;
                        bool k_ready = false;
                        if (has_k) ring.advance<NS>();
                        for (uint32_t t = 0; t < 2; ++t) {
                            const uint32_t nt = t ? wk.n1 : wk.n0;
                            if (j >= nt) continue;
                            uint32_t& cp = t ? cp1 : cp0;
                            mbar_wait(bar(B_PFULL + t), cp & 1);
                            if (j == 0) mbar_wait(bar(B_OEMPTY + t), (it & 1) ^ 1);
                            tc_fence_after();
                            issue_pv(t, vstage, j > 0, 0, 6);
                            mbar_wait(bar(B_PFULLB + t), cp & 1); ++cp;
                            tc_fence_after();
                            issue_pv(t, vstage, j > 0, 6, 8);
                            if (j == nt - 1) mma_commit(bar(B_OFULL + t));
                            if (t == 1) mma_commit(bar(B_KVFULL + NS + vstage));     // V_j: tile 1 is the last user
                            if (j + 1 < nt) {
                                if (!k_ready) { mbar_wait(bar(B_KVFULL + kst), kph); tc_fence_after(); k_ready = true; }
                                issue_qk(t, kst);
                                mma_commit(bar(B_SFULL + t));
                                if (j + 1 == nt - 1) mma_commit(bar(B_QEMPTY + t));
                                if (t == 1) mma_commit(bar(B_KVFULL + NS + kst));    // K_{j+1}: same
                            }
                        }
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
                for (uint32_t w; next_work(p, it, w); ++it) {
                    const Work wk = decode(p, w);
                    load_kv(tmK, 0, wk.bkv);
                    for (uint32_t t = 0; t < 2; ++t) {
                        mbar_wait(bar(B_QEMPTY + t), (it & 1) ^ 1);
                        const uint32_t full = bar(B_QFULL + t);
                        mbar_expect_tx(full, C::TILE_BYTES);
#pragma unroll
                        for (int c = 0; c < C::CHUNKS; ++c)
                            tma_load_3d(sb + C::OFF_Q + t * C::TILE_BYTES + c * C::CHUNK_BYTES, tmQ, full, c * 64,
                                        (int32_t)(wk.row0 + t * 128), (int32_t)wk.bh);
                    }
                    for (uint32_t j = 0; j < wk.n1; ++j) {
                        load_kv(tmV, j, wk.bkv);
                        if (j + 1 < wk.n1) load_kv(tmK, j + 1, wk.bkv);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 14) tmem_dealloc<512>(tmem);
}

}  // namespace fwd100v3

#undef AULE_FWD100
#define AULE_FWD100(NAME, DD, BF, EMU)                                                                  \
    extern "C" __global__ void __launch_bounds__(512, 1) NAME(const __grid_constant__ CUtensorMap tmQ,   \
                                                              const __grid_constant__ CUtensorMap tmK,   \
                                                              const __grid_constant__ CUtensorMap tmV,   \
                                                              const __grid_constant__ CUtensorMap tmO,   \
                                                              const aule_kp::FwdParams p) {              \
        fwd100v3::fwd_body<DD, BF, EMU>(&tmQ, &tmK, &tmV, &tmO, p);                                        \
    }

// v3 of the forward kernel (S/P aliased, per-tile S buffers), kept as an A/B baseline for v4.
AULE_FWD100(aule_fwd_sm100_bf16_d128_v3, 128, true, 0)
