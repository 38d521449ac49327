// Fused FlashAttention-2 forward for fp32 inputs on the sm_100a tensor cores (kind::tf32).                (round 2)
//
// The reference's GPU path feeds fp32 tensors to tl.dot, which runs them as tf32 (python/aule/triton_flash.py:405-411, Triton's
// default allow_tf32); its Vulkan shaders are exact fp32 (attention_f32.comp).  Here fp32 inputs run the exact CUDA-core
// kernel of attn_simt.cu unless the caller asks for tf32 (dtype code 3 at the C ABI; aule.flash_attention(allow_tf32=True)).
//
// Same kernel as attn_fwd_sm100.cu (v5 slot stream: roles, barriers, work scheduler, mask rules are shared code) with
// fp32 operands.  An fp32 [128 rows][64] tile has the bytes of a bf16 [128][128] one, so the shared-memory geometry, the TMA
// boxes (32 fp32 = 128 bytes, 128B swizzle) and the descriptor arithmetic are those of the D = 128 bf16 kernel:
//   Q K^T : 8 instructions of K = 8 (32 bytes per row, the same 32-byte k-step as 16 bf16)
//   P     : stays fp32 in tensor memory, 128 columns per tile (S [0,128) | P0 [128,256) | P1 [256,384) | O0 [384,448) | O1 [448,512))
//   P V   : 16 instructions of K = 8 keys; V is the MN-major B operand (8 rows of 128 bytes per k-step)
// so head_dim <= 64 only (D = 128 would need 2 x 128 more TMEM columns).  Inputs are truncated to tf32 by the MMA; P is
// truncated by the softmax warps before it is summed, so l normalises exactly the weights the MMA uses.
#include "sm100_ptx.cuh"
#include "kernel_params.h"

namespace fwd100 {

template <int EMUN, int EMUD, bool TRACE>
__device__ __forceinline__ void fwd_tf32_body(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                              const FwdParams& p) {
    using C = aule_kp::FwdCfgT32;
    constexpr int D = C::D;                                             // logical (padded) head_dim: 64 fp32 = two 128-byte chunks
    constexpr int VAR = 1 + 131072;                                     // the shipped softmax loop of attn_fwd_sm100.cu
    constexpr bool BF16 = false, TRUNC_PACK = false;
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
    // V is the MN-major B operand.  For 32-bit operands the only MN-major shared-memory layout the tensor core reads is the
    // 128-byte swizzle with a 32-byte atom (descriptor layout type 1; CUTLASS Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over
    // [4 rows][128 bytes]): the V tensor map is encoded with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, a k-step of 8 keys is two
    // 4-row atoms 512 bytes apart (SBO), the two 32-column halves of D are one chunk apart (LBO).
    constexpr uint64_t HI_V = (smem_desc_hi(C::CHUNK_BYTES, 512) & ~(uint64_t(7) << 61)) | (uint64_t(1) << 61);
    constexpr uint32_t IDESC_QK = instr_desc_tf32(128, 128, false);
    constexpr uint32_t IDESC_PV = instr_desc_tf32(128, C::PV_N, true);

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
                auto wait_pv = [&]() {                              // P_t is still being read by PV_t of the previous block until pv_done
                    tr.ev(14, g);
                    if (g > 0 && !pv_waited) WAIT(bar(B_PVDONE + t), (g - 1) & 1);
                    tr.ev(15, g);
                    tc_fence_after();
                };
                // P stays fp32 in TMEM (the tf32 MMA reads the upper 19 bits): the exponentials are truncated to tf32 HERE, before
                // they are summed, so that the row sum l normalises exactly the weights the MMA uses.
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float2 x = make_float2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1]));
                        float2 e;
                        if ((i % EMUD) < EMUN) {
                            e = ex2_emu2(x);
                        } else {
                            e.x = ex2(x.x);
                            e.y = ex2(x.y);
                        }
                        e.x = __uint_as_float(__float_as_uint(e.x) & 0xFFFFE000u);
                        e.y = __uint_as_float(__float_as_uint(e.y) & 0xFFFFE000u);
                        if (i & 1) acc1 = __fadd2_rn(acc1, e); else acc0 = __fadd2_rn(acc0, e);
                        s[c][2 * i] = __float_as_uint(e.x); s[c][2 * i + 1] = __float_as_uint(e.y);
                    }
                    if (c == 2) {
                        wait_pv();
                        tmem_st32(tP, s[0]);
                        tmem_st32(tP + 32, s[1]);
                        tmem_st32(tP + 64, s[2]);                   // keys 0..95 ready: PV k-steps 0..11 may start
                        tmem_wait_st();
                        tc_fence_before();
                        ARRIVE(bar(B_PFULL + t));
                        tr.ev(16, g);
                    } else if (c == 3) {
                        tmem_st32(tP + 96, s[3]);
                        tmem_wait_st();
                        tc_fence_before();
                        ARRIVE(bar(B_PFULLB + t));
                        tr.ev(17, g);
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
                uint8_t* optr = reinterpret_cast<uint8_t*>(p.o) + orow * (size_t)(p.D_real * 4);
                mbar_wait(bar(B_OFULL + t), it & 1);
                tc_fence_after();
                // The MMA TRUNCATES its fp32 operands to tf32 (13 mantissa bits dropped: a relative shrink of 2^-11 / mantissa, 3.6e-4
                // on average for non-degenerate data).  Q K^T carries that bias twice -- the host folds (1 + 7.2e-4) into the softmax
                // scale it passes to this kernel -- and V once, undone here; P is truncated before it is summed, so it cancels in
                // O = (P V) / l.  Without the compensation the output error is 2e-3 of its scale (measured), with it 3e-4.
                const float inv = (l > 0.f) ? 1.00036f / l : 0.f;   // rows without a visible key: O = 0, LSE = -inf
#pragma unroll
                for (int c = 0; c < D / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld32(tO + c * 32, o);
                    tmem_wait_ld();
                    if (c == D / 32 - 1) {                          // O_t fully read: MMA may overwrite it
                        tc_fence_before();
                        ARRIVE(bar(B_OEMPTY + t));
                    }
#pragma unroll
                    for (int u = 0; u < 32; ++u) o[u] = __float_as_uint(__uint_as_float(o[u]) * inv);
                    if (row_ok) {
                        if (p.D_real == (uint32_t)D) {              // 32 fp32 = 128 B of this row: four 256-bit stores (whole 32-byte sectors)
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(optr + c * 128 + u * 32),
                                             "r"(o[8 * u]), "r"(o[8 * u + 1]), "r"(o[8 * u + 2]), "r"(o[8 * u + 3]),
                                             "r"(o[8 * u + 4]), "r"(o[8 * u + 5]), "r"(o[8 * u + 6]), "r"(o[8 * u + 7]) : "memory");
                        } else {                                    // padded head_dim (D_real % 4 == 0): 128-bit pieces inside the row
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if ((uint32_t)(c * 32 + u * 4 + 4) <= p.D_real)
                                    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(optr + c * 128 + u * 16),
                                                 "r"(o[4 * u]), "r"(o[4 * u + 1]), "r"(o[4 * u + 2]), "r"(o[4 * u + 3]) : "memory");
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
                    for (int kk = 0; kk < D / 8; ++kk) {                 // K = 8 fp32 (32 bytes) per instruction
                        const uint32_t off = ((kk / 4) * C::CHUNK_BYTES + (kk % 4) * 32) >> 4;
                        mma_ss_tf32(d, mk(HI_K_HI, a_lo + off), mk(HI_K_HI, b_lo + off), IDESC_QK, kk > 0);
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
                    for (int kk = 0; kk < 12; ++kk)                      // K = 8 keys per instruction: 8 fp32 P columns, 8 V rows (1 KB)
                        mma_ts_tf32(d, a + kk * 8, mk(HI_V_HI, b_lo + kk * (1024 >> 4)), IDESC_PV, (!first || kk > 0) ? 1u : 0u);
                    mbar_wait<HOT_HINT>(bar(B_PFULLB + t), gpv & 1);
                    tr.ev(3 + 16 * t, gpv);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 12; kk < 16; ++kk)
                        mma_ts_tf32(d, a + kk * 8, mk(HI_V_HI, b_lo + kk * (1024 >> 4)), IDESC_PV, 1u);
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
                        tma_load_3d(dst + c * C::CHUNK_BYTES, map, full, c * 32, (int32_t)(j * 128), (int32_t)bkv);
                    ring.advance<NS>();
                };
                auto load_q = [&](const Cur& c, uint32_t t) {
                    const uint32_t qi = 2 * c.it + t, qslot = qi % 3;
                    mbar_wait(bar(B_QEMPTY + qslot), ((qi / 3) & 1) ^ 1);
                    const uint32_t full = bar(B_QFULL + qslot);
                    mbar_expect_tx(full, C::TILE_BYTES);
#pragma unroll
                    for (int ch = 0; ch < C::CHUNKS; ++ch)
                        tma_load_3d(sb + C::OFF_Q + qslot * C::TILE_BYTES + ch * C::CHUNK_BYTES, tmQ, full, ch * 32,
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

extern "C" __global__ void __launch_bounds__(512, 1) aule_fwd_sm100_tf32_d64(const __grid_constant__ CUtensorMap tmQ,
                                                                               const __grid_constant__ CUtensorMap tmK,
                                                                               const __grid_constant__ CUtensorMap tmV,
                                                                               const aule_kp::FwdParams p) {
    fwd100::fwd_tf32_body<1, 4, false>(&tmQ, &tmK, &tmV, p);
}
