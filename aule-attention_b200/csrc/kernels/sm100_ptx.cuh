// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (MMA / TMEM alloc / ld / st / commit / fences), setmaxnreg.
// Hand-written for this engine; bit layouts follow the PTX ISA 8.7 tcgen05
// chapter (shared-memory matrix descriptor, instruction descriptor).
#pragma once
#include <cuda.h>
#include <stdint.h>

#ifndef AULE_WATCHDOG
#define AULE_WATCHDOG 0          // 1 (bring-up): bound every mbarrier spin, print and trap instead of hanging the GPU
#endif

namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// HINT_NS: suspend-time hint of mbarrier.try_wait (the thread may sleep up to that long before re-checking;
// a completing phase wakes it).  0 = plain try_wait (hardware default time limit).
template <uint32_t HINT_NS = 1000000u>
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    if constexpr (HINT_NS == 0) {
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(HINT_NS)
            : "memory");
    }
    return ok != 0;
}
// Wait until the phase with the given parity has completed.
// Non-blocking poll (mbarrier.test_wait never suspends the thread): for "is this stage free yet?" checks made on the side of
// another wait -- a try_wait on a barrier that is not complete may sleep for the hardware's time limit.
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
template <uint32_t HINT_NS = 1000000u>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if AULE_WATCHDOG
    if (mbar_try_wait<HINT_NS>(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait<HINT_NS>(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {      // ~2 s: a pipeline bug, not a slow kernel
            printf("[aule] mbarrier watchdog: block %d thread %d bar 0x%x parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
#else
    while (!mbar_try_wait<HINT_NS>(bar, parity)) {}
#endif
}

// ------------------------------------------------------------------ proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// global -> shared, 3-D tile, completes `bytes` on `bar`.
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar,
                                            int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// global -> L2 only (no shared-memory destination): warms the tile a later tma_load_3d will fetch.
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// shared -> global, 3-D tile (out-of-bounds rows are clipped by the hardware).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src_smem,
                                             int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem_slot) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem_slot), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {          // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives lane (base_lane + i),
// columns [col, col+32).  `taddr` = (lane << 16) | column.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
        : "memory");
}

// ------------------------------------------------------------------ tcgen05.mma
// Shared-memory matrix descriptor (64-bit):
//   [ 0,14) start address >> 4      [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4   [46,48) version = 1 (sm_100)
//   [49,52) base offset = 0         [52] LBO mode = 0      [61,64) swizzle: 2 = 128B
// K-major, 128B swizzle  : rows of 64 x 16-bit (128 B); 8-row groups are 1024-B atoms,
//                          SBO = 1024 (next 8 rows); LBO unused.
// MN-major, 128B swizzle : 64 MN elements (128 B) contiguous, 8 K-rows per 1024-B atom,
//                          SBO = 1024 (next 8 K rows), LBO = stride between 64-wide MN chunks.
__host__ __device__ constexpr uint64_t smem_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) | (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ uint64_t smem_desc(uint64_t hi_bits, uint32_t smem_addr) {
    return hi_bits | uint64_t((smem_addr >> 4) & 0x3FFF);
}
// Instruction descriptor for kind::f16 (fp32 accumulate):
//   [4,6) D fmt: 1 = f32   [7,10) A fmt   [10,13) B fmt  (0 = f16, 1 = bf16)
//   [15] A major (0 = K)   [16] B major (0 = K, 1 = MN)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t instr_desc_f16(bool bf16, uint32_t M, uint32_t N, bool b_mn_major) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((b_mn_major ? 1u : 0u) << 16) |
           ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]      (one thread issues)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::tf32 (fp32 operands read as tf32, fp32 accumulate): instruction descriptor with A / B format 2 = TF32; K = 8 per instruction.
__host__ __device__ constexpr uint32_t instr_desc_tf32(uint32_t M, uint32_t N, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive when all previously issued tcgen05 async ops of this thread are complete.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ------------------------------------------------------------------ registers
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ------------------------------------------------------------------ math
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp2 for a pair on the FMA/ALU pipes (no MUFU): Cody-Waite split x = floor(x) + f, degree-3
// polynomial for 2^f on [0,1) (max rel. error ~1e-4, far below the bf16 rounding of P), exponent
// patched in with an integer add.  Valid for x <= 127; x < -127 is clamped (result ~1e-38 ~ 0).
__device__ __forceinline__ float2 ex2_emu2(float2 x) {
    const float magic = 12582912.f;                       // 2^23 + 2^22
    x.x = fmaxf(x.x, -127.f);
    x.y = fmaxf(x.y, -127.f);
    float2 xr;                                            // x + magic, rounded DOWN: floor(x) in the low mantissa bits
    asm("{\n\t.reg .b64 a, b, d;\n\t"
        "mov.b64 a, {%2, %3};\n\t"
        "mov.b64 b, {%4, %4};\n\t"
        "add.rm.ftz.f32x2 d, a, b;\n\t"
        "mov.b64 {%0, %1}, d;\n\t}"
        : "=f"(xr.x), "=f"(xr.y)
        : "f"(x.x), "f"(x.y), "f"(magic));
    const float2 fl = __fadd2_rn(xr, make_float2(-magic, -magic));            // floor(x) as float
    const float2 f = __ffma2_rn(fl, make_float2(-1.f, -1.f), x);               // fractional part in [0,1)
    float2 p = __ffma2_rn(f, make_float2(0.077119089663028717041015625f, 0.077119089663028717041015625f),
                          make_float2(0.227564394474029541015625f, 0.227564394474029541015625f));
    p = __ffma2_rn(p, f, make_float2(0.695146143436431884765625f, 0.695146143436431884765625f));
    p = __ffma2_rn(p, f, make_float2(1.f, 1.f));
    float2 r;
    r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(xr.x) << 23));
    r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(xr.y) << 23));
    return r;
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    if constexpr (BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

}  // namespace sm100
