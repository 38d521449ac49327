// Microbenchmark (round 2): issue cost per warp instruction and per SM sub-partition of the instructions the softmax
// warps of attn_fwd_sm100.cu are made of, including the packed 16-bit exp2 forms and TMEM loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu ; run on a B200.
// Every mode runs 16 independent dependency chains per thread (ILP 16) with asm volatile bodies.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define R16(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15)

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters, long long* cyc) {
    float a[16];
    uint32_t u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = -0.001f * (threadIdx.x + i) - 0.5f; u[i] = 0x3c003c00u + threadIdx.x + i; }
    __shared__ uint32_t tslot;
    uint32_t tmem = 0;
    if (MODE == 20 || MODE == 21) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tslot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem = tslot + ((((threadIdx.x >> 5) & 3) * 32) << 16);
    }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // MUFU.EX2 f32
#define X(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            R16(X)
#undef X
        } else if (MODE == 1) {     // ex2 bf16x2
#define X(i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
            R16(X)
#undef X
        } else if (MODE == 2) {     // ex2 f16x2
#define X(i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
            R16(X)
#undef X
        } else if (MODE == 3) {     // F2FP bf16x2 <- 2 x f32
#define X(i) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
            R16(X)
#undef X
        } else if (MODE == 4) {     // F2FP f16x2
#define X(i) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
            R16(X)
#undef X
        } else if (MODE == 5) {     // 8 MUFU f32 + 8 F2FP interleaved (shared pipe?)
#define X(i) if ((i) & 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); else asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 2) & 15]));
            R16(X)
#undef X
        } else if (MODE == 6) {     // FFMA2
#define X(i) if (!((i) & 1)) asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2}; mov.b64 z, {%3, %3}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(a[(i) + 1]) : "f"(0.999f), "f"(-0.01f));
            R16(X) R16(X)
#undef X
        } else if (MODE == 7) {     // FADD2
#define X(i) if (!((i) & 1)) asm volatile("{.reg .b64 x, z; mov.b64 x, {%0, %1}; mov.b64 z, {%2, %2}; add.rn.f32x2 x, x, z; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(a[(i) + 1]) : "f"(-0.01f));
            R16(X) R16(X)
#undef X
        } else if (MODE == 8) {     // FMNMX3
#define X(i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 5) & 15]), "f"(a[(i + 9) & 15]));
            R16(X)
#undef X
        } else if (MODE == 9) {     // PRMT
#define X(i) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 3) & 15]));
            R16(X)
#undef X
        } else if (MODE == 10) {    // FFMA scalar
#define X(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(-0.01f));
            R16(X)
#undef X
        } else if (MODE == 11) {    // shl + iadd (exponent patch of the polynomial exp2)
#define X(i) asm volatile("{.reg .b32 t; shl.b32 t, %1, 23; add.s32 %0, %0, t;}" : "+r"(u[i]) : "r"(u[(i + 3) & 15]));
            R16(X)
#undef X
        } else if (MODE == 12) {    // current MUFU pair: FFMA2 + 2 MUFU + FADD2 + PRMT        (8 pairs)
#define X(i) if (!((i) & 1)) { float e0, e1; asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %4}; mov.b64 z, {%5, %5}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}" : "=f"(e0), "=f"(e1) : "f"(a[i]), "f"(a[(i) + 1]), "f"(0.999f), "f"(-0.01f)); \
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e1)); \
            asm volatile("{.reg .b64 x, z; mov.b64 x, {%0, %1}; mov.b64 z, {%2, %3}; add.rn.f32x2 x, x, z; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(a[(i) + 1]) : "f"(e0), "f"(e1)); \
            asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(__float_as_uint(e0)), "r"(__float_as_uint(e1))); }
            R16(X)
#undef X
        } else if (MODE == 13) {    // f16x2 pair: FFMA2 + F2FP(f16x2) + 1 MUFU f16x2 + 2 cvt f32<-f16 + FADD2   (8 pairs)
#define X(i) if (!((i) & 1)) { float e0, e1; uint32_t h; asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %4}; mov.b64 z, {%5, %5}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}" : "=f"(e0), "=f"(e1) : "f"(a[i]), "f"(a[(i) + 1]), "f"(0.999f), "f"(-0.01f)); \
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(e1), "f"(e0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h)); \
            asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=f"(e0), "=f"(e1) : "r"(h)); \
            asm volatile("{.reg .b64 x, z; mov.b64 x, {%0, %1}; mov.b64 z, {%2, %3}; add.rn.f32x2 x, x, z; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(a[(i) + 1]) : "f"(e0), "f"(e1)); u[i] = h; }
            R16(X)
#undef X
        } else if (MODE == 14) {    // bf16x2 pair: FFMA2 + F2FP(bf16x2) + 1 MUFU bf16x2 + unpack by shifts + FADD2   (8 pairs)
#define X(i) if (!((i) & 1)) { float e0, e1; uint32_t h; asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %4}; mov.b64 z, {%5, %5}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}" : "=f"(e0), "=f"(e1) : "f"(a[i]), "f"(a[(i) + 1]), "f"(0.999f), "f"(-0.01f)); \
            asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(e1), "f"(e0)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h)); \
            e0 = __uint_as_float(h << 16); e1 = __uint_as_float(h & 0xffff0000u); \
            asm volatile("{.reg .b64 x, z; mov.b64 x, {%0, %1}; mov.b64 z, {%2, %3}; add.rn.f32x2 x, x, z; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(a[(i) + 1]) : "f"(e0), "f"(e1)); u[i] = h; }
            R16(X)
#undef X
        } else if (MODE == 20) {    // tcgen05.ld 32x32b.x32 (4 per iteration = one 128-column row block) + wait
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(tmem + c * 32) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                u[c] += r[0] ^ r[31];
            }
        } else if (MODE == 21) {    // same, one wait for all four loads
            uint32_t r[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[c][0]), "=r"(r[c][1]), "=r"(r[c][2]), "=r"(r[c][3]), "=r"(r[c][4]), "=r"(r[c][5]), "=r"(r[c][6]), "=r"(r[c][7]), "=r"(r[c][8]), "=r"(r[c][9]), "=r"(r[c][10]), "=r"(r[c][11]), "=r"(r[c][12]), "=r"(r[c][13]), "=r"(r[c][14]), "=r"(r[c][15]),
                               "=r"(r[c][16]), "=r"(r[c][17]), "=r"(r[c][18]), "=r"(r[c][19]), "=r"(r[c][20]), "=r"(r[c][21]), "=r"(r[c][22]), "=r"(r[c][23]), "=r"(r[c][24]), "=r"(r[c][25]), "=r"(r[c][26]), "=r"(r[c][27]), "=r"(r[c][28]), "=r"(r[c][29]), "=r"(r[c][30]), "=r"(r[c][31])
                             : "r"(tmem + c * 32) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 4; ++c) u[c] += r[c][0] ^ r[c][31];
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float(u[i] & 0x3fffffffu);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (MODE == 20 || MODE == 21) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tslot) : "memory");
    }
}

template <int MODE>
void run(const char* name, int opsPerIter) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int warps : {4, 8, 16}) {
        const int iters = 2000;
        k<MODE><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, iters, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double perWarpInstr = (double)c / ((double)iters * opsPerIter);
        double perSmsp = perWarpInstr / (warps / 4.0);
        printf("%-44s warps/SMSP=%d  cyc/iter=%8.1f  cyc per op (one warp)=%7.2f  cyc per op per SMSP=%7.2f %s\n", name, warps / 4,
               (double)c / iters, perWarpInstr, perSmsp, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("MUFU.EX2 f32 x16", 16);
    run<1>("ex2.bf16x2 x16 (32 exps)", 16);
    run<2>("ex2.f16x2 x16 (32 exps)", 16);
    run<3>("F2FP bf16x2 x16", 16);
    run<4>("F2FP f16x2 x16", 16);
    run<5>("8 MUFU f32 + 8 F2FP interleaved", 16);
    run<6>("FFMA2 x16", 16);
    run<7>("FADD2 x16", 16);
    run<8>("FMNMX3 x16", 16);
    run<9>("PRMT x16", 16);
    run<10>("FFMA x16", 16);
    run<11>("SHL+IADD x16 (32 instr)", 16);
    run<12>("pair: FFMA2+2MUFU+FADD2+PRMT (8 pairs)", 8);
    run<13>("pair: FFMA2+F2FP+MUFU.f16x2+2cvt+FADD2 (8 pairs)", 8);
    run<14>("pair: FFMA2+F2FP+MUFU.bf16x2+unpack+FADD2 (8)", 8);
    run<20>("tcgen05.ld x32 + wait, x4", 4);
    run<21>("tcgen05.ld x32 x4, one wait", 4);
    return 0;
}
