// Microbenchmark: per-SM throughput of MUFU.EX2, FFMA2, FADD2, F2FP and the polynomial exp2
// with 1, 2, 4 warps per SMSP (148 CTAs). Build: nvcc -arch=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = -0.001f * (threadIdx.x + i);
    float2 acc = make_float2(0.f, 0.f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // 16 MUFU
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = ex2(a[i]);
        } else if (MODE == 1) {     // softmax-like: FFMA2 + 2 MUFU + FADD2 + F2FP per pair
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 x = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(0.999f, 0.999f), make_float2(-0.01f, -0.01f));
                float2 e = make_float2(ex2(x.x), ex2(x.y));
                acc = __fadd2_rn(acc, e);
                unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e.y), "f"(e.x));
                a[i] = x.x; a[i + 1] = __uint_as_float(r) * 1e-30f + x.y;
            }
        } else if (MODE == 2) {     // 16 FFMA2 (32 flops-pairs)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 x = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(0.999f, 0.999f), make_float2(-0.01f, -0.01f));
                a[i] = x.x; a[i + 1] = x.y;
            }
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 x = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(1.001f, 1.001f), make_float2(0.01f, 0.01f));
                a[i] = x.x; a[i + 1] = x.y;
            }
        } else if (MODE == 3) {     // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 0.999f, -0.01f);
        } else if (MODE == 4) {     // 16 F2FP
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned r0, r1;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r0) : "f"(a[i]), "f"(a[i + 1]));
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r1) : "f"(a[i + 1]), "f"(a[i]));
                a[i] = __uint_as_float(r0 & 0x3fffffffu); a[i + 1] = __uint_as_float(r1 & 0x3fffffffu);
            }
        } else if (MODE == 5) {     // 16 FMNMX3-ish
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaxf(fmaxf(a[i], a[(i + 1) & 15]), a[(i + 2) & 15] - 1.f);
        }
    }
    long long t1 = clock64();
    float s = acc.x + acc.y;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int opsPerIter) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 2000;
        k<MODE><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double perWarpInstr = (double)c / ((double)iters * opsPerIter);                 // cycles per warp-instruction, this warp
        double perSmsp = perWarpInstr / (warps / 4.0);                                  // cycles per warp-instruction per SMSP
        printf("%-28s warps/SM=%2d  cycles/iter=%8.1f  cyc per op (one warp)=%6.2f  cyc per op per SMSP=%6.2f\n", name, warps,
               (double)c / iters, perWarpInstr, perSmsp);
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("MUFU.EX2 x16", 16);
    run<1>("softmax pair loop (8 pairs)", 16);
    run<2>("FFMA2 x16", 16);
    run<3>("FFMA x16", 16);
    run<4>("F2FP x16", 16);
    run<5>("FMNMX x32", 32);
    return 0;
}
