// Microbenchmark (round 2): issue rate of the forward kernel's two MMA shapes on one SM, 148 CTAs.
//   SS: D[tmem] = A[smem] * B[smem]   (Q K^T: Q and K tiles both read from shared memory, M=128 N=128 K=16 per instruction)
//   TS: D[tmem] = A[tmem] * B[smem]   (P V  : P from tensor memory, V from shared memory)
// One thread issues groups of 8 instructions (one 128x128x128 bf16 product = 512 cycles at 8192 FLOP/clk/SM) with two
// groups in flight; optionally other warps hammer shared memory with stores / loads to show how much an SS MMA's operand
// fetch (8 KB per instruction = 128 B/clk, the whole shared-memory bandwidth) suffers from competing traffic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I aule-attention_b200/csrc/kernels -o tools/microbench/mma_rate tools/microbench/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"

using namespace sm100;

// mode: 0 SS, 1 TS, 2 alternating SS / TS (the kernel's steady state)
// noise: 0 none, 1 four warps storing to shared memory (32 B/clk-ish), 2 eight warps loading+storing
__global__ void __launch_bounds__(512, 1) k(long long* cyc, int mode, int noise, int groups) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[4];
    __shared__ uint32_t tslot;
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(&tslot));
    // zero the operand tiles (finite values)
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    constexpr uint64_t HI_K = smem_desc_hi(16, 1024), HI_V = smem_desc_hi(16384, 1024);
    constexpr uint32_t ID_QK = instr_desc_f16(true, 128, 128, false), ID_PV = instr_desc_f16(true, 128, 128, true);
    volatile __shared__ int stop;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t sQ = sb, sK = sb + 32768, sV = sb + 65536;
            auto mk = [](uint64_t hi, uint32_t addr) { return hi | uint64_t((addr >> 4) & 0x3FFF); };
            long long t0 = clock64();
            for (int g = 0; g < groups; ++g) {
                const int b = g & 1;
                if (g >= 2) mbar_wait(smem_u32(&bars[b]), ((g - 2) >> 1) & 1);
                const bool ss = mode == 0 || (mode == 2 && !(g & 1));
                if (ss) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t off = (kk / 4) * 16384 + (kk % 4) * 32;
                        mma_ss(tmem + 128 * b, mk(HI_K, sQ + off), mk(HI_K, sK + off), ID_QK, kk > 0);
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        mma_ts(tmem + 256 + 128 * b, tmem + 8 * kk, mk(HI_V, sV + kk * 2048), ID_PV, kk > 0);
                }
                mma_commit(smem_u32(&bars[b]));
            }
            for (int g = groups - 2; g < groups; ++g) mbar_wait(smem_u32(&bars[g & 1]), (g >> 1) & 1);
            if (blockIdx.x == 0) cyc[0] = clock64() - t0;
            stop = 1;
        }
    } else if (noise == 3 && warp >= 4 && warp < 12) {
        // competing TENSOR-MEMORY traffic: what the softmax warps do (load 128 fp32 columns, store 64) on columns the MMAs do not use
        const uint32_t ta = tmem + (((warp & 3) * 32) << 16) + 384;
        uint32_t r[32], acc = 0;
        while (!stop) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { tmem_ld32(ta + 32 * (c & 3), r); tmem_wait_ld(); acc ^= r[0]; }
            tmem_st16(ta, r); tmem_st16(ta + 16, r + 16); tmem_st16(ta + 32, r); tmem_st16(ta + 48, r + 16);
            tmem_wait_st();
        }
        if (acc == 0xdeadbeef) cyc[1] = acc;
    } else if (noise && noise < 3 && warp >= 4 && warp < (noise == 1 ? 8 : 12)) {
        // competing shared-memory traffic in a separate 64 KB region
        uint4* base = reinterpret_cast<uint4*>(smem + 98304) + threadIdx.x;
        uint4 v = make_uint4(lane, 1, 2, 3);
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (noise == 2) { uint4 w = base[(i * 512) & 3071]; v.x ^= w.y; }
                base[(i * 512 + 256) & 3071] = v;
            }
        }
        if (v.x == 0xdeadbeef) cyc[1] = v.x;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

int main() {
    long long* cyc;
    cudaMalloc(&cyc, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int groups = 2000;
    const char* mn[] = {"SS (Q K^T form)", "TS (P V form)", "SS / TS alternating"};
    const char* nn[] = {"no other smem traffic", "4 warps storing to smem", "8 warps loading+storing smem", "8 warps tcgen05.ld/st (TMEM)"};
    for (int noise = 0; noise < 4; ++noise)
        for (int mode = 0; mode < 3; ++mode) {
            k<<<148, 512, 200 * 1024>>>(cyc, mode, noise, groups);
            cudaError_t e = cudaDeviceSynchronize();
            long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-22s | %-30s: %7.1f cycles per 8-instruction group (512 = peak)  %s\n", mn[mode], nn[noise], (double)h / groups,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
