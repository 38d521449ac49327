// Microbenchmark (round 2): how fast can one SM / the whole GPU add fp32 partial tiles into global memory?
// The fused backward (attn_bwd_fused_sm100.cu) drains dQ partials with coalesced scalar `red.global.add.f32` and was
// measured at 5.7-7.4 cycles per 128-byte warp reduction (~3.7 TB/s over the GPU).  This compares the alternatives:
//   mode 0  red.global.add.f32          one 4-byte element per lane  (128 B per warp instruction)   -- what the kernel does
//   mode 1  red.global.add.v2.f32       8 bytes per lane             (256 B per warp instruction)
//   mode 2  red.global.add.v4.f32       16 bytes per lane            (512 B per warp instruction)
//   mode 3  cp.reduce.async.bulk.global.shared::cta.add.f32 of CHUNK bytes from shared memory (one thread issues; the TMA
//           engine does the rest), CHUNK = 2/8/16/32 KB, up to 4 groups in flight
// Every CTA adds `tile` bytes per step (64 KB = one [128 q][128 d] fp32 dQ partial) to a destination tile chosen so that
//   share = 1: every CTA owns its tiles (no two CTAs touch an address),
//   share = 8: 8 CTAs walk the same tiles at the same time (the backward: CTAs of different key blocks add into the same
//              query rows).
// Destination footprint: 32 MB (L2-resident) or 1 GB.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_rate red_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int TILE = 64 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* dst, size_t ntiles, int steps, int share, int chunk, long long* cyc) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* sf = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < TILE / 4; i += blockDim.x) sf[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t group = blockIdx.x / share;
    const size_t ngroups = (gridDim.x + share - 1) / share;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        const size_t tile = (group + (size_t)s * ngroups) % ntiles;
        float* base = dst + tile * (TILE / 4);
        if (MODE == 0) {
            // 16 warps x 32 lanes: warp w covers rows w, w+16, ... of 128 floats?  keep it simple: element index = it*512 + tid
#pragma unroll 8
            for (int it = 0; it < TILE / 4 / 512; ++it)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(base + it * 512 + threadIdx.x), "f"(1.0f) : "memory");
        } else if (MODE == 1) {
#pragma unroll 8
            for (int it = 0; it < TILE / 8 / 512; ++it)
                asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(base + 2 * (it * 512 + threadIdx.x)), "f"(1.0f) : "memory");
        } else if (MODE == 2) {
#pragma unroll 8
            for (int it = 0; it < TILE / 16 / 512; ++it)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(base + 4 * (it * 512 + threadIdx.x)), "f"(1.0f) : "memory");
        } else {
            if (threadIdx.x == 0) {
                for (int off = 0; off < TILE; off += chunk) {
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                                     reinterpret_cast<uint8_t*>(base) + off),
                                 "r"(smem_u32(smem + off)), "r"(chunk)
                                 : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");   // up to 4 tiles in flight (the source is constant)
            }
        }
    }
    if (MODE == 3 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, float* dst, size_t footprint, int grid, int steps, int share, int chunk, long long* dcyc) {
    const size_t ntiles = footprint / TILE;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 512, TILE>>>(dst, ntiles, 4, share, chunk, dcyc);          // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 512, TILE>>>(dst, ntiles, steps, share, chunk, dcyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); exit(1); }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024];
    cudaMemcpy(h, dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += (double)h[i];
    avg /= grid;
    const double bytes = (double)grid * steps * TILE;
    printf("%-28s footprint %5zu MB share %d chunk %5d: %8.3f ms  %7.1f GB/s  %6.1f B/clk/SM  (%.0f cycles per 64 KB tile per CTA)\n", name,
           footprint >> 20, share, chunk, ms, bytes / ms * 1e-6, (double)steps * TILE / avg, avg / steps);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* dst = nullptr;
    const size_t big = 1ull << 30;
    cudaMalloc(&dst, big);
    cudaMemset(dst, 0, big);
    long long* dcyc = nullptr;
    cudaMalloc(&dcyc, sizeof(long long) * 1024);
    const int steps = 400;
    for (size_t fp : {size_t(32) << 20, big}) {
        for (int share : {1, 8}) {
            run<0>("red.f32", dst, fp, sms, steps, share, 0, dcyc);
            run<1>("red.v2.f32", dst, fp, sms, steps, share, 0, dcyc);
            run<2>("red.v4.f32", dst, fp, sms, steps, share, 0, dcyc);
            for (int chunk : {2048, 8192, 16384, 65536}) run<3>("cp.reduce.async.bulk.f32", dst, fp, sms, steps, share, chunk, dcyc);
        }
    }
    // one SM alone (is the limit per SM or global?)
    run<0>("red.f32 (1 CTA)", dst, size_t(32) << 20, 1, steps, 1, 0, dcyc);
    run<2>("red.v4.f32 (1 CTA)", dst, size_t(32) << 20, 1, steps, 1, 0, dcyc);
    run<3>("cp.reduce.bulk (1 CTA)", dst, size_t(32) << 20, 1, steps, 1, 16384, dcyc);
    return 0;
}
