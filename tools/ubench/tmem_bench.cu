// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps issuing
// (the row warps of the Delayed-Acceptance kernels read accumulator tiles and keep chain state in TMEM).
// One CTA per SM, NW warps, every warp loops over x16 loads (or stores) of its own lane quarter.
// Prints bytes per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../tinyda_b200/csrc -o tmem_bench tmem_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tda_tc_prims.cuh"
using namespace tda;

__global__ void __launch_bounds__(1024, 1) k(int iters, int mode, long long* cyc, uint32_t* sink) {
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[16], acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const uint32_t col = (uint32_t)(((it * 16) + (warp >> 2) * 64) & 511) & ~15u;
        if (mode == 0) {
            tc::tmem_ld16(base + col, v);
            if ((it & 3) == 3) {
                tc::tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i++) acc += v[i];
            }
        } else {
            tc::tmem_st16(base + col, v);
            if ((it & 3) == 3) tc::tmem_wait_st();
        }
    }
    if (mode == 0) tc::tmem_wait_ld(); else tc::tmem_wait_st();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(s_tmem, 512);
}

int main() {
    long long* d_cyc; uint32_t* d_sink;
    cudaMalloc(&d_cyc, 8); cudaMalloc(&d_sink, 4);
    const int iters = 4096;
    for (int mode = 0; mode < 2; mode++)
        for (int nw : {4, 8, 16, 32}) {
            k<<<148, nw * 32>>>(iters, mode, d_cyc, d_sink);
            k<<<148, nw * 32>>>(iters, mode, d_cyc, d_sink);
            long long c = 0;
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)nw * iters * 16 * 32 * 4;
            printf("%s x16  warps=%2d  cycles=%lld  bytes/clk/SM=%.1f  cycles per x16 per warp=%.1f  (%s)\n", mode ? "tcgen05.st" : "tcgen05.ld", nw, c,
                   bytes / (double)c, (double)c / iters, cudaGetErrorString(e));
        }
    return 0;
}
