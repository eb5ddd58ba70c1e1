// Micro-benchmark: throughput of the tc16 RNG-warp inner loop in isolation (Philox4x32-10 + Box-Muller
// + fp16 pack + STS.128), one CTA per SM, NW warps per CTA.  Prints cycles per 128x64 z image per scheduler.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../tinyda_b200/csrc -o rng_bench rng_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tda_common.cuh"
#include "tda_tc_prims.cuh"
using namespace tda;

template <int UNROLL, int ROUNDS>
__device__ __forceinline__ uint4 philox_r(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < ROUNDS; i++) {
        uint32_t hi0, lo0, hi1, lo1;
        mulwide32(ctr.x, M0, hi0, lo0);
        mulwide32(ctr.z, M1, hi1, lo1);
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

template <int UNROLL, int ROUNDS>
__global__ void __launch_bounds__(1024, 1) k(int steps, unsigned long long seed, long long* cyc, uint32_t* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = (warp & 3) * 32 + lane;
    unsigned char* dst = smem + (warp >> 2) * 16384 + (row >> 3) * 1024 + (row & 7) * 16;
    const long long gchain = blockIdx.x * 1024 + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int st = 0; st < steps; st++) {
        const unsigned long long blk0 = (unsigned long long)st * 16;
#pragma unroll UNROLL
        for (int kg = 0; kg < 8; kg++) {
            uint4 c0 = make_uint4((uint32_t)(blk0 + 2 * kg), 0, (uint32_t)gchain, STREAM_Z);
            uint4 c1 = make_uint4((uint32_t)(blk0 + 2 * kg + 1), 0, (uint32_t)gchain, STREAM_Z);
            const uint4 b0 = philox_r<UNROLL, ROUNDS>(c0, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            const uint4 b1 = philox_r<UNROLL, ROUNDS>(c1, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            float s[8];
            bm_pair(b0.x, b0.y, BM_C_X4096, s[0], s[1]);
            bm_pair(b0.z, b0.w, BM_C_X4096, s[2], s[3]);
            bm_pair(b1.x, b1.y, BM_C_X4096, s[4], s[5]);
            bm_pair(b1.z, b1.w, BM_C_X4096, s[6], s[7]);
            uint4 w;
            w.x = tc::pack_f16x2(s[0], s[1]); w.y = tc::pack_f16x2(s[2], s[3]);
            w.z = tc::pack_f16x2(s[4], s[5]); w.w = tc::pack_f16x2(s[6], s[7]);
            *reinterpret_cast<uint4*>(dst + kg * 128) = w;
        }
        __syncwarp();
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = *reinterpret_cast<uint32_t*>(smem + threadIdx.x * 4);
}

template <int UNROLL, int ROUNDS>
void run(int nw, const char* name) {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
    const int steps = 2000;
    size_t smem = (size_t)(nw / 4) * 16384;
    cudaFuncSetAttribute(k<UNROLL, ROUNDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<UNROLL, ROUNDS><<<148, nw * 32, smem>>>(steps, 1234, cyc, sink);
    k<UNROLL, ROUNDS><<<148, nw * 32, smem>>>(steps, 1234, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    // images per step per CTA = nw/4; per scheduler: nw/4 warps... report cycles per image-per-scheduler-warp
    printf("%-28s warps %2d: %8.0f cycles/step, %6.1f cycles per warp-normal-row (64 normals x 32 lanes), IPC-equiv %s (%s)\n", name, nw, c / steps,
           c / steps / (nw / 4.0), "-", cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    for (int nw : {4, 8, 16}) {
        run<1, 10>(nw, "philox10 unroll1");
        run<2, 10>(nw, "philox10 unroll2");
        run<4, 10>(nw, "philox10 unroll4");
        run<2, 7>(nw, "philox7 unroll2");
    }
    return 0;
}
