// Tensor-core (tcgen05 / TMEM) fast path for the two-level Delayed-Acceptance hot path with a
// linear forward operator and isotropic likelihoods (BASELINE cfg2), float32 engine only.
//
// Arithmetic: every contraction is a 3xTF32 split product (hi*hi + lo*hi + hi*lo, fp32
// accumulation in TMEM) so that log-likelihood differences keep fp32-grade accuracy -- a single
// TF32 pass would flip accept decisions.
//
// This file also holds a one-CTA self-test GEMM (tda_tc_gemm_selftest in the C ABI) that
// exercises exactly the descriptor / TMEM conventions the DA kernel relies on.
#pragma once
#include <string>
#include <vector>
#include "tda_common.cuh"
#include "tda_tc_prims.cuh"

namespace tda {

// ---------------------------------------------------------------------------------------------
// self-test: D[128][N] = A[128][64] @ B[64][N] on the tensor cores, 3xTF32
//   a_in_tmem = 1: A operand staged in TMEM by tcgen05.st (what the DA kernel does)
//   a_in_tmem = 0: A operand staged in shared memory (canonical K-major, no swizzle)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1)
tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bc_hi, const float* __restrict__ Bc_lo,
                        int N, float* __restrict__ D, int a_in_tmem) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int K = 64;
    float* sB_hi = reinterpret_cast<float*>(smem);
    float* sB_lo = sB_hi + 256 * K;
    float* sA_hi = sB_lo + 256 * K;
    float* sA_lo = sA_hi + 128 * K;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA_lo + 128 * K);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4);
    uint64_t *bar_b = bars, *bar_req = bars + 1, *bar_resp = bars + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 4) tc::tmem_alloc(s_tmem, 512);
    if (tid == 128) {
        tc::mbar_init(bar_b, 1);
        tc::mbar_init(bar_req, 128);
        tc::mbar_init(bar_resp, 1);
        tc::fence_mbar_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *s_tmem;
    const uint32_t bytes_b = (uint32_t)N * K * 4;

    if (warp == 4) {
        if (lane == 0) {
            tc::mbar_expect_tx(bar_b, 2 * bytes_b);
            tc::bulk_g2s(sB_hi, Bc_hi, bytes_b, bar_b);
            tc::bulk_g2s(sB_lo, Bc_lo, bytes_b, bar_b);
            tc::mbar_wait(bar_b, 0);
            tc::mbar_wait(bar_req, 0);
            tc::fence_after_sync();
            const uint32_t idesc = tc::idesc_tf32(128, N);
            const uint32_t d_t = tbase + 128;
            uint32_t accumulate = 0;
            for (int pass = 0; pass < 3; pass++) {
                const int a_part = (pass == 1) ? 1 : 0;     // hi, lo, hi
                const float* sb = (pass == 2) ? sB_lo : sB_hi;
                const float* sa = a_part ? sA_lo : sA_hi;
                for (int k0 = 0; k0 < K; k0 += 8) {
                    uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(sb) + (k0 / 4) * 128, 128, (K / 4) * 128);
                    if (a_in_tmem) {
                        tc::mma_tf32_ts(d_t, tbase + a_part * 64 + k0, bdesc, idesc, accumulate);
                    } else {
                        uint64_t adesc = tc::smem_desc_kmajor(tc::smem_u32(sa) + (k0 / 4) * 128, 128, (K / 4) * 128);
                        tc::mma_tf32_ss(d_t, adesc, bdesc, idesc, accumulate);
                    }
                    accumulate = 1;
                }
            }
            tc::mma_commit(bar_resp);
        }
        __syncwarp();
    } else {
        const int r = tid;                       // row = TMEM lane
        const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
        uint32_t hi[16], lo[16];
        for (int c0 = 0; c0 < K; c0 += 16) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float x = A[r * K + c0 + j];
                float h = tc::tf32_hi(x);
                hi[j] = __float_as_uint(h);
                lo[j] = __float_as_uint(x - h);
                if (!a_in_tmem) {
                    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sA_hi) + tc::canon_offset_f32(r, c0 + j, K)) = h;
                    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sA_lo) + tc::canon_offset_f32(r, c0 + j, K)) = x - h;
                }
            }
            if (a_in_tmem) {
                tc::tmem_st16(lane_base + c0, hi);
                tc::tmem_st16(lane_base + 64 + c0, lo);
            }
        }
        if (a_in_tmem) tc::tmem_wait_st();
        else tc::fence_proxy_async_smem();
        tc::fence_before_sync();
        tc::mbar_arrive(bar_req);
        tc::mbar_wait(bar_resp, 0);
        tc::fence_after_sync();
        uint32_t v[16];
        for (int c0 = 0; c0 < N; c0 += 16) {
            tc::tmem_ld16(lane_base + 128 + c0, v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (c0 + j < N) D[r * N + c0 + j] = __uint_as_float(v[j]);
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tbase, 512);
}

inline int tc_gemm_selftest_host(const float* A, const float* B, int N, float* D, int a_in_tmem, int split, std::string& err) {
    constexpr int K = 64;
    if (N < 8 || N > 256 || (N % 8) != 0) { err = "selftest: N must be a multiple of 8 in [8, 256]"; return -1; }
    std::vector<float> bhi((size_t)N * K), blo((size_t)N * K), a((size_t)128 * K);
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) {
            float x = B[(size_t)k * N + n];
            float h = split ? tc::tf32_hi(x) : x;
            size_t o = tc::canon_offset_f32(n, k, K) / 4;
            bhi[o] = h;
            blo[o] = x - h;
        }
    float *dA = nullptr, *dBh = nullptr, *dBl = nullptr, *dD = nullptr;
    cudaError_t e;
    auto chk = [&](cudaError_t c, const char* what) { if (c != cudaSuccess && err.empty()) err = std::string(what) + ": " + cudaGetErrorString(c); return c; };
    chk(cudaMalloc(&dA, 128 * K * 4), "malloc");
    chk(cudaMalloc(&dBh, (size_t)N * K * 4), "malloc");
    chk(cudaMalloc(&dBl, (size_t)N * K * 4), "malloc");
    chk(cudaMalloc(&dD, (size_t)128 * N * 4), "malloc");
    chk(cudaMemcpy(dA, A, 128 * K * 4, cudaMemcpyHostToDevice), "h2d");
    chk(cudaMemcpy(dBh, bhi.data(), (size_t)N * K * 4, cudaMemcpyHostToDevice), "h2d");
    chk(cudaMemcpy(dBl, blo.data(), (size_t)N * K * 4, cudaMemcpyHostToDevice), "h2d");
    const size_t smem = (size_t)(2 * 256 * K + 2 * 128 * K) * 4 + 64;
    chk(cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr");
    if (err.empty()) {
        tc_gemm_selftest_kernel<<<1, 160, smem>>>(dA, dBh, dBl, N, dD, a_in_tmem);
        chk(cudaGetLastError(), "launch");
        chk(cudaDeviceSynchronize(), "sync");
        chk(cudaMemcpy(D, dD, (size_t)128 * N * 4, cudaMemcpyDeviceToHost), "d2h");
    }
    (void)e;
    cudaFree(dA); cudaFree(dBh); cudaFree(dBl); cudaFree(dD);
    return err.empty() ? 0 : -2;
}

template <typename R>
struct DaTcState {
    std::string err;
    bool eligible(const tda_config&, const Params<R>&) const { return false; }
    int run(Params<R>&, const tda_config&, long long, int, cudaStream_t) { err = "tensor-core path not built"; return -5; }
    void destroy() {}
};

}  // namespace tda
