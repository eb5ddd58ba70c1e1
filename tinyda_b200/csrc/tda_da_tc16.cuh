// Tensor-core (tcgen05 / TMEM) Delayed-Acceptance kernel, third generation ("tc16"): two-level DA,
// pCN proposal, linear forward operators, Gaussian likelihoods (BASELINE cfg2), float32 engine.
// Native shape d = 64, m_c a multiple of 16 (<= 128), m_f a multiple of 64 (<= 1920), isotropic
// likelihoods; prepare() maps the rest onto it on the host: d in {16, 32, 48} and other output counts
// by zero padding (kernel template PAD), diagonal / dense likelihoods (distributions.py:304-315,
// :246-301) by whitening the operators and the data with the Cholesky factor of the precision.
// History: fine-level Links, and -- the reference's default store_coarse_chain=True,
// sampler.py:421-436 -- the parameters, log-likelihood and accept flag of every coarse Link; the
// fields the kernel does not hold (coarse log-prior, Link.model_output) are rebuilt from the recorded
// parameters when first fetched (EngineT::fill_lazy_history in tda_engine.cu).
//
// What changed against tda_da_tc.cuh (3xTF32, draws on the row threads):
//
//  * Arithmetic: every operand is a TWO-TERM FP16 SPLIT at a power-of-two scale chosen on the host
//    (x * 2^s = hi + lo, hi = fp16(x 2^s), lo = fp16(x 2^s - hi)); products are
//    hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM.  fp16 carries the same 11 significant
//    bits as tf32, so the accuracy is that of the 3xTF32 scheme (~2^-22 per element), but
//    kind::f16 runs at twice the tf32 rate.  The scales keep every hi part in [2^3, 2^15) so the
//    lo parts stay in fp16's normal range; an overflowing proposal yields NaN -> alpha = 0.
//  * Normals: Philox4x32-10 + Box-Muller run on DEDICATED warps that stay one coarse step ahead
//    and write z straight into shared memory as an MMA A operand.  The stream is the "z16"
//    stream of tda_common.cuh (normals on the fp16 grid), so z is ONE exact operand image.
//  * One MMA hop per coarse step: with theta' = a theta + b z T (pCN), the coarse model output is
//        F_c(theta') = theta @ (a G_c^T)  +  z @ (b T G_c^T)
//    i.e. two accumulating MMA groups into the same TMEM columns (A = theta from TMEM, 3 products;
//    A = z from shared memory, 2 products) with the composed operator b T G_c^T built once on the
//    host; xi = z @ T goes to 64 more columns and is only read when a chain accepts
//    (theta <- a theta + b xi).  The row threads keep A_theta = split(theta) up to date.
//  * One MMA-issuing warp per tile (blocking mbarrier waits instead of a polling loop).
//  * Register budget moved between warpgroups with setmaxnreg.
//
// Reference semantics: chain.py:325-444 (DAChain), proposal.py:261-369 (pCN), exactly as
// Tile::base_step / Tile::upper_step in tda_kernels.cuh; the state buffers are shared with the other
// kernels, so runs can be interleaved.
#pragma once
#include <cuda_fp16.h>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>
#include "tda_common.cuh"
#include "tda_tc_prims.cuh"

namespace tda {

constexpr int T16_K = 64;                       // parameters (contraction length)
constexpr int T16_CH = 64;                      // columns per streamed fine-operator chunk
constexpr int T16_NST = 4;                      // ring stages
constexpr int T16_MAX_MC = 128;
constexpr int T16_MAX_MF = 1920;                // the negated data vector stays resident in shared memory
// The warp scheduler prefers the highest warp id among ready warps: the MMA issuers come first,
// then the RNG warps (the throughput-critical producers; alone they need ~4400 cycles per coarse
// step of both tiles), then the row warps, which fill the remaining issue slots.
constexpr int T16_ROW_WARP0 = 0;                // warps 0..15  (warpgroups 0-3)
constexpr int T16_RNG_WARP0 = 16;               // warps 16..23 (warpgroups 4-5)
constexpr int T16_RNG_WARPS = 8;
constexpr int T16_MMA_WARP0 = 24;               // warps 24, 25 (one per tile)
constexpr int T16_PROD_WARP = 26;               // warp 27 idles (completes warpgroup 6)
constexpr int T16_THREADS = 28 * 32;
constexpr int T16_HK = T16_K / 2;               // theta columns per row thread
constexpr int T16_IMG = 128 * T16_K * 2;        // one [128 x 64] fp16 image: 16 KB
constexpr int T16_TIMG = 64 * T16_K * 2;        // one [64 x 64] fp16 image: 8 KB
constexpr int T16_CHUNK_BYTES = 2 * T16_TIMG;   // hi + lo
#ifndef T16_RNG_ILP
#define T16_RNG_ILP 1          // 16-normal groups a generator thread has in flight; 2 (with 56 / 88 registers for generator / row threads) measured 780 M against 797 M transitions/s: the generator is not latency-bound, it competes with the row threads for issue slots
#endif
#if T16_RNG_ILP == 2
constexpr int T16_REGS_ROW = 88, T16_REGS_RNG = 56, T16_REGS_AUX = 40;   // 512*16 taken = 256*(-16) ... see the budget below
#else
constexpr int T16_REGS_ROW = 96, T16_REGS_RNG = 40, T16_REGS_AUX = 40;
#endif   // 512*24 taken = 256*32 + 128*32 released (launch: 72)

constexpr int T16_OFF_G = 0;                                   // a G_c^T   (hi | lo)
constexpr int T16_OFF_M = T16_OFF_G + 2 * T16_IMG;             // b T G_c^T (hi | lo)
constexpr int T16_OFF_T = T16_OFF_M + 2 * T16_IMG;             // T         (hi | lo)
constexpr int T16_OFF_RING = T16_OFF_T + 2 * T16_TIMG;
constexpr int T16_OFF_Z = T16_OFF_RING + T16_NST * T16_CHUNK_BYTES;   // [tile 2][buffer 2] images
constexpr int T16_OFF_PART = T16_OFF_Z + 4 * T16_IMG;          // [buf 2][tile 2][half 2][128] f32
constexpr int T16_OFF_U = T16_OFF_PART + 2 * 2 * 2 * 128 * 4;  // [buf 2][tile 2][128] f32
constexpr int T16_OFF_PF = T16_OFF_U + 2 * 2 * 128 * 4;        // [tile 2][half 2][val 2][128] f32
constexpr int T16_OFF_NY = T16_OFF_PF + 2 * 2 * 2 * 128 * 4;   // f32: [-y_c 128 | -(mu @ LP) 64 | -y_f mf]
constexpr int T16_NY_FLOATS = T16_MAX_MC + T16_K + T16_MAX_MF;
constexpr int T16_OFF_BARS = T16_OFF_NY + T16_NY_FLOATS * 4;
constexpr size_t T16_SMEM_BYTES = T16_OFF_BARS + 512;
static_assert(T16_SMEM_BYTES <= 232448, "shared memory budget");


struct DaTc16Params {
    const __half* G_hl;       // [hi mc x 64 | lo mc x 64] canonical K-major, a G_c^T  * 2^sG
    const __half* M_hl;       // [hi mc x 64 | lo mc x 64]                  b T G_c^T * 2^sM
    const __half* T_hl;       // [hi 64 x 64 | lo 64 x 64]                  T         * 2^sT
    const __half* F_chunks;   // n_chunks x [hi 64 x 64 | lo 64 x 64]  fine operator columns, then LP
    const float* ny;          // [-(y_c - b_c) 128 | -(mu @ LP) 64 | -(y_f - b_f) mf]
    int mc, mf, n_chunks, J;
    float var_c, var_f, prior_logconst;
    float ca;                 // sqrt(1 - beta^2)
    float cxi;                // beta * 2^s_theta / 2^(s_z + s_T): scaled-theta increment per unit of D_xi
    float sc_c, sc_f, sc_p;   // accumulator -> model output: 2^-(s_theta + s_G), 2^-(s_theta + s_Gf), 2^-(s_theta + s_LP)
    float th_scale, th_unscale;
    int n_pairs;              // units per block of iterations: tile pairs, or single tiles when `solo`
    int solo;                 // fewer tile pairs than half the SMs: every CTA advances ONE 128-chain tile (the warps of
                              // the second tile idle), so that twice as many SMs work (strong scaling)
    int half;                 // (with solo) fewer tiles than half the SMs: tiles of 64 chains -- only the warps of TMEM
                              // lanes 0..63 work, the MMA's rows 64..127 are padding -- so that the SMs fill up again
    int ib, nb;               // work units: iterations per block, blocks per launch (unit = tile pair x block)
    int* progress;            // [n_pairs] tile completions of this launch (2 per finished block)
    long long* dbg;           // optional timeline probe (tools/tc16_timeline.py): [role 4][256] clock64 stamps of CTA 0
    PhiloxRoundKeys rk;       // round keys of the engine's seed (constant-bank operands of the generator warps)
    int direct_wait;          // single-tile modes: every row warp waits on the MMA's mbarrier itself
};

// ---------------------------------------------------------------------------------------------
// self-test: D[128][N] = A[128][64] @ B[64][N], fp16 two-term split, the conventions of the DA
// kernel (A packed in TMEM or canonical in shared memory, B canonical in shared memory).
// A is scaled by sA in the kernel; B arrives pre-scaled / pre-split; D is unscaled by the host.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1)
tc16_gemm_selftest_kernel(const float* __restrict__ A, const __half* __restrict__ B_hl, int N, float sA,
                          float* __restrict__ D, int a_in_tmem) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int K = T16_K;
    unsigned char* sB_hi = smem;                       // up to 256 x 64 halves = 32 KB
    unsigned char* sB_lo = smem + 256 * K * 2;
    unsigned char* sA_hi = smem + 2 * 256 * K * 2;     // 16 KB each
    unsigned char* sA_lo = sA_hi + T16_IMG;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA_lo + T16_IMG);
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
    const uint32_t bytes_b = (uint32_t)N * K * 2;

    if (warp == 4) {
        if (lane == 0) {
            tc::mbar_expect_tx(bar_b, 2 * bytes_b);
            tc::bulk_g2s(sB_hi, B_hl, bytes_b, bar_b);
            tc::bulk_g2s(sB_lo, B_hl + (size_t)N * K, bytes_b, bar_b);
            tc::mbar_wait(bar_b, 0);
            tc::mbar_wait(bar_req, 0);
            tc::fence_after_sync();
            const uint32_t idesc = tc::idesc_f16(128, N);
            const uint32_t d_t = tbase + 128;
            uint32_t accumulate = 0;
            for (int pass = 0; pass < 3; pass++) {
                const int a_lo = (pass == 1);
                const unsigned char* sb = (pass == 2) ? sB_lo : sB_hi;
                const unsigned char* sa = a_lo ? sA_lo : sA_hi;
                for (int ks = 0; ks < K / 16; ks++) {
                    const uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(sb) + ks * 256, 128, (K / 8) * 128);
                    if (a_in_tmem) {
                        tc::mma_f16_ts(d_t, tbase + a_lo * 32 + ks * 8, bdesc, idesc, accumulate);
                    } else {
                        const uint64_t adesc = tc::smem_desc_kmajor(tc::smem_u32(sa) + ks * 256, 128, (K / 8) * 128);
                        tc::mma_f16_ss(d_t, adesc, bdesc, idesc, accumulate);
                    }
                    accumulate = 1;
                }
            }
            tc::mma_commit(bar_resp);
        }
        __syncwarp();
    } else {
        const int r = tid;
        const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
        uint32_t hi[16], lo[16];
        for (int c0 = 0; c0 < K; c0 += 32) {
#pragma unroll
            for (int j = 0; j < 16; j++)
                tc::split_f16x2(A[r * K + c0 + 2 * j] * sA, A[r * K + c0 + 2 * j + 1] * sA, hi[j], lo[j]);
            if (a_in_tmem) {
                tc::tmem_st16(lane_base + c0 / 2, hi);
                tc::tmem_st16(lane_base + 32 + c0 / 2, lo);
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    *reinterpret_cast<uint32_t*>(sA_hi + tc::canon_offset_f16(r, c0 + 2 * j, K)) = hi[j];
                    *reinterpret_cast<uint32_t*>(sA_lo + tc::canon_offset_f16(r, c0 + 2 * j, K)) = lo[j];
                }
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

// power-of-two scale that brings `maxabs` into [2^(top-1), 2^top)
inline int pow2_scale_for(double maxabs, int top) {
    if (!(maxabs > 0.0) || !std::isfinite(maxabs)) return 0;
    int e;
    std::frexp(maxabs, &e);          // maxabs = f * 2^e, f in [0.5, 1)
    return top - e;
}

// canonical K-major fp16 hi / lo images of B[n][k] = factor * W[k][n0 + n] * 2^s   (W row-major [K][ldw])
inline void canon_split16(const std::vector<double>& W, int ldw, int n0, int rows, double factor, int s, __half* hi, __half* lo) {
    // factor * 2^s once: scaling by a power of two commutes with the rounding of factor * W (no
    // subnormals at these magnitudes), so the images are the ones ldexp(factor * W, s) gives
    const double scale = std::ldexp(factor, s);
    for (int k = 0; k < T16_K; k++)
        for (int n = 0; n < rows; n++) {
            const float x = (float)(scale * W[(size_t)k * ldw + n0 + n]);
            const __half h = __float2half_rn(x);
            const __half l = __float2half_rn(x - __half2float(h));
            const size_t o = tc::canon_offset_f16(n, k, T16_K) / 2;
            hi[o] = h;
            lo[o] = l;
        }
}

inline int tc16_gemm_selftest_host(const float* A, const float* B, int N, float* D, int a_in_tmem, std::string& err) {
    constexpr int K = T16_K;
    if (N < 16 || N > 256 || (N % 16) != 0) { err = "selftest16: N must be a multiple of 16 in [16, 256]"; return -1; }
    double ma = 0, mb = 0;
    for (int i = 0; i < 128 * K; i++) ma = std::fmax(ma, std::fabs((double)A[i]));
    for (int i = 0; i < K * N; i++) mb = std::fmax(mb, std::fabs((double)B[i]));
    const int sa = pow2_scale_for(ma, 14), sb = pow2_scale_for(mb, 14);
    std::vector<double> W((size_t)K * N);
    for (size_t i = 0; i < W.size(); i++) W[i] = B[i];
    std::vector<__half> bhl((size_t)2 * N * K);
    canon_split16(W, N, 0, N, 1.0, sb, bhl.data(), bhl.data() + (size_t)N * K);
    float *dA = nullptr, *dD = nullptr;
    __half* dB = nullptr;
    auto chk = [&](cudaError_t c, const char* what) { if (c != cudaSuccess && err.empty()) err = std::string(what) + ": " + cudaGetErrorString(c); return c; };
    chk(cudaMalloc(&dA, 128 * K * 4), "malloc");
    chk(cudaMalloc(&dB, bhl.size() * 2), "malloc");
    chk(cudaMalloc(&dD, (size_t)128 * N * 4), "malloc");
    chk(cudaMemcpy(dA, A, 128 * K * 4, cudaMemcpyHostToDevice), "h2d");
    chk(cudaMemcpy(dB, bhl.data(), bhl.size() * 2, cudaMemcpyHostToDevice), "h2d");
    const size_t smem = (size_t)2 * 256 * K * 2 + 2 * T16_IMG + 64;
    chk(cudaFuncSetAttribute(tc16_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr");
    if (err.empty()) {
        tc16_gemm_selftest_kernel<<<1, 160, smem>>>(dA, dB, N, (float)std::ldexp(1.0, sa), dD, a_in_tmem);
        chk(cudaGetLastError(), "launch");
        chk(cudaDeviceSynchronize(), "sync");
        chk(cudaMemcpy(D, dD, (size_t)128 * N * 4, cudaMemcpyDeviceToHost), "d2h");
        const float un = (float)std::ldexp(1.0, -(sa + sb));
        for (size_t i = 0; i < (size_t)128 * N; i++) D[i] *= un;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return err.empty() ? 0 : -2;
}

// ---------------------------------------------------------------------------------------------
// the DA kernel
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T* t16_opaque(T* ptr) {
    asm volatile("" : "+l"(ptr));
    return ptr;
}

// "my TMEM / shared-memory accesses are done" -> one arrival per warp
__device__ __forceinline__ void t16_warp_arrive(uint64_t* bar, int lane) {
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(bar);
}

// 3 products (A = theta hi | lo in TMEM, B hi | lo in shared memory), N columns, K = 16 nks (64 unless d < 64)
__device__ __forceinline__ void t16_issue_theta(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, uint32_t idesc, uint32_t accumulate,
                                                int nks = T16_K / 16) {
#pragma unroll
    for (int pass = 0; pass < 3; pass++) {
        const uint32_t a = a_tmem + (pass == 1 ? 32 : 0);
        const uint64_t b0 = tc::smem_desc_kmajor(pass == 2 ? b_lo : b_hi, 128, (T16_K / 8) * 128);
#pragma unroll
        for (int ks = 0; ks < T16_K / 16; ks++) {
            if (ks < nks) {
                tc::mma_f16_ts(d_tmem, a + ks * 8, b0 + (uint64_t)(ks * 16), idesc, accumulate);
                accumulate = 1;
            }
        }
    }
}

// z-operand products: z @ B_hi + z @ B_lo (+ z_lo @ B_hi when the stream is injected)
__device__ __forceinline__ void t16_issue_z(uint32_t d_tmem, uint32_t z_hi, uint32_t z_lo, bool with_lo, uint32_t b_hi, uint32_t b_lo,
                                            uint32_t idesc, uint32_t accumulate, int nks = T16_K / 16) {
    const int npass = with_lo ? 3 : 2;
    for (int pass = 0; pass < npass; pass++) {
        const uint64_t a0 = tc::smem_desc_kmajor(pass == 2 ? z_lo : z_hi, 128, (T16_K / 8) * 128);
        const uint64_t b0 = tc::smem_desc_kmajor(pass == 1 ? b_lo : b_hi, 128, (T16_K / 8) * 128);
#pragma unroll
        for (int ks = 0; ks < T16_K / 16; ks++) {
            if (ks < nks) {
                tc::mma_f16_ss(d_tmem, a0 + (uint64_t)(ks * 16), b0 + (uint64_t)(ks * 16), idesc, accumulate);
                accumulate = 1;
            }
        }
    }
}

// One 16-normal group (z16 stream) of one chain, packed to fp16 and stored into the chain's row of
// a canonical K-major z image (row_ptr = image + row offset; columns 16 q4 .. 16 q4 + 15).
__device__ __forceinline__ void t16_z_group(const PhiloxRoundKeys& rk, long long gchain, unsigned long long group, unsigned char* row_ptr, int q4) {
    const unsigned long long b3 = 3 * group;
    float s[16];
    z16_group_scaled(philox_block_z_rk(rk, gchain, b3), philox_block_z_rk(rk, gchain, b3 + 1), philox_block_z_rk(rk, gchain, b3 + 2), s);
    uint4 w0, w1;
    w0.x = tc::pack_f16x2(s[0], s[1]);   w0.y = tc::pack_f16x2(s[2], s[3]);
    w0.z = tc::pack_f16x2(s[4], s[5]);   w0.w = tc::pack_f16x2(s[6], s[7]);
    w1.x = tc::pack_f16x2(s[8], s[9]);   w1.y = tc::pack_f16x2(s[10], s[11]);
    w1.z = tc::pack_f16x2(s[12], s[13]); w1.w = tc::pack_f16x2(s[14], s[15]);
    *reinterpret_cast<uint4*>(row_ptr + (2 * q4) * 128) = w0;
    *reinterpret_cast<uint4*>(row_ptr + (2 * q4 + 1) * 128) = w1;
}

// Two groups of one chain at once: six independent Philox blocks and sixteen Box-Muller pairs in one basic block,
// so that the generator warps -- two per scheduler, each a chain of dependent integer multiplies and special-function
// results -- overlap their latencies (the z images pace the paired mode: ~6000 cycles per image with one group after
// the other, measured with the clock64 probe).
__device__ __forceinline__ void t16_z_group2(const PhiloxRoundKeys& rk, long long gchain, unsigned long long group0, unsigned char* row_ptr,
                                             int qa, int qb) {
    const unsigned long long a3 = 3 * (group0 + qa), b3 = 3 * (group0 + qb);
    const uint4 pa0 = philox_block_z_rk(rk, gchain, a3), pb0 = philox_block_z_rk(rk, gchain, b3);
    const uint4 pa1 = philox_block_z_rk(rk, gchain, a3 + 1), pb1 = philox_block_z_rk(rk, gchain, b3 + 1);
    const uint4 pa2 = philox_block_z_rk(rk, gchain, a3 + 2), pb2 = philox_block_z_rk(rk, gchain, b3 + 2);
    float sa[16], sb[16];
    z16_group_scaled(pa0, pa1, pa2, sa);
    z16_group_scaled(pb0, pb1, pb2, sb);
    uint4 w0, w1, v0, v1;
    w0.x = tc::pack_f16x2(sa[0], sa[1]);   w0.y = tc::pack_f16x2(sa[2], sa[3]);
    w0.z = tc::pack_f16x2(sa[4], sa[5]);   w0.w = tc::pack_f16x2(sa[6], sa[7]);
    w1.x = tc::pack_f16x2(sa[8], sa[9]);   w1.y = tc::pack_f16x2(sa[10], sa[11]);
    w1.z = tc::pack_f16x2(sa[12], sa[13]); w1.w = tc::pack_f16x2(sa[14], sa[15]);
    v0.x = tc::pack_f16x2(sb[0], sb[1]);   v0.y = tc::pack_f16x2(sb[2], sb[3]);
    v0.z = tc::pack_f16x2(sb[4], sb[5]);   v0.w = tc::pack_f16x2(sb[6], sb[7]);
    v1.x = tc::pack_f16x2(sb[8], sb[9]);   v1.y = tc::pack_f16x2(sb[10], sb[11]);
    v1.z = tc::pack_f16x2(sb[12], sb[13]); v1.w = tc::pack_f16x2(sb[14], sb[15]);
    *reinterpret_cast<uint4*>(row_ptr + (2 * qa) * 128) = w0;
    *reinterpret_cast<uint4*>(row_ptr + (2 * qa + 1) * 128) = w1;
    *reinterpret_cast<uint4*>(row_ptr + (2 * qb) * 128) = v0;
    *reinterpret_cast<uint4*>(row_ptr + (2 * qb + 1) * 128) = v1;
}

// Work distribution.  A launch advances n_pairs tile pairs by `iterations`; 256 pairs on 148 SMs would
// leave 27 % of the machine idle in the second wave, so the iterations are cut into nb blocks and the
// units (pair, block) are dealt round-robin, block-major: CTA c runs units c, c + grid, c + 2 grid, ...
// Consecutive blocks of a pair generally run on different SMs; the chain state travels through global
// memory (as it does between launches) and `progress[pair]` orders the hand-off.  Unit u depends only
// on unit u - n_pairs, every CTA walks its units in increasing order and all CTAs are co-resident
// (grid <= #SMs, one CTA per SM), so the wait cannot deadlock.
struct T16Unit {
    int pair, blk, it0, it1;       // 32-bit on purpose: the RNG / MMA warps run on 40 registers
};
__device__ __forceinline__ bool t16_unit(const DaTc16Params& q, int iters, int k, T16Unit& u) {
    const int idx = (int)blockIdx.x + k * (int)gridDim.x;
    if (idx >= q.n_pairs * q.nb) return false;
    u.blk = idx / q.n_pairs;
    u.pair = idx - u.blk * q.n_pairs;
    u.it0 = u.blk * q.ib;
    u.it1 = min(u.it0 + q.ib, iters);
    return true;
}
__device__ __forceinline__ int t16_ld_acquire(const int* ptr) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void t16_red_release_add(int* ptr, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

// (Measured on B200 and rejected: letting the row threads generate a quarter or a half of the
// normals one step ahead, in the shadow of the MMA -- 678 M / 636 M transitions/s against 701 M with
// the RNG warps alone; the row threads are the critical path as soon as the normals are ahead.)

// PAD = false: d == 64, the native shape (no guards on the parameter rows); PAD = true: d in {16, 32, 48}
template <bool PAD>
__global__ void __launch_bounds__(T16_THREADS, 1)
da_tc16_kernel(const __grid_constant__ Params<float> p, const __grid_constant__ DaTc16Params q) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sG = smem + T16_OFF_G;
    unsigned char* sM = smem + T16_OFF_M;
    unsigned char* sT = smem + T16_OFF_T;
    unsigned char* ring = smem + T16_OFF_RING;
    unsigned char* zbuf = smem + T16_OFF_Z;
    float* s_part = reinterpret_cast<float*>(smem + T16_OFF_PART);
    float* s_u = reinterpret_cast<float*>(smem + T16_OFF_U);
    float* s_pf = reinterpret_cast<float*>(smem + T16_OFF_PF);
    const float* s_ny = reinterpret_cast<const float*>(smem + T16_OFF_NY);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T16_OFF_BARS);
    uint64_t* bar_res = bars;                       // resident operands landed
    uint64_t* bar_reqA = bars + 1;                  // [tile]     rows -> MMA: D[0:128) consumed
    uint64_t* bar_reqB = bars + 3;                  // [tile]     rows -> MMA: A_theta current, D_xi consumed
    uint64_t* bar_respA = bars + 5;                 // [tile]     MMA  -> rows: F_c(theta') complete in D[0:128)
    uint64_t* bar_respB = bars + 7;                 // [tile]     MMA  -> rows: xi complete in D[128:192)
    uint64_t* bar_reqF = bars + 9;                  // [tile][3] rows -> MMA: fine-chunk accumulator consumed
    uint64_t* bar_respF = bars + 15;                // [tile][3] MMA  -> rows: fine chunk complete
    uint64_t* bar_zfull = bars + 21;                // [tile][buffer] RNG  -> MMA
    uint64_t* bar_zfree = bars + 25;                // [tile][buffer] MMA  -> RNG
    uint64_t* bar_full = bars + 29;                 // [NST]
    uint64_t* bar_empty = bars + 29 + T16_NST;      // [NST]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 29 + 2 * T16_NST);

    // the shuffle tells the compiler that the warp index (and everything derived from it: tile,
    // TMEM / shared-memory bases, MMA descriptors) is warp-uniform -> uniform registers
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (warp == T16_MMA_WARP0) tc::tmem_alloc(s_tmem, 512);
    if (tid == T16_PROD_WARP * 32) {
        tc::mbar_init(bar_res, 1);
        const int nrw = q.half ? 4 : 8;                 // row warps of a tile that take part
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(bar_reqA + i, nrw);
            tc::mbar_init(bar_reqB + i, nrw);
            tc::mbar_init(bar_respA + i, 1);
            tc::mbar_init(bar_respB + i, 1);
        }
        for (int i = 0; i < 6; i++) {
            tc::mbar_init(bar_reqF + i, nrw);
            tc::mbar_init(bar_respF + i, 1);
        }
        for (int i = 0; i < 4; i++) {
            tc::mbar_init(bar_zfull + i, q.solo ? 8 : 4);          // single-tile modes: all eight RNG warps serve the tile
            tc::mbar_init(bar_zfree + i, 1);
        }
        for (int s = 0; s < T16_NST; s++) { tc::mbar_init(bar_full + s, 1); tc::mbar_init(bar_empty + s, q.solo ? 1 : 2); }
        tc::fence_mbar_init();
    }
    // d < 64 (multiples of 16): the operators are zero-padded to 64 rows on the host and the normals'
    // columns d..63 stay zero -- the RNG warps only write the first d columns of a z image
    const int d = PAD ? p.d : T16_K;
    if (PAD || q.half) {
        uint4* zz = reinterpret_cast<uint4*>(zbuf);
        for (int i = tid; i < 4 * T16_IMG / 16; i += T16_THREADS) zz[i] = make_uint4(0u, 0u, 0u, 0u);
        tc::fence_proxy_async_smem();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = __shfl_sync(0xffffffffu, *s_tmem, 0);

    const int J = q.J, mc = q.mc, NCH = q.n_chunks;
    const int iters = (int)p.iterations;          // per launch; checked on the host
    const bool inj = p.rng_mode == TDA_RNG_INJECTED;
    const bool solo = q.solo != 0;
    const bool half = q.half != 0;                 // implies solo
    const int nbt = half ? 128 : 256;              // threads of a tile's named barriers

    if (warp >= T16_RNG_WARP0 && warp < T16_RNG_WARP0 + T16_RNG_WARPS) {
        // =====================================================================================
        // RNG warps: thread = one chain of one tile; 64 normals per coarse step -> z image(s)
        // =====================================================================================
        tc::setmaxnreg_dec<T16_REGS_RNG>();
        // paired mode: four warps per tile, a thread draws the 64 normals of its chain.  Single-tile modes: ALL
        // eight RNG warps work for the one tile -- the generator's latency per z image (a serial chain of
        // ~250 instructions per 16-normal group) paces a lone tile -- so a chain's groups are dealt to 2
        // (128-chain tile) or 4 (64-chain tile) threads.
        const int wi = warp - T16_RNG_WARP0;
        const int t = solo ? 0 : (wi >> 2);
        const int row = half ? (wi & 1) * 32 + lane : (wi & 3) * 32 + lane;
        const int part = half ? (wi >> 1) : solo ? (wi >> 2) : 0;
        const int nparts = half ? 4 : solo ? 2 : 1;
        unsigned char* zt = zbuf + (size_t)t * 2 * T16_IMG + (row >> 3) * ((T16_K / 8) * 128) + (row & 7) * 16;
        uint64_t* zfull = bar_zfull + t * 2;
        uint64_t* zfree = bar_zfree + t * 2;
        unsigned n = 0;                                    // coarse steps produced
        T16Unit un;
        for (int uk = 0; t16_unit(q, iters, uk, un); uk++) {
            const int g = half ? un.pair * 64 + row : solo ? un.pair * 128 + row : un.pair * 256 + t * 128 + row;
            const long long gchain = p.chain_offset + g;
            const bool live = g < p.C;
            long long tb = p.t_base + (long long)un.it0 * J;
            const int nsteps = (un.it1 - un.it0) * J;
            for (int st = 0; st < nsteps; st++, n++, tb++) {
                const int b = inj ? 0 : (int)(n & 1);
                const unsigned use = inj ? n : (n >> 1);
                const bool dbg_on = q.dbg && blockIdx.x == 0 && warp == T16_RNG_WARP0 && lane == 0 && n < 64;
                if (dbg_on) q.dbg[0 * 256 + 3 * n] = clock64();
                if (use >= 1) tc::mbar_wait(zfree + b, (uint32_t)((use - 1) & 1));
                if (dbg_on) q.dbg[0 * 256 + 3 * n + 1] = clock64();
                unsigned char* dst = zt + (size_t)b * T16_IMG;
                if (!inj) {
                    // z16 stream: d normals = d / 16 groups of 3 Philox blocks
                    const unsigned long long grp0 = (unsigned long long)(tb * (d >> 4));
#if T16_RNG_ILP == 2
                    {
                        const int ng = d >> 4;
                        int q4 = part;
#pragma unroll 1
                        for (; q4 + nparts < ng; q4 += 2 * nparts) t16_z_group2(q.rk, gchain, grp0, dst, q4, q4 + nparts);
                        if (q4 < ng) t16_z_group(q.rk, gchain, grp0 + q4, dst, q4);
                    }
#else
#pragma unroll 1
                    for (int q4 = part; q4 < (d >> 4); q4 += nparts) t16_z_group(q.rk, gchain, grp0 + q4, dst, q4);
#endif
                } else {
                    const long long z0 = tb * d;
                    for (int kg = part; kg < (d >> 3); kg += nparts) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const long long idx = z0 + kg * 8 + 2 * i;
                            const float x0 = (live && idx < p.zlen) ? p.zs[(size_t)g * p.zlen + idx] * Z16_SCALE : 0.0f;
                            const float x1 = (live && idx + 1 < p.zlen) ? p.zs[(size_t)g * p.zlen + idx + 1] * Z16_SCALE : 0.0f;
                            tc::split_f16x2(x0, x1, hi[i], lo[i]);
                        }
                        *reinterpret_cast<uint4*>(dst + kg * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(dst + T16_IMG + kg * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                tc::fence_proxy_async_smem();          // generic-proxy writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(zfull + b);
                if (dbg_on) q.dbg[0 * 256 + 3 * n + 2] = clock64();
            }
        }
    } else if (warp >= T16_MMA_WARP0) {
        tc::setmaxnreg_dec<T16_REGS_AUX>();
        if (warp == T16_PROD_WARP) {
            // ===== producer: resident operands once, then the fine-operator chunk ring =====
            if (lane == 0) {
                unsigned total_chunks = 0;
                {
                    T16Unit un;
                    for (int uk = 0; t16_unit(q, iters, uk, un); uk++) total_chunks += (unsigned)(un.it1 - un.it0) * NCH;
                }
                const uint32_t bI = (uint32_t)mc * T16_K * 2;       // one mc-row image
                const uint32_t bNY = (uint32_t)(T16_MAX_MC + T16_K + q.mf) * 4;
                tc::mbar_expect_tx(bar_res, 4 * bI + 2 * T16_TIMG + bNY);
                tc::bulk_g2s(smem + T16_OFF_NY, q.ny, bNY, bar_res);
                tc::bulk_g2s(sG, q.G_hl, bI, bar_res);
                tc::bulk_g2s(sG + T16_IMG, q.G_hl + (size_t)mc * T16_K, bI, bar_res);
                tc::bulk_g2s(sM, q.M_hl, bI, bar_res);
                tc::bulk_g2s(sM + T16_IMG, q.M_hl + (size_t)mc * T16_K, bI, bar_res);
                tc::bulk_g2s(sT, q.T_hl, 2 * T16_TIMG, bar_res);
                for (unsigned g = 0; g < total_chunks; g++) {
                    const int st = (int)(g % T16_NST);
                    if (g >= T16_NST) tc::mbar_wait(bar_empty + st, (uint32_t)(((g / T16_NST) - 1) & 1));
                    const int c = (int)(g % NCH);
                    tc::mbar_expect_tx(bar_full + st, T16_CHUNK_BYTES);
                    tc::bulk_g2s(ring + (size_t)st * T16_CHUNK_BYTES, q.F_chunks + (size_t)c * (T16_CHUNK_BYTES / 2), T16_CHUNK_BYTES, bar_full + st);
                }
            }
        } else if (warp < T16_PROD_WARP) {
            // ===== MMA issuer of tile t: the whole warp runs the (uniform) control flow so that
            // descriptors live in uniform registers; one elected lane issues =====
            const int t = warp - T16_MMA_WARP0;
            if (!(solo && t == 1)) {
            tc::mbar_wait(bar_res, 0);
            const uint32_t tA = tbase + t * 256, tD = tA + 64;
            const uint32_t sG_hi = tc::smem_u32(sG), sG_lo = sG_hi + T16_IMG;
            const uint32_t sM_hi = tc::smem_u32(sM), sM_lo = sM_hi + T16_IMG;
            const uint32_t sT_hi = tc::smem_u32(sT), sT_lo = sT_hi + T16_TIMG;
            const uint32_t z_base = tc::smem_u32(zbuf) + t * 2 * T16_IMG;
            const uint32_t ring_base = tc::smem_u32(ring);
            const uint32_t idesc_c = tc::idesc_f16(128, mc), idesc_64 = tc::idesc_f16(128, 64);
            uint64_t* reqA = bar_reqA + t;
            uint64_t* reqB = bar_reqB + t;
            uint64_t* respA = bar_respA + t;
            uint64_t* respB = bar_respB + t;
            uint64_t* reqF = bar_reqF + t * 3;
            uint64_t* respF = bar_respF + t * 3;
            uint32_t pa = 0, pb = 0, pf = 0;              // pf: bit b = parity of reqF[b]
            const int nks = PAD ? (d >> 4) : T16_K / 16; // K steps that carry parameters (the rest is zero padding)
            unsigned n = 0, gch = 0;
            int total_it = 0;
            {
                T16Unit un;
                for (int uk = 0; t16_unit(q, iters, uk, un); uk++) total_it += un.it1 - un.it0;
            }
            for (int itg = 0; itg < total_it; itg++) {
                // Per coarse step three MMA groups with their own dependencies:
                //   G1  z @ (b T G_c^T) -> D[0:128)   needs the normals and the rows' residual pass of the
                //                                      previous step -- it runs WHILE the rows still update theta
                //   G2  theta @ (a G_c^T), accumulating on G1   needs the updated A_theta
                //   G3  xi = z @ T      -> D[128:192) needs the rows' previous xi read (same signal as G2)
                for (int j = 0; j < J; j++, n++) {
                    const int b = inj ? 0 : (int)(n & 1);
                    const unsigned use = inj ? n : (n >> 1);
                    const uint32_t z_hi = z_base + b * T16_IMG, z_lo = z_base + T16_IMG;
                    const bool dbg_on = q.dbg && blockIdx.x == 0 && t == 0 && lane == 0 && n < 40;
                    if (dbg_on) q.dbg[1 * 256 + 6 * n] = clock64();
                    tc::mbar_wait(reqA, pa); pa ^= 1;
                    if (dbg_on) q.dbg[1 * 256 + 6 * n + 1] = clock64();
                    tc::mbar_wait(bar_zfull + t * 2 + b, (uint32_t)(use & 1));
                    tc::fence_after_sync();
                    if (dbg_on) q.dbg[1 * 256 + 6 * n + 2] = clock64();
                    if (tc::elect_one()) t16_issue_z(tD, z_hi, z_lo, inj, sM_hi, sM_lo, idesc_c, 0, nks);
                    __syncwarp();
                    if (dbg_on) q.dbg[1 * 256 + 6 * n + 3] = clock64();
                    tc::mbar_wait(reqB, pb); pb ^= 1;
                    tc::fence_after_sync();
                    if (dbg_on) q.dbg[1 * 256 + 6 * n + 4] = clock64();
                    if (tc::elect_one()) {
                        t16_issue_theta(tD, tA, sG_hi, sG_lo, idesc_c, 1, nks);
                        tc::mma_commit(respA);
                        t16_issue_z(tD + 128, z_hi, z_lo, inj, sT_hi, sT_lo, idesc_64, 0, nks);
                        tc::mma_commit(respB);
                        tc::mma_commit(bar_zfree + t * 2 + b);
                    }
                    __syncwarp();
                    if (dbg_on) q.dbg[1 * 256 + 6 * n + 5] = clock64();
                }
                // fine stage: both chunk accumulators are free and A_theta is current once the rows
                // have signalled the end of the last coarse step
                tc::mbar_wait(reqA, pa); pa ^= 1;
                tc::mbar_wait(reqB, pb); pb ^= 1;
                // three 64-column accumulators (the xi columns are idle during the fine stage): the
                // MMAs run up to two chunks ahead of the rows' residual pass
                for (int c = 0, b = 0; c < NCH; c++, gch++, b = (b == 2 ? 0 : b + 1)) {
                    const int st = (int)(gch % T16_NST);
                    if (c >= 3) {
                        tc::mbar_wait(reqF + b, (pf >> b) & 1u);
                        pf ^= 1u << b;
                    }
                    tc::mbar_wait(bar_full + st, (uint32_t)((gch / T16_NST) & 1));
                    tc::fence_after_sync();
                    if (tc::elect_one()) {
                        const uint32_t b_hi = ring_base + st * T16_CHUNK_BYTES, b_lo = b_hi + T16_TIMG;
                        t16_issue_theta(tD + b * T16_CH, tA, b_hi, b_lo, idesc_64, 0, nks);
                        tc::mma_commit(respF + b);
                        tc::mma_commit(bar_empty + st);
                    }
                    __syncwarp();
                }
            }
            }
        }
        __syncwarp();
    } else {
        // =====================================================================================
        // row threads: two per chain (column halves)
        // =====================================================================================
        tc::setmaxnreg_inc<T16_REGS_ROW>();
        const int rw = warp - T16_ROW_WARP0;
        const int t = rw >> 3;                         // tile within the pair
        const int h = (rw >> 2) & 1;                   // column half
        const int wq = rw & 3;                         // TMEM lane quarter (= warp % 4)
        const int cl = wq * 32 + lane;                 // chain within the tile
        const int col0 = h * T16_HK;                   // first theta column of this thread
        const int nk = PAD ? min(T16_HK, max(0, d - col0)) : T16_HK;   // columns of this thread that exist (d <= 64)
        const uint32_t tA = tbase + ((uint32_t)(wq * 32) << 16) + t * 256;
        const uint32_t tD = tA + 64;
        uint64_t* reqA = bar_reqA + t;
        uint64_t* reqB = bar_reqB + t;
        uint64_t* respA = bar_respA + t;
        uint64_t* respB = bar_respB + t;
        uint64_t* reqF = bar_reqF + t * 3;
        uint64_t* respF = bar_respF + t * 3;
        uint32_t phA = 0, phB = 0, phF = 0;           // phF: bit b = parity of respF[b]
        int sbuf = 0;
        const LevelP<float>& l0 = p.lv[0];
        const LevelP<float>& l1 = p.lv[1];
        const float inv2vc = -0.5f / q.var_c, inv2vf = -0.5f / q.var_f;
        const float ca = q.ca, cxi = q.cxi, sc_c = q.sc_c, sc_f = q.sc_f, sc_p = q.sc_p;
        const unsigned long long sc_c2 = f2pack(sc_c, sc_c);
        const float th_scale = q.th_scale, th_unscale = q.th_unscale;
        const int ngc = mc >> 4, gc0 = h ? (ngc + 1) / 2 : 0, gc1 = h ? ngc : (ngc + 1) / 2;
        // One warp per tile waits on the MMA's mbarrier; the other seven block in a named barrier
        // (blocked warps take no issue slots, spinning mbarrier waiters do).
        const bool leader = (h == 0 && wq == 0);
        // Single-tile modes: the few active warps leave issue slots free, so every row warp waits on the mbarrier
        // itself -- one synchronisation hop less on the step's dependency chain (MMA -> rows -> MMA), which is what
        // paces a lone tile.  (No phase can be missed: the MMA behind the next completion of `bar` needs every row
        // warp's arrival after this one.)
        const bool direct_wait = solo && q.direct_wait;
        auto wait_mma = [&](uint64_t* bar, uint32_t& ph) {
            if (direct_wait) {
                tc::mbar_wait(bar, ph);
                ph ^= 1;
                tc::fence_after_sync();
                return;
            }
            if (leader) tc::mbar_wait(bar, ph);
            ph ^= 1;
            tc::named_bar_sync(3 + t, nbt);
            tc::fence_after_sync();
        };
        const bool idle = (solo && t == 1) || (half && wq >= 2);
        if (!idle) tc::mbar_wait(bar_res, 0);   // the data vector arrives with the resident operands

        T16Unit un;
        for (int uk = 0; !idle && t16_unit(q, iters, uk, un); uk++) {
            const int pair = un.pair;
            const int g = half ? pair * 64 + cl : solo ? pair * 128 + cl : pair * 256 + t * 128 + cl;     // chain slot (padded arrays)
            if (un.blk > 0) {
                // the previous block of this pair ran on another SM: wait for both of its tiles
                if (leader) while (t16_ld_acquire(q.progress + pair) < (solo ? 1 : 2) * un.blk) __nanosleep(200);
                tc::named_bar_sync(3 + t, nbt);
            }
            const long long gchain = p.chain_offset + g;
            const bool live = g < p.C;
            const size_t cs = (size_t)p.Cs;
            const size_t off0 = (size_t)col0 * cs + g;      // this thread's first column, this chain
            float th[T16_HK];                                // current coarse state, scaled by 2^s_theta
            {
                // (.cg loads: the state may have been written by another SM during this launch)
                const float* src = t16_opaque(l1.theta + off0);
#pragma unroll
                for (int k = 0; k < T16_HK; k++) th[k] = (k < nk) ? __ldcg(src + k * cs) * th_scale : 0.0f;
            }
            float like_c = __ldcg(l0.like + g), like_cs = like_c, like_f = __ldcg(l1.like + g), prior_f = __ldcg(l1.prior + g);
            long long ucur = __ldcg(p.ucur + g);
            int nacc_c = 0, nacc_f = 0;
            int acc_any = 0;

            auto store_A = [&]() {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; i++) tc::split_f16x2(th[2 * i], th[2 * i + 1], hi[i], lo[i]);
                tc::tmem_st16(tA + h * 16, hi);
                tc::tmem_st16(tA + 32 + h * 16, lo);
                tc::tmem_wait_st();
            };
            auto draw_u = [&]() -> float {
                if (inj) return (live && ucur < p.ulen) ? p.us[(size_t)g * p.ulen + ucur] : 0.5f;
                return philox_uniform<float>(p.seed, gchain, ucur);
            };

            store_A();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(reqA); tc::mbar_arrive(reqB); }

            for (int it = un.it0; it < un.it1; it++) {
                for (int j = 0; j < J; j++) {
                    // the accept-test uniform does not depend on the MMA: draw it while waiting
                    float u_mine = 0.0f;
                    if (h == 0) u_mine = draw_u();
                    ucur++;
                    const int dn = it * J + j;
                    const bool dbg_on = q.dbg && blockIdx.x == 0 && uk == 0 && t == 0 && lane == 0 && dn < 40 && (leader || (h == 1 && wq == 3));
                    long long* dbg = q.dbg + (leader ? 2 : 3) * 256 + 6 * dn;
                    if (dbg_on) dbg[0] = clock64();
                    wait_mma(respA, phA);
                    if (dbg_on) dbg[1] = clock64();
                    float ssq = 0.0f;
                    unsigned long long ssq2a = 0ull, ssq2b = 0ull;      // packed (even, odd) partial sums
                    int gc = gc0;
                    for (; gc + 1 < gc1; gc += 2) {
                        uint32_t v0[16], v1[16];
                        tc::tmem_ld16(tD + gc * 16, v0);
                        tc::tmem_ld16(tD + gc * 16 + 16, v1);
                        const float4* ny4 = reinterpret_cast<const float4*>(s_ny + gc * 16);
                        tc::tmem_wait_ld();
#pragma unroll
                        for (int i4 = 0; i4 < 4; i4++) {
                            const float4 y0 = ny4[i4], y1 = ny4[4 + i4];
                            const unsigned long long r0 = f2fma(f2pack(__uint_as_float(v0[4 * i4 + 0]), __uint_as_float(v0[4 * i4 + 1])), sc_c2, f2pack(y0.x, y0.y));
                            const unsigned long long r1 = f2fma(f2pack(__uint_as_float(v0[4 * i4 + 2]), __uint_as_float(v0[4 * i4 + 3])), sc_c2, f2pack(y0.z, y0.w));
                            const unsigned long long r2 = f2fma(f2pack(__uint_as_float(v1[4 * i4 + 0]), __uint_as_float(v1[4 * i4 + 1])), sc_c2, f2pack(y1.x, y1.y));
                            const unsigned long long r3 = f2fma(f2pack(__uint_as_float(v1[4 * i4 + 2]), __uint_as_float(v1[4 * i4 + 3])), sc_c2, f2pack(y1.z, y1.w));
                            ssq2a = f2fma(r0, r0, ssq2a); ssq2b = f2fma(r1, r1, ssq2b);
                            ssq2a = f2fma(r2, r2, ssq2a); ssq2b = f2fma(r3, r3, ssq2b);
                        }
                    }
                    if (gc < gc1) {
                        uint32_t v[16];
                        tc::tmem_ld16(tD + gc * 16, v);
                        tc::tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const float r = fmaf(__uint_as_float(v[i]), sc_c, s_ny[gc * 16 + i]);
                            ssq = fmaf(r, r, ssq);
                        }
                    }
                    t16_warp_arrive(reqA, lane);           // D[0:128) consumed: the next step's z products may start
                    if (dbg_on) dbg[2] = clock64();
                    {
                        float e0, e1, e2, e3;
                        f2unpack(ssq2a, e0, e1);
                        f2unpack(ssq2b, e2, e3);
                        ssq += (e0 + e1) + (e2 + e3);
                    }
                    float* sp = s_part + ((sbuf * 2 + t) * 2) * 128;
                    float* su = s_u + (sbuf * 2 + t) * 128;
                    sp[h * 128 + cl] = ssq;
                    if (h == 0) su[cl] = u_mine;
                    sbuf ^= 1;
                    tc::named_bar_sync(1 + t, nbt);
                    const float like_p = inv2vc * (sp[cl] + sp[128 + cl]);
                    const float u = su[cl];
                    const float alpha = isnan(like_p) ? 0.0f : expf(like_p - like_c);
                    const bool acc = u < alpha;
                    if (dbg_on) dbg[3] = clock64();
                    wait_mma(respB, phB);
                    if (dbg_on) dbg[4] = clock64();
                    if (__any_sync(0xffffffffu, acc)) {
                        uint32_t x0[16], x1[16];
                        tc::tmem_ld16(tD + 128 + col0, x0);
                        tc::tmem_ld16(tD + 128 + col0 + 16, x1);
                        tc::tmem_wait_ld();
                        if (acc) {
                            const unsigned long long ca2 = f2pack(ca, ca), cxi2 = f2pack(cxi, cxi);
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const unsigned long long a0 = f2mul(f2pack(th[2 * i], th[2 * i + 1]), ca2);
                                const unsigned long long a1 = f2mul(f2pack(th[16 + 2 * i], th[16 + 2 * i + 1]), ca2);
                                f2unpack(f2fma(f2pack(__uint_as_float(x0[2 * i]), __uint_as_float(x0[2 * i + 1])), cxi2, a0), th[2 * i], th[2 * i + 1]);
                                f2unpack(f2fma(f2pack(__uint_as_float(x1[2 * i]), __uint_as_float(x1[2 * i + 1])), cxi2, a1), th[16 + 2 * i], th[16 + 2 * i + 1]);
                            }
                        }
                        store_A();
                    }
                    if (acc) { like_c = like_p; acc_any = 1; nacc_c++; }
                    t16_warp_arrive(reqB, lane);           // A_theta current, xi consumed
                    if (dbg_on) dbg[5] = clock64();
                    // ---- coarse-level record (chain_coarse_i, sampler.py:421-436): the state after this
                    // step with its log-likelihood and accept flag; Link.prior / Link.model_output of
                    // these records are filled from the stored parameters when they are first fetched
                    // (engine: fill_lazy_history) ----
                    if (l0.store) {
                        const long long r0 = p.rec[0] + (long long)it * J + j;
                        if (r0 < l0.hist_cap) {
                            if (l0.store & TDA_STORE_THETA) {
                                float* dst = t16_opaque(l0.h_theta + (size_t)r0 * d * cs + off0);
#pragma unroll
                                for (int k = 0; k < T16_HK; k++)
                                    if (k < nk) __stcs(dst + k * cs, th[k] * th_unscale);
                            }
                            if (h == 0) {
                                if (l0.store & TDA_STORE_STATS) __stcs(l0.h_like + (size_t)r0 * cs + g, like_c);
                                if (l0.store & TDA_STORE_ACCEPT) l0.h_acc[(size_t)r0 * cs + g] = (uint8_t)acc;
                            }
                        }
                    }
                }
                // ---- fine level: F_f = theta @ G_f^T streamed in 64-column chunks, then theta @ LP ----
                float u2 = 0.0f;
                if (h == 0) u2 = draw_u();
                float ssq_f = 0.0f, ssq_p = 0.0f;
                for (int c = 0, b = 0; c < NCH; c++, b = (b == 2 ? 0 : b + 1)) {
                    {
                        uint32_t phb = (phF >> b) & 1u;
                        wait_mma(respF + b, phb);
                        phF ^= 1u << b;
                    }
                    uint32_t v0[16], v1[16];
                    tc::tmem_ld16(tD + b * T16_CH + col0, v0);
                    tc::tmem_ld16(tD + b * T16_CH + col0 + 16, v1);
                    const bool last = (c == NCH - 1);
                    const float4* ny4 = reinterpret_cast<const float4*>(last ? s_ny + T16_MAX_MC + col0
                                                                             : s_ny + T16_MAX_MC + T16_K + c * T16_CH + col0);
                    const float scl = last ? sc_p : sc_f;
                    tc::tmem_wait_ld();
                    if (c + 3 < NCH) t16_warp_arrive(reqF + b, lane);
                    unsigned long long a2 = 0ull, b2 = 0ull;
                    const unsigned long long scl2 = f2pack(scl, scl);
#pragma unroll
                    for (int i4 = 0; i4 < 4; i4++) {
                        const float4 y0 = ny4[i4], y1 = ny4[4 + i4];
                        const unsigned long long r0 = f2fma(f2pack(__uint_as_float(v0[4 * i4 + 0]), __uint_as_float(v0[4 * i4 + 1])), scl2, f2pack(y0.x, y0.y));
                        const unsigned long long r1 = f2fma(f2pack(__uint_as_float(v0[4 * i4 + 2]), __uint_as_float(v0[4 * i4 + 3])), scl2, f2pack(y0.z, y0.w));
                        const unsigned long long r2 = f2fma(f2pack(__uint_as_float(v1[4 * i4 + 0]), __uint_as_float(v1[4 * i4 + 1])), scl2, f2pack(y1.x, y1.y));
                        const unsigned long long r3 = f2fma(f2pack(__uint_as_float(v1[4 * i4 + 2]), __uint_as_float(v1[4 * i4 + 3])), scl2, f2pack(y1.z, y1.w));
                        a2 = f2fma(r0, r0, a2); b2 = f2fma(r1, r1, b2);
                        a2 = f2fma(r2, r2, a2); b2 = f2fma(r3, r3, b2);
                    }
                    float acc2;
                    {
                        float e0, e1, e2, e3;
                        f2unpack(a2, e0, e1);
                        f2unpack(b2, e2, e3);
                        acc2 = (e0 + e1) + (e2 + e3);
                    }
                    if (last) ssq_p += acc2;
                    else ssq_f += acc2;
                }
                float* sf = s_pf + (t * 2) * 256;
                sf[h * 256 + cl] = ssq_f;
                sf[h * 256 + 128 + cl] = ssq_p;
                float* su = s_u + (sbuf * 2 + t) * 128;
                if (h == 0) su[cl] = u2;
                sbuf ^= 1;
                tc::named_bar_sync(1 + t, nbt);
                const float like_fp = inv2vf * (sf[cl] + sf[256 + cl]);
                const float prior_p = -0.5f * (q.prior_logconst + (sf[128 + cl] + sf[256 + 128 + cl]));
                int accf = 0;
                if (acc_any) {
                    const float alpha2 = expf(like_fp - like_f + like_cs - like_c);
                    accf = (su[cl] < alpha2) ? 1 : 0;
                    ucur++;
                }
                if (accf) {
                    like_f = like_fp; prior_f = prior_p; like_cs = like_c; nacc_f++;
                    float* dst = t16_opaque(l1.theta + off0);
#pragma unroll
                    for (int k = 0; k < T16_HK; k++)
                        if (k < nk) dst[k * cs] = th[k] * th_unscale;
                } else {
                    like_c = like_cs;
                    const float* src = t16_opaque(l1.theta + off0);
#pragma unroll
                    for (int k = 0; k < T16_HK; k++)
                        if (k < nk) th[k] = __ldcg(src + k * cs) * th_scale;
                }
                acc_any = 0;
                // the A operand must hold the (possibly rewound) state before the next coarse job
                if (!__all_sync(0xffffffffu, accf)) store_A();
                if (it + 1 < un.it1) {
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) { tc::mbar_arrive(reqA); tc::mbar_arrive(reqB); }
                }
                // ---- fine-level record (coalesced: consecutive lanes = consecutive chains) ----
                const long long r = p.rec[1] + it;
                if (r < l1.hist_cap) {
                    if (l1.store & TDA_STORE_THETA) {
                        float* dst = t16_opaque(l1.h_theta + (size_t)r * d * cs + off0);
#pragma unroll
                        for (int k = 0; k < T16_HK; k++)
                            if (k < nk) dst[k * cs] = th[k] * th_unscale;
                    }
                    if (h == 0) {
                        if (l1.store & TDA_STORE_STATS) { l1.h_prior[(size_t)r * p.Cs + g] = prior_f; l1.h_like[(size_t)r * p.Cs + g] = like_f; }
                        if (l1.store & TDA_STORE_ACCEPT) l1.h_acc[(size_t)r * p.Cs + g] = (uint8_t)accf;
                    }
                }
                {
                    float* s1 = t16_opaque(p.sum1 + off0);
                    float* s2 = t16_opaque(p.sum2 + off0);
#pragma unroll
                    for (int k = 0; k < T16_HK; k++) {
                        const float x = th[k] * th_unscale;
                        if (k < nk) {
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(s1 + k * cs), "f"(x) : "memory");
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(s2 + k * cs), "f"(x * x) : "memory");
                        }
                    }
                }
            }
            // ---- write the chain state back (layout shared with the other kernels) ----
            {
                float* dst = t16_opaque(l0.theta + off0);
#pragma unroll
                for (int k = 0; k < T16_HK; k++)
                    if (k < nk) dst[k * cs] = th[k] * th_unscale;
            }
            if (h == 0) {
                l0.like[g] = like_c; l0.prior[g] = prior_f; l1.like[g] = like_f; l1.prior[g] = prior_f;
                l0.sv_like[1][g] = like_c; l0.sv_prior[1][g] = prior_f;
                l0.acc_sub[g] = 0;
                l0.n_acc[g] = __ldcg(l0.n_acc + g) + nacc_c;
                l1.n_acc[g] = __ldcg(l1.n_acc + g) + nacc_f;
                p.ucur[g] = ucur;
            }
            if (q.nb > 1) {
                // publish the block: state stores -> fence -> tile barrier -> release increment
                __threadfence();
                tc::named_bar_sync(3 + t, nbt);
                if (leader && lane == 0) t16_red_release_add(q.progress + pair, 1);
            }
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == T16_MMA_WARP0) tc::tmem_dealloc(tbase, 512);
}

template <typename R>
struct DaTc16State {
    std::string err;
    bool prepared = false;
    cudaStream_t fetch_stream = nullptr;   // engine's copy stream: prepare() reads the operands on it
    float theta_limit = 0.0f;
    bool eligible(const tda_config&, const Params<R>&) const { return false; }
    int prepare(const Params<R>&, const tda_config&) { err = "fp16-split tensor-core DA kernel is float32 only"; return 1; }
    int timeline(long long*) { return -5; }
    int run(Params<R>&, const tda_config&, long long, int, cudaStream_t) { err = "fp16-split tensor-core DA kernel is float32 only"; return -5; }
    void destroy() {}
};

template <>
struct DaTc16State<float> {
    std::string err;
    __half *dG = nullptr, *dM = nullptr, *dT = nullptr, *dF = nullptr;
    float* dNY = nullptr;
    long long* dDbg = nullptr;
    int* dProgress = nullptr;
    int progress_len = 0;
    bool prepared = false;
    cudaStream_t fetch_stream = nullptr;   // engine's copy stream: prepare() reads the operands on it
    float theta_limit = 0.0f;     // |theta| of a current state must stay below this (fp16 operand image at scale 2^s_theta)
    DaTc16Params q{};

    bool eligible(const tda_config& c, const Params<float>& P) const {
        if (c.dtype != TDA_F32 || c.n_levels != 2 || c.aem || c.randomize_subchain || c.mtm_k || c.prop_kind != TDA_PROP_PCN || c.adaptive) return false;
        // d, m_c and m_f are zero-padded to 64 / a multiple of 16 / a multiple of 64 (prepare)
        if (c.d > T16_K || c.d < 16 || (c.d % 16) != 0) return false;
        for (int l = 0; l < 2; l++)
            if (c.level[l].model_kind != TDA_MODEL_LINEAR || c.level[l].lik_kind > TDA_LIK_DENSE) return false;
        if (c.level[0].m < 1 || c.level[0].m > T16_MAX_MC) return false;
        if (c.level[1].m < 1 || c.level[1].m > T16_MAX_MF) return false;
        // Link.prior of the coarse records and Link.model_output of both levels are rebuilt from the
        // stored parameters when first fetched (engine: fill_lazy_history): they need the parameters
        if ((c.level[0].store & (TDA_STORE_STATS | TDA_STORE_OUTPUT)) && !(c.level[0].store & TDA_STORE_THETA)) return false;
        if ((c.level[1].store & TDA_STORE_OUTPUT) && !(c.level[1].store & TDA_STORE_THETA)) return false;
        if ((P.Cs % 256) != 0) return false;
        if (!(c.scaling > 0.0 && c.scaling < 1.0)) return false;
        return true;
    }

    // timeline probe: allocate the stamp buffer (subsequent runs fill it), or read it back
    int timeline(long long* host_out) {
        if (!dDbg) {
            if (cudaMalloc(&dDbg, 4 * 256 * sizeof(long long)) != cudaSuccess) return -3;
            cudaMemset(dDbg, 0, 4 * 256 * sizeof(long long));
            return 0;
        }
        cudaDeviceSynchronize();
        return cudaMemcpy(host_out, dDbg, 4 * 256 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
    }

    void destroy() {
        if (dDbg) cudaFree(dDbg);
        dDbg = nullptr;
        if (dProgress) cudaFree(dProgress);
        dProgress = nullptr;
        progress_len = 0;
        if (dG) cudaFree(dG);
        if (dM) cudaFree(dM);
        if (dT) cudaFree(dT);
        if (dF) cudaFree(dF);
        if (dNY) cudaFree(dNY);
        dG = dM = dT = dF = nullptr;
        dNY = nullptr;
        prepared = false;
    }

    static double maxabs(const std::vector<double>& W, int ldw, int n0, int rows, double factor) {
        double m = 0;
        for (int k = 0; k < T16_K; k++)
            for (int n = 0; n < rows; n++) m = std::fmax(m, std::fabs(factor * W[(size_t)k * ldw + n0 + n]));
        return m;
    }

    // returns 1 when the problem cannot be scaled into fp16 (caller falls back to another kernel)
    int prepare(const Params<float>& P, const tda_config& c) {
        const bool prof = std::getenv("TDA_PROFILE") != nullptr;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        const auto t_start = now();
        auto fetch = [&](const float* dev, size_t n, std::vector<double>& h) {
            std::vector<float> f(n);
            cudaError_t e = cudaMemcpy(f.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost);
            h.assign(f.begin(), f.end());
            return e;
        };
        // The constant operands were complete when the last upload returned (uploads end with a synchronisation),
        // so they are read on the engine's copy stream -- not behind the initial-Link kernel queued on the run
        // stream -- into one page-locked staging buffer kept for the life of the process, with one wait at the end.
        struct Pending { const float* dev; size_t n; std::vector<double>* h; };
        std::vector<Pending> pend;
        auto fetch_all = [&]() -> cudaError_t {
            static std::mutex mu;
            static float* stage = nullptr;
            static size_t stage_cap = 0;
            size_t total = 0;
            for (const Pending& x : pend) total += x.n;
            if (!fetch_stream || !total) {
                for (const Pending& x : pend) { cudaError_t e1 = fetch(x.dev, x.n, *x.h); if (e1 != cudaSuccess) return e1; }
                return cudaSuccess;
            }
            std::lock_guard<std::mutex> lk(mu);
            if (stage_cap < total) {
                if (stage) cudaFreeHost(stage);
                stage = nullptr; stage_cap = 0;
                cudaError_t e1 = cudaHostAlloc((void**)&stage, total * sizeof(float), cudaHostAllocPortable);
                if (e1 != cudaSuccess) return e1;
                stage_cap = total;
            }
            size_t off = 0;
            for (const Pending& x : pend) {
                cudaError_t e1 = cudaMemcpyAsync(stage + off, x.dev, x.n * sizeof(float), cudaMemcpyDeviceToHost, fetch_stream);
                if (e1 != cudaSuccess) return e1;
                off += x.n;
            }
            cudaError_t e1 = cudaStreamSynchronize(fetch_stream);
            if (e1 != cudaSuccess) return e1;
            off = 0;
            for (const Pending& x : pend) { x.h->assign(stage + off, stage + off + x.n); off += x.n; }
            return cudaSuccess;
        };
        // real shapes (d0, mc0, mf0) and the padded ones the kernel runs on: zero rows / columns add
        // nothing to a contraction, zero data under a zero operator column gives a zero residual
        const int d0 = c.d, mc0 = c.level[0].m, mf0 = c.level[1].m;
        const int mc = (mc0 + 15) / 16 * 16, mf = (mf0 + T16_CH - 1) / T16_CH * T16_CH;
        std::vector<double> T, LP, Ac, Af, bc, bf, dc, df, mu, sc;
        cudaError_t e = cudaSuccess;
        if (P.ldD < T16_K || P.lv[0].ldA < mc || P.lv[1].ldA < mf) { err = "tc16: operand leading dimensions"; return 1; }
        pend = {{P.T, (size_t)d0 * P.ldD, &T},        {P.LP, (size_t)d0 * P.ldD, &LP},
                {P.lv[0].A, (size_t)d0 * P.lv[0].ldA, &Ac}, {P.lv[1].A, (size_t)d0 * P.lv[1].ldA, &Af},
                {P.lv[0].b, (size_t)mc0, &bc},        {P.lv[1].b, (size_t)mf0, &bf},
                {P.lv[0].data, (size_t)mc0, &dc},     {P.lv[1].data, (size_t)mf0, &df},
                {P.prior_mean, (size_t)d0, &mu},      {P.scaling, (size_t)P.Cs, &sc}};
        e = fetch_all();
        if (e != cudaSuccess) { err = std::string("tc16 prepare: ") + cudaGetErrorString(e); return -2; }
        const auto t_fetched = now();
        auto pad = [&](std::vector<double>& W, int ld, int cols) {
            for (int k = 0; k < d0; k++)
                for (int n = cols; n < ld; n++) W[(size_t)k * ld + n] = 0.0;
            W.resize((size_t)T16_K * ld, 0.0);
        };
        pad(T, P.ldD, d0); pad(LP, P.ldD, d0); pad(Ac, P.lv[0].ldA, mc0); pad(Af, P.lv[1].ldA, mf0);
        mu.resize(T16_K, 0.0);
        // Diagonal and dense Gaussian likelihoods (distributions.py:304-315, :246-301) are folded into
        // the operators: with prec = L L^T (diagonal: L = diag(var^-1/2)) the log-likelihood
        // -0.5 r^T prec r is -0.5 |L^T r|^2 -- an isotropic unit-variance likelihood of the whitened
        // model L^T G against the whitened data L^T (y - b)
        std::vector<double> rc(mc0), rf(mf0);
        for (int j = 0; j < mc0; j++) rc[j] = dc[j] - bc[j];
        for (int j = 0; j < mf0; j++) rf[j] = df[j] - bf[j];
        double lik_var[2] = {c.level[0].lik_var, c.level[1].lik_var};
        for (int l = 0; l < 2; l++) {
            const int kind = c.level[l].lik_kind, m0 = l ? mf0 : mc0, ld = P.lv[l].ldA;
            std::vector<double>& A = l ? Af : Ac;
            std::vector<double>& r = l ? rf : rc;
            if (kind == TDA_LIK_ISO) continue;
            lik_var[l] = 1.0;
            if (kind == TDA_LIK_DIAG) {
                std::vector<double> var;
                e = fetch(P.lv[l].var, m0, var);
                if (e != cudaSuccess) { err = std::string("tc16 prepare: ") + cudaGetErrorString(e); return -2; }
                for (int n = 0; n < m0; n++) {
                    if (!(var[n] > 0.0)) { err = "tc16: non-positive likelihood variance"; return 1; }
                    const double w = 1.0 / std::sqrt(var[n]);
                    for (int k = 0; k < d0; k++) A[(size_t)k * ld + n] *= w;
                    r[n] *= w;
                }
            } else {
                std::vector<double> Lc;
                e = fetch(P.lv[l].prec, (size_t)m0 * m0, Lc);
                if (e != cudaSuccess) { err = std::string("tc16 prepare: ") + cudaGetErrorString(e); return -2; }
                for (int j = 0; j < m0; j++) {            // in-place lower Cholesky factor of the precision
                    double dj = Lc[(size_t)j * m0 + j];
                    for (int k = 0; k < j; k++) dj -= Lc[(size_t)j * m0 + k] * Lc[(size_t)j * m0 + k];
                    if (!(dj > 0.0)) { err = "tc16: likelihood precision is not positive definite in float32"; return 1; }
                    dj = std::sqrt(dj);
                    Lc[(size_t)j * m0 + j] = dj;
                    for (int i = j + 1; i < m0; i++) {
                        double v = Lc[(size_t)i * m0 + j];
                        for (int k = 0; k < j; k++) v -= Lc[(size_t)i * m0 + k] * Lc[(size_t)j * m0 + k];
                        Lc[(size_t)i * m0 + j] = v / dj;
                    }
                }
                std::vector<double> row(m0);
                for (int k = 0; k <= d0; k++) {            // rows of G^T, then the data residual: x <- x L
                    double* x = (k < d0) ? &A[(size_t)k * ld] : r.data();
                    for (int n = 0; n < m0; n++) {
                        double v = 0;
                        for (int j = n; j < m0; j++) v += x[j] * Lc[(size_t)j * m0 + n];
                        row[n] = v;
                    }
                    for (int n = 0; n < m0; n++) x[n] = row[n];
                }
            }
        }
        // the pCN step is folded into the operators: it must be the same for every chain
        const double beta = sc[0];
        for (int i = 0; i < P.C; i++)
            if (sc[i] != beta) { err = "tc16: per-chain pCN step sizes"; return 1; }
        if (!(beta > 0.0 && beta < 1.0)) { err = "tc16: pCN step outside (0, 1)"; return 1; }
        const double a = std::sqrt(1.0 - (double)(float)beta * (double)(float)beta);
        const int ldD = P.ldD, ldc = P.lv[0].ldA, ldf = P.lv[1].ldA;

        // composed coarse operator M[k][n] = sum_j T[k][j] * G_c^T[j][n]
        std::vector<double> M((size_t)T16_K * mc, 0.0);
        for (int k = 0; k < T16_K; k++)
            for (int j = 0; j < T16_K; j++) {         // same summation order over j for every (k, n)
                const double t = T[(size_t)k * ldD + j];
                const double* g = &Ac[(size_t)j * ldc];
                double* mrow = &M[(size_t)k * mc];
                for (int n = 0; n < mc; n++) mrow[n] += t * g[n];
            }
        // scales: theta from the prior (mean +- 12 sd must stay below 2^15), operators to [2^13, 2^14)
        double th_max = 0;
        for (int k = 0; k < T16_K; k++) {
            double var = 0;     // prior covariance diagonal = sum_j T[j][k]^2  (xi = z @ T ~ N(0, C))
            for (int j = 0; j < T16_K; j++) var += T[(size_t)j * ldD + k] * T[(size_t)j * ldD + k];
            th_max = std::fmax(th_max, std::fabs(mu[k]) + 12.0 * std::sqrt(var));
        }
        const int s_th = pow2_scale_for(th_max, 15);
        const int s_z = 12;
        int s_G = pow2_scale_for(maxabs(Ac, ldc, 0, mc, a), 14);
        int s_M = s_th + s_G - s_z;
        const double mM = maxabs(M, mc, 0, mc, beta);
        const int s_M_max = pow2_scale_for(mM, 15);
        if (s_M > s_M_max) { s_G -= (s_M - s_M_max); s_M = s_M_max; }
        const int s_T = pow2_scale_for(maxabs(T, ldD, 0, 64, 1.0), 14);
        const int s_Gf = pow2_scale_for(maxabs(Af, ldf, 0, mf, 1.0), 14);
        const int s_LP = pow2_scale_for(maxabs(LP, ldD, 0, 64, 1.0), 14);
        auto bad = [](int s) { return s < -20 || s > 40; };
        if (bad(s_th) || bad(s_G) || bad(s_M) || bad(s_T) || bad(s_Gf) || bad(s_LP)) { err = "tc16: operands do not fit the fp16 range"; return 1; }
        if (maxabs(Ac, ldc, 0, mc, a) * std::ldexp(1.0, s_G) < 64.0) { err = "tc16: coarse operator and composed operator differ too much in scale"; return 1; }

        const int nfc = mf / T16_CH, nch = nfc + 1;
        std::vector<__half> hG((size_t)2 * mc * T16_K), hM((size_t)2 * mc * T16_K), hT((size_t)2 * 64 * T16_K),
            hF((size_t)nch * 2 * T16_CH * T16_K);
        canon_split16(Ac, ldc, 0, mc, a, s_G, hG.data(), hG.data() + (size_t)mc * T16_K);
        canon_split16(M, mc, 0, mc, beta, s_M, hM.data(), hM.data() + (size_t)mc * T16_K);
        canon_split16(T, ldD, 0, 64, 1.0, s_T, hT.data(), hT.data() + (size_t)64 * T16_K);
        for (int ch = 0; ch < nfc; ch++)
            canon_split16(Af, ldf, ch * T16_CH, T16_CH, 1.0, s_Gf, hF.data() + (size_t)ch * 2 * T16_CH * T16_K,
                          hF.data() + (size_t)ch * 2 * T16_CH * T16_K + T16_CH * T16_K);
        canon_split16(LP, ldD, 0, T16_CH, 1.0, s_LP, hF.data() + (size_t)nfc * 2 * T16_CH * T16_K,
                      hF.data() + (size_t)nfc * 2 * T16_CH * T16_K + T16_CH * T16_K);
        const auto t_composed = now();
        destroy();
        if (e == cudaSuccess) e = cudaMalloc(&dG, hG.size() * 2);
        if (e == cudaSuccess) e = cudaMalloc(&dM, hM.size() * 2);
        if (e == cudaSuccess) e = cudaMalloc(&dT, hT.size() * 2);
        if (e == cudaSuccess) e = cudaMalloc(&dF, hF.size() * 2);
        if (e == cudaSuccess) e = cudaMemcpy(dG, hG.data(), hG.size() * 2, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dM, hM.data(), hM.size() * 2, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dT, hT.data(), hT.size() * 2, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dF, hF.data(), hF.size() * 2, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { err = std::string("tc16 prepare: ") + cudaGetErrorString(e); return -2; }
        std::vector<float> ny((size_t)T16_MAX_MC + T16_K + mf, 0.f);
        for (int j = 0; j < mc0; j++) ny[j] = -(float)rc[j];
        for (int n = 0; n < T16_K; n++) {
            double s = 0;
            for (int k = 0; k < T16_K; k++) s += mu[k] * LP[(size_t)k * ldD + n];
            ny[T16_MAX_MC + n] = -(float)s;
        }
        for (int j = 0; j < mf0; j++) ny[T16_MAX_MC + T16_K + j] = -(float)rf[j];
        e = cudaMalloc(&dNY, ny.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(dNY, ny.data(), ny.size() * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { err = std::string("tc16 prepare: ") + cudaGetErrorString(e); return -2; }
        q.G_hl = dG; q.M_hl = dM; q.T_hl = dT; q.F_chunks = dF; q.ny = dNY;
        q.mc = mc; q.mf = mf; q.n_chunks = nch; q.J = c.subchain[0];
        q.var_c = (float)lik_var[0]; q.var_f = (float)lik_var[1];
        q.prior_logconst = (float)c.prior_logconst;
        q.ca = (float)a;
        q.cxi = (float)(beta * std::ldexp(1.0, s_th - s_z - s_T));
        q.sc_c = (float)std::ldexp(1.0, -(s_th + s_G));
        q.sc_f = (float)std::ldexp(1.0, -(s_th + s_Gf));
        q.sc_p = (float)std::ldexp(1.0, -(s_th + s_LP));
        q.th_scale = (float)std::ldexp(1.0, s_th);
        q.th_unscale = (float)std::ldexp(1.0, -s_th);
        theta_limit = (float)std::ldexp(65000.0, -s_th);
        if (prof)
            std::fprintf(stderr, "[tc16 prepare] operands D2H %.2f ms, host composition %.2f ms, device images %.2f ms\n",
                         ms(t_start, t_fetched), ms(t_fetched, t_composed), ms(t_composed, now()));
        prepared = true;
        return 0;
    }

    int run(Params<float>& P, const tda_config& c, long long iterations, int sm_count, cudaStream_t st) {
        if (!prepared) { int r = prepare(P, c); if (r) return r; }
        cudaError_t e = cudaSuccess;
        q.n_pairs = P.Cs / 256;
        q.solo = 0; q.half = 0;
        if (2 * q.n_pairs <= sm_count && !getenv("TDA_TC16_NO_SOLO")) {
            q.solo = 1; q.n_pairs = P.Cs / 128;
            if (2 * q.n_pairs <= sm_count && !getenv("TDA_TC16_NO_HALF")) { q.half = 1; q.n_pairs = P.Cs / 64; }
        }
        q.dbg = dDbg;
        q.rk = philox_round_keys(P.seed);
        q.direct_wait = getenv("TDA_TC16_LEADER_WAIT") ? 0 : 1;
        const int grid = q.n_pairs < sm_count ? q.n_pairs : sm_count;
        // iteration blocks: the smallest block count (<= 16) whose round-robin deal of the
        // (pair, block) units fills at least 97 % of the last wave, else the best one
        int nb = 1;
        if (q.n_pairs > grid) {
            double best = 0.0;
            const int nb_max = iterations < 16 ? (int)iterations : 16;
            for (int cand = 1; cand <= nb_max; cand++) {
                const long long units = (long long)q.n_pairs * cand;
                const double eff = (double)units / (double)(((units + grid - 1) / grid) * grid);
                if (eff > best + 1e-9) { best = eff; nb = cand; }
                if (eff >= 0.97) { nb = cand; break; }
            }
        }
        if (const char* forced = getenv("TDA_TC16_BLOCKS")) {     // test hook: force the block count
            const int f = atoi(forced);
            if (f >= 1) nb = f < iterations ? f : (int)iterations;
        }
        q.ib = (int)((iterations + nb - 1) / nb);
        q.nb = (int)((iterations + q.ib - 1) / q.ib);
        if (q.nb > 1) {
            if (progress_len < q.n_pairs) {
                if (dProgress) cudaFree(dProgress);
                dProgress = nullptr;
                if (cudaMalloc(&dProgress, (size_t)q.n_pairs * sizeof(int)) != cudaSuccess) { err = "tc16: cudaMalloc progress"; return -3; }
                progress_len = q.n_pairs;
            }
            e = cudaMemsetAsync(dProgress, 0, (size_t)q.n_pairs * sizeof(int), st);
            if (e != cudaSuccess) { err = std::string("tc16 progress: ") + cudaGetErrorString(e); return -2; }
        }
        q.progress = dProgress;
        const size_t smem = T16_SMEM_BYTES;
        auto kern = (P.d == T16_K) ? da_tc16_kernel<false> : da_tc16_kernel<true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { err = std::string("tc16 attr: ") + cudaGetErrorString(e); return -2; }
        P.mode = MODE_RUN;
        P.iterations = iterations;
        P.z_round = 1;
        kern<<<grid, T16_THREADS, smem, st>>>(P, q);
        e = cudaGetLastError();
        if (e != cudaSuccess) { err = std::string("tc16 launch: ") + cudaGetErrorString(e); return -2; }
        return 0;
    }
};

}  // namespace tda
