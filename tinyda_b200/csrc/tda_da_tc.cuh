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

// ---------------------------------------------------------------------------------------------
// Tensor-core Delayed-Acceptance kernel.
//
// One CTA = 2 tiles x 128 chains.  A chain is a TMEM lane; TWO threads serve each chain (16 row
// warps: warp w -> tile w/8, column half (w/4)%2, TMEM lane quarter w%4), each owning half of
// the columns of every row-wise operation (proposal draws, theta', residual sums) so that four
// row warps per scheduler hide each other's latencies; the two partial sums of a chain meet in
// shared memory across one named barrier per step, and both threads then take the identical
// accept decision.  Warp 16 lane 0 issues every tcgen05.mma; warp 17 lane 0 runs the bulk-copy
// producer.  Per tile, TMEM holds the A operand (hi | lo, 2 x 64 columns) and 128 accumulator
// columns (one 128-column buffer for the coarse jobs, two 64-column buffers for the streamed
// fine-operator chunks so that the MMA of chunk c+1 overlaps the residual reduction of chunk c).
// All B operands live in shared memory in the canonical K-major no-swizzle layout, pre-split
// into hi / lo on the host: the proposal factor T and the coarse operator G_c^T stay resident,
// the fine operator [G_f^T | LP] streams through a ring of 64-column chunks filled by
// cp.async.bulk.  Row warps and the MMA thread hand jobs over through mbarriers (req: "A
// written / D buffer free", resp: tcgen05.commit); the MMA thread polls both tiles and serves
// whichever is ready, so one tile's CUDA-core work (Philox, Box-Muller, splits, residual
// reductions) overlaps the other tile's MMAs.
//
// Reference semantics are those of Tile::base_step / Tile::upper_step in tda_kernels.cuh
// (chain.py:325-444) specialised to pCN + isotropic likelihoods + linear models; the state
// buffers are shared with the generic kernel, so runs of the two kernels can be interleaved.
// ---------------------------------------------------------------------------------------------
constexpr int TC_K = 64;                    // parameters (contraction length)
constexpr int TC_CH = 64;                   // columns per streamed chunk
constexpr int TC_NST = 3;                   // ring stages
constexpr int TC_MAX_MC = 128;
constexpr int TC_MAX_MF = 4096;
constexpr int TC_ROW_WARPS = 16;
constexpr int TC_MMA_WARP = 16;
constexpr int TC_PROD_WARP = 17;
constexpr int TC_THREADS = 32 * 18;
constexpr int TC_HK = TC_K / 2;             // theta columns per thread
constexpr int TC_CHUNK_BYTES = TC_CH * TC_K * 4 * 2;   // hi + lo
inline size_t tc_smem_bytes(int rwmh, int nst) {
    return (size_t)(2 * 64 * TC_K + 2 * TC_MAX_MC * TC_K + (rwmh ? 2 * 64 * TC_K : 0)) * 4 + (size_t)nst * TC_CHUNK_BYTES +
           (size_t)(3 * 2 * 2 * 2 * 128) * 4 + 256;
}

__constant__ float c_yc[TC_MAX_MC];         // coarse data - offset
__constant__ float c_yf[TC_MAX_MF];         // fine data - offset
__constant__ float c_lp[TC_K];              // prior_mean @ LP

struct DaTcParams {
    const float* T_hl;       // canonical [hi 64x64 | lo 64x64]
    const float* Gc_hl;      // canonical [hi mc x 64 | lo mc x 64]
    const float* F_chunks;   // n_chunks x [hi 64x64 | lo 64x64]   (fine operator columns, then LP)
    int mc, mf, n_chunks, J;
    float var_c, var_f, prior_logconst;
    int n_pairs;
    // GaussianRandomWalk instead of pCN (proposal.py:247-258): theta' = theta + s xi, and the acceptance needs the
    // proposal's log-prior as well -- a third job per coarse step, theta' @ LP, against a resident copy of LP; the
    // ring then has two stages instead of three (shared memory)
    int rwmh, nst;
    const float* LP_hl;      // canonical [hi 64x64 | lo 64x64] (the last chunk of F_chunks)
};

__device__ __forceinline__ void tc_issue_job(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi_saddr, uint32_t b_lo_saddr, int N) {
    const uint32_t idesc = tc::idesc_tf32(128, N);
    uint32_t accumulate = 0;
#pragma unroll
    for (int pass = 0; pass < 3; pass++) {
        const uint32_t a = a_tmem + (pass == 1 ? TC_K : 0);
        const uint64_t b0 = tc::smem_desc_kmajor(pass == 2 ? b_lo_saddr : b_hi_saddr, 128, (TC_K / 4) * 128);
#pragma unroll
        for (int ks = 0; ks < TC_K / 8; ks++) {
            tc::mma_tf32_ts(d_tmem, a + ks * 8, b0 + (uint64_t)(ks * 16), idesc, accumulate);
            accumulate = 1;
        }
    }
}

// split 8 fp32 values into tf32 hi / lo and store them as A-operand columns [col, col+8)
__device__ __forceinline__ void tc_store_split8(uint32_t tA, int col, const float* x) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = __float_as_uint(tc::tf32_hi(x[i]));
    tc::tmem_st8(tA + col, w);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = __float_as_uint(x[i] - __uint_as_float(w[i]));
    tc::tmem_st8(tA + TC_K + col, w);
}

// hides a pointer's provenance from the optimiser so that the 32 per-column addresses derived
// from it are recomputed at the use site instead of being hoisted out of the run loop (and spilled)
template <typename T>
__device__ __forceinline__ T* tc_opaque(T* ptr) {
    asm volatile("" : "+l"(ptr));
    return ptr;
}

// row warps: "my TMEM writes / reads are done" -> one arrival per warp
__device__ __forceinline__ void tc_warp_arrive(uint64_t* bar, int lane) {
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(bar);
}

// RW = false: pCN (two jobs per coarse step, three ring stages); RW = true: random walk (see DaTcParams::rwmh)
template <bool RW>
__global__ void __launch_bounds__(TC_THREADS, 1)
da_tc_kernel(const __grid_constant__ Params<float> p, const __grid_constant__ DaTcParams q) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sT = reinterpret_cast<float*>(smem);                         // 32 KB (hi | lo)
    float* sGc = sT + 2 * 64 * TC_K;                                    // 64 KB (hi | lo), sized for mc = 128
    float* sLP = sGc + 2 * TC_MAX_MC * TC_K;                            // 32 KB (hi | lo), random-walk variant only
    constexpr int NST = RW ? 2 : TC_NST;
    unsigned char* ring = reinterpret_cast<unsigned char*>(sLP + (RW ? 2 * 64 * TC_K : 0));   // NST x 32 KB
    float* s_part = reinterpret_cast<float*>(ring + NST * TC_CHUNK_BYTES);      // [buf 2][tile 2][half 2][128]
    float* s_pf = s_part + 2 * 2 * 2 * 128;                                      // [tile 2][half 2][val 2][128]
    float* s_part2 = s_pf + 2 * 2 * 2 * 128;                                     // [buf 2][tile 2][half 2][128] prior partials
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_part2 + 2 * 2 * 2 * 128);
    uint64_t* bar_res = bars;            // resident operands landed
    uint64_t* bar_req = bars + 1;        // [tile 2][buffer 2]
    uint64_t* bar_resp = bars + 5;       // [tile 2][buffer 2]
    uint64_t* bar_full = bars + 9;       // [NST]
    uint64_t* bar_empty = bars + 9 + TC_NST;   // [NST]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 9 + 2 * TC_NST);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == TC_MMA_WARP) tc::tmem_alloc(s_tmem, 512);
    if (tid == TC_PROD_WARP * 32) {
        tc::mbar_init(bar_res, 1);
        for (int i = 0; i < 4; i++) { tc::mbar_init(bar_req + i, 8); tc::mbar_init(bar_resp + i, 1); }
        for (int s = 0; s < NST; s++) { tc::mbar_init(bar_full + s, 1); tc::mbar_init(bar_empty + s, 2); }
        tc::fence_mbar_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *s_tmem;

    const int J = q.J, mc = q.mc, NCH = q.n_chunks;
    const int my_pairs = (q.n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const long long iters = p.iterations;

    if (warp == TC_PROD_WARP) {
        // ===== producer: resident operands once, then the fine-operator chunk ring =====
        if (lane == 0) {
            const long long total_chunks = (long long)my_pairs * iters * NCH;
            const uint32_t bT = 2u * 64 * TC_K * 4, bG = 2u * (uint32_t)mc * TC_K * 4;
            tc::mbar_expect_tx(bar_res, bT + bG + (RW ? bT : 0u));
            tc::bulk_g2s(sT, q.T_hl, bT, bar_res);
            if (RW) tc::bulk_g2s(sLP, q.LP_hl, bT, bar_res);
            // hi part and lo part of G_c are stored back to back in global; in shared memory the lo
            // part starts at the fixed offset used by the MMA warp (mc rows each)
            tc::bulk_g2s(sGc, q.Gc_hl, bG / 2, bar_res);
            tc::bulk_g2s(sGc + TC_MAX_MC * TC_K, q.Gc_hl + (size_t)mc * TC_K, bG / 2, bar_res);
            for (long long g = 0; g < total_chunks; g++) {
                const int st = (int)(g % NST);
                if (g >= NST) tc::mbar_wait(bar_empty + st, (uint32_t)(((g / NST) - 1) & 1));
                const int c = (int)(g % NCH);
                tc::mbar_expect_tx(bar_full + st, TC_CHUNK_BYTES);
                tc::bulk_g2s(ring + (size_t)st * TC_CHUNK_BYTES, q.F_chunks + (size_t)c * (TC_CHUNK_BYTES / 4), TC_CHUNK_BYTES, bar_full + st);
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // ===== MMA issuer: serves whichever tile has its next job ready =====
        if (lane == 0) {
            tc::mbar_wait(bar_res, 0);
            const uint32_t sT_hi = tc::smem_u32(sT), sT_lo = sT_hi + 64 * TC_K * 4;
            const uint32_t sG_hi = tc::smem_u32(sGc), sG_lo = sG_hi + TC_MAX_MC * TC_K * 4;
            const uint32_t sL_hi = tc::smem_u32(sLP), sL_lo = sL_hi + 64 * TC_K * 4;
            constexpr int CJ = RW ? 3 : 2;                  // jobs per coarse step
            const int JOBS = CJ * J + NCH;
            const long long total = (long long)my_pairs * iters * JOBS;
            long long n0 = 0, n1 = 0, g0 = 0, g1 = 0;      // jobs done / chunks consumed per tile
            int q0 = 0, q1 = 0;                             // position inside the iteration
            uint32_t rph = 0;                               // bit (t*2+b): parity of req[t][b]
            while (n0 < total || n1 < total) {
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    long long& n = t ? n1 : n0;
                    long long& gch = t ? g1 : g0;
                    int& qq = t ? q1 : q0;
                    if (n >= total) continue;
                    const uint32_t tA = tbase + t * 256, tD = tA + 128;
                    if (qq < CJ * J) {
                        uint64_t* rq = bar_req + t * 2;
                        if (!tc::mbar_test(rq, (rph >> (t * 2)) & 1u)) continue;
                        rph ^= 1u << (t * 2);
                        tc::fence_after_sync();
                        const int kind = qq % CJ;
                        if (kind == 1) tc_issue_job(tD, tA, sG_hi, sG_lo, mc);          // F_c = theta' @ G_c^T
                        else if (kind == 0) tc_issue_job(tD, tA, sT_hi, sT_lo, 64);     // xi = z @ T
                        else tc_issue_job(tD, tA, sL_hi, sL_lo, 64);                    // (theta') @ LP (random walk)
                        tc::mma_commit(bar_resp + t * 2);
                    } else {
                        const int c = qq - CJ * J, b = c & 1;
                        const int st = (int)(gch % NST);
                        uint64_t* rq = bar_req + t * 2 + b;
                        if (!tc::mbar_test(rq, (rph >> (t * 2 + b)) & 1u)) continue;
                        if (!tc::mbar_test(bar_full + st, (uint32_t)((gch / NST) & 1))) continue;
                        rph ^= 1u << (t * 2 + b);
                        tc::fence_after_sync();
                        const uint32_t b_hi = tc::smem_u32(ring + (size_t)st * TC_CHUNK_BYTES), b_lo = b_hi + TC_CH * TC_K * 4;
                        tc_issue_job(tD + b * TC_CH, tA, b_hi, b_lo, TC_CH);
                        tc::mma_commit(bar_resp + t * 2 + b);
                        tc::mma_commit(bar_empty + st);
                        gch++;
                    }
                    n++;
                    qq = (qq + 1 == JOBS) ? 0 : qq + 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ===== row threads: two per chain =====
        const int t = warp >> 3;                       // tile within the pair
        const int h = (warp >> 2) & 1;                 // column half
        const int wq = warp & 3;                       // TMEM lane quarter
        const int cl = wq * 32 + lane;                 // chain within the tile
        const int col0 = h * TC_HK;                    // first theta column of this thread
        const uint32_t tA = tbase + ((uint32_t)(wq * 32) << 16) + t * 256;
        const uint32_t tD = tA + 128;
        uint64_t* req = bar_req + t * 2;
        uint64_t* resp = bar_resp + t * 2;
        uint32_t ph0 = 0, ph1 = 0;
        int sbuf = 0;
        const LevelP<float>& l0 = p.lv[0];
        const LevelP<float>& l1 = p.lv[1];
        const float inv2vc = -0.5f / q.var_c, inv2vf = -0.5f / q.var_f;
        // coarse residual columns of this thread: 16-column groups [gc0, gc1)
        const int ngc = mc >> 4, gc0 = h ? (ngc + 1) / 2 : 0, gc1 = h ? ngc : (ngc + 1) / 2;
        const bool inj = p.rng_mode == TDA_RNG_INJECTED;

        for (int pr = 0; pr < my_pairs; pr++) {
            const int pair = (int)blockIdx.x + pr * (int)gridDim.x;
            const int g = pair * 256 + t * 128 + cl;                   // chain slot (padded arrays)
            const long long gchain = p.chain_offset + g;
            const bool live = g < p.C;
            // d < 64 (16, 32, 48): the operators are zero-padded to 64 rows / columns on the host; this thread owns the
            // nk parameter columns [col0, col0 + nk) that exist, the padded ones stay zero (their normals are not drawn)
            const int d = p.d;
            const int nk = max(0, min(TC_HK, d - col0));
            float th[TC_HK];
#pragma unroll
            for (int k = 0; k < TC_HK; k++) th[k] = (k < nk) ? l1.theta[(size_t)(col0 + k) * p.Cs + g] : 0.0f;
            const size_t cs = (size_t)p.Cs;
            const size_t off0 = (size_t)col0 * cs + g;      // this thread's first column, this chain
            float like_c = l0.like[g], like_cs = like_c, like_f = l1.like[g], prior_f = l1.prior[g];
            float prior_c = l0.prior[g], prior_cs = prior_c;        // random walk: log-prior of the coarse state / of the subchain start
            long long ucur = p.ucur[g];
            int nacc_c = 0, nacc_f = 0;
            int acc_any = 0;
            float ca, cb;
            {
                const float s = __ldcg(p.scaling + g);
                ca = RW ? 1.0f : sqrtf(1.0f - s * s);
                cb = s;
            }
            long long tb = p.t_base;
            // adaptive global scaling (proposal.py:228-245): the `accepted` window of the last `period` entries -- coarse
            // decisions and the alignment entry of every fine iteration -- as a ring in global memory shared with the
            // lock-step kernel; kept by the h == 0 thread of the chain (same scheme as the whitened-state kernel)
            const bool adaptive = p.adaptive != 0;
            int wsum = (adaptive && h == 0) ? __ldcg(p.win_sum + g) : 0;
            long long wc = p.wcount;
            int wpos = adaptive ? (int)(wc % p.period) : 0;
            auto window_append = [&](int a) {
                if (h == 0) {
                    uint8_t* slot = p.win + (size_t)wpos * cs + g;
                    const int old = (wc >= p.period) ? (int)__ldcg(slot) : 0;
                    *slot = (uint8_t)a;
                    wsum += a - old;
                }
                wc++;
                wpos = (wpos + 1 == p.period) ? 0 : wpos + 1;
            };

            for (long long it = 0; it < iters; it++) {
                for (int j = 0; j < J; j++) {
                    // ---- proposal draws (this thread's 32 of the 64 normals) -> A (hi | lo) ----
                    const long long z0 = tb * d + col0;
#pragma unroll
                    for (int c0 = 0; c0 < TC_HK; c0 += 8) {
                        float v[8];
#pragma unroll
                        for (int b4 = 0; b4 < 2; b4++) {
                            if (c0 >= nk) {                         // padded parameter columns (d is a multiple of 16)
#pragma unroll
                                for (int i = 0; i < 4; i++) v[b4 * 4 + i] = 0.0f;
                            } else if (inj) {
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    long long idx = z0 + c0 + b4 * 4 + i;
                                    v[b4 * 4 + i] = (live && idx < p.zlen) ? p.zs[(size_t)g * p.zlen + idx] : 0.0f;
                                }
                            } else {
                                normals4<float>(philox_block(p.seed, gchain, STREAM_Z, (unsigned long long)(z0 >> 2) + (c0 >> 2) + b4), v + b4 * 4);
                            }
                        }
                        tc_store_split8(tA, col0 + c0, v);
                    }
                    tc::tmem_wait_st();
                    tc_warp_arrive(req, lane);
                    // ---- xi -> theta' -> A ----
                    tc::mbar_wait(resp, ph0); ph0 ^= 1;
                    tc::fence_after_sync();
#pragma unroll
                    for (int c0 = 0; c0 < TC_HK; c0 += 8) {
                        uint32_t v[8];
                        float x[8];
                        tc::tmem_ld8(tD + col0 + c0, v);
                        tc::tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 8; i++) x[i] = ca * th[c0 + i] + cb * __uint_as_float(v[i]);
                        tc_store_split8(tA, col0 + c0, x);
                    }
                    tc::tmem_wait_st();
                    tc_warp_arrive(req, lane);
                    // ---- coarse residual (this thread's column groups), accept / reject ----
                    tc::mbar_wait(resp, ph0); ph0 ^= 1;
                    tc::fence_after_sync();
                    float ssq = 0.0f;
                    for (int gc = gc0; gc < gc1; gc++) {
                        uint32_t v[16];
                        tc::tmem_ld16(tD + gc * 16, v);
                        tc::tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            float r = __uint_as_float(v[i]) - c_yc[gc * 16 + i];
                            ssq = fmaf(r, r, ssq);
                        }
                    }
                    float ssq_pc = 0.0f;
                    if (RW) {
                        // third job: (theta') @ LP into the accumulator the residual pass has just released
                        tc_warp_arrive(req, lane);
                        tc::mbar_wait(resp, ph0); ph0 ^= 1;
                        tc::fence_after_sync();
                        uint32_t v0[16], v1[16];
                        tc::tmem_ld16(tD + col0, v0);
                        tc::tmem_ld16(tD + col0 + 16, v1);
                        tc::tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const float w0 = __uint_as_float(v0[i]) - c_lp[col0 + i];
                            const float w1 = __uint_as_float(v1[i]) - c_lp[col0 + 16 + i];
                            ssq_pc = fmaf(w0, w0, ssq_pc);
                            ssq_pc = fmaf(w1, w1, ssq_pc);
                        }
                    }
                    float* sp = s_part + ((sbuf * 2 + t) * 2) * 128;
                    float* sp2 = s_part2 + ((sbuf * 2 + t) * 2) * 128;
                    sp[h * 128 + cl] = ssq;
                    if (RW) sp2[h * 128 + cl] = ssq_pc;
                    sbuf ^= 1;
                    tc::named_bar_sync(1 + t, 256);
                    const float like_p = inv2vc * (sp[cl] + sp[128 + cl]);
                    float prior_p = 0.0f, alpha;
                    if (RW) {
                        // chain.py:105-117 with proposal.py:253-258: the full posterior ratio
                        prior_p = -0.5f * (q.prior_logconst + (sp2[cl] + sp2[128 + cl]));
                        alpha = isnan(prior_p + like_p) ? 0.0f : expf((prior_p + like_p) - (prior_c + like_c));
                    } else {
                        alpha = isnan(like_p) ? 0.0f : expf(like_p - like_c);
                    }
                    float u;
                    if (inj) u = (live && ucur < p.ulen) ? p.us[(size_t)g * p.ulen + ucur] : 0.5f;
                    else u = philox_uniform<float>(p.seed, gchain, ucur);
                    ucur++;
                    const bool acc = u < alpha;
                    if (__any_sync(0xffffffffu, acc)) {
                        // theta' = hi + lo exactly: read it back from the A operand
#pragma unroll
                        for (int c0 = 0; c0 < TC_HK; c0 += 8) {
                            uint32_t vh[8], vl[8];
                            tc::tmem_ld8(tA + col0 + c0, vh);
                            tc::tmem_ld8(tA + TC_K + col0 + c0, vl);
                            tc::tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 8; i++)
                                th[c0 + i] = acc ? __uint_as_float(vh[i]) + __uint_as_float(vl[i]) : th[c0 + i];
                        }
                    }
                    if (acc) { like_c = like_p; prior_c = prior_p; acc_any = 1; nacc_c++; }
                    tb++;
                    // ---- coarse-level record (chain_coarse_i, sampler.py:421-436): the state after this step, its
                    // log-likelihood and accept flag; Link.prior / Link.model_output of these records are filled from
                    // the stored parameters when they are first fetched (engine: fill_lazy_history) ----
                    if (l0.store) {
                        const long long r0 = p.rec[0] + it * J + j;
                        if (r0 < l0.hist_cap) {
                            if (l0.store & TDA_STORE_THETA) {
                                float* dst = tc_opaque(l0.h_theta + (size_t)r0 * d * cs + off0);
#pragma unroll
                                for (int k = 0; k < TC_HK; k++)
                                    if (k < nk) __stcs(dst + k * cs, th[k]);
                            }
                            if (h == 0) {
                                if (l0.store & TDA_STORE_STATS) __stcs(l0.h_like + (size_t)r0 * cs + g, like_c);
                                if (l0.store & TDA_STORE_ACCEPT) l0.h_acc[(size_t)r0 * cs + g] = (uint8_t)acc;
                            }
                        }
                    }
                    if (adaptive) {
                        window_append(acc ? 1 : 0);
                        if ((tb % p.period) == 0) {                   // tb: proposal.t after this step
                            if (h == 0) {
                                const long long kk = tb / p.period - 1;
                                const float rate = (float)wsum / (float)p.period;
                                const float s_old = __ldcg(p.scaling + g);
                                p.scaling[g] = expf(logf(s_old) + powf(p.gamma, (float)(-(double)kk)) * (rate - p.alpha_star));
                            }
                            tc::named_bar_sync(1 + t, 256);          // the other half-thread of the chain reads the new step
                            const float s_new = __ldcg(p.scaling + g);
                            ca = RW ? 1.0f : sqrtf(1.0f - s_new * s_new);
                            cb = s_new;
                        }
                    }
                }
                // ---- fine level: A <- current coarse state ----
#pragma unroll
                for (int c0 = 0; c0 < TC_HK; c0 += 8) tc_store_split8(tA, col0 + c0, th + c0);
                tc::tmem_wait_st();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) { tc::mbar_arrive(req); tc::mbar_arrive(req + 1); }
                float ssq_f = 0.0f, ssq_p = 0.0f;
                for (int c = 0; c < NCH; c++) {
                    const int b = c & 1;
                    if (b) { tc::mbar_wait(resp + 1, ph1); ph1 ^= 1; }
                    else { tc::mbar_wait(resp, ph0); ph0 ^= 1; }
                    tc::fence_after_sync();
                    uint32_t v0[16], v1[16];
                    tc::tmem_ld16(tD + b * TC_CH + col0, v0);
                    tc::tmem_ld16(tD + b * TC_CH + col0 + 16, v1);
                    tc::tmem_wait_ld();
                    if (c + 2 < NCH) tc_warp_arrive(req + b, lane);
                    if (c < NCH - 1) {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            float r0 = __uint_as_float(v0[i]) - c_yf[c * TC_CH + col0 + i];
                            float r1 = __uint_as_float(v1[i]) - c_yf[c * TC_CH + col0 + 16 + i];
                            ssq_f = fmaf(r0, r0, ssq_f);
                            ssq_f = fmaf(r1, r1, ssq_f);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            float w0 = __uint_as_float(v0[i]) - c_lp[col0 + i];
                            float w1 = __uint_as_float(v1[i]) - c_lp[col0 + 16 + i];
                            ssq_p = fmaf(w0, w0, ssq_p);
                            ssq_p = fmaf(w1, w1, ssq_p);
                        }
                    }
                }
                float* sf = s_pf + (t * 2) * 256;
                sf[h * 256 + cl] = ssq_f;
                sf[h * 256 + 128 + cl] = ssq_p;
                tc::named_bar_sync(1 + t, 256);
                const float like_fp = inv2vf * (sf[cl] + sf[256 + cl]);
                const float prior_p = -0.5f * (q.prior_logconst + (sf[128 + cl] + sf[256 + 128 + cl]));
                int accf = 0;
                if (acc_any) {
                    const float alpha2 = expf(like_fp - like_f + like_cs - like_c);
                    float u;
                    if (inj) u = (live && ucur < p.ulen) ? p.us[(size_t)g * p.ulen + ucur] : 0.5f;
                    else u = philox_uniform<float>(p.seed, gchain, ucur);
                    ucur++;
                    accf = (u < alpha2) ? 1 : 0;
                }
                if (accf) {
                    like_f = like_fp; prior_f = prior_p; like_cs = like_c; prior_cs = prior_c; nacc_f++;
                    float* dst = tc_opaque(l1.theta + off0);
#pragma unroll
                    for (int k = 0; k < TC_HK; k++)
                        if (k < nk) dst[k * cs] = th[k];
                } else {
                    like_c = like_cs;
                    prior_c = prior_cs;
                    const float* src = tc_opaque(l1.theta + off0);
#pragma unroll
                    for (int k = 0; k < TC_HK; k++)
                        if (k < nk) th[k] = src[k * cs];
                }
                acc_any = 0;
                if (adaptive) window_append(accf);                   // the alignment entry (chain.py:391, :397)
                // ---- fine-level record (coalesced: consecutive lanes = consecutive chains) ----
                const long long r = p.rec[1] + it;
                if (r < l1.hist_cap) {
                    if (l1.store & TDA_STORE_THETA) {
                        float* dst = tc_opaque(l1.h_theta + (size_t)r * d * cs + off0);
#pragma unroll
                        for (int k = 0; k < TC_HK; k++)
                            if (k < nk) dst[k * cs] = th[k];
                    }
                    if (h == 0) {
                        if (l1.store & TDA_STORE_STATS) { l1.h_prior[(size_t)r * p.Cs + g] = prior_f; l1.h_like[(size_t)r * p.Cs + g] = like_f; }
                        if (l1.store & TDA_STORE_ACCEPT) l1.h_acc[(size_t)r * p.Cs + g] = (uint8_t)accf;
                    }
                }
                {
                    float* s1 = tc_opaque(p.sum1 + off0);
                    float* s2 = tc_opaque(p.sum2 + off0);
#pragma unroll
                    for (int k = 0; k < TC_HK; k++) {
                        // fire-and-forget reductions (one adder per address: deterministic), no load latency
                        if (k < nk) {
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(s1 + k * cs), "f"(th[k]) : "memory");
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(s2 + k * cs), "f"(th[k] * th[k]) : "memory");
                        }
                    }
                }
            }
            // ---- write the chain state back (layout shared with the generic kernel) ----
            {
                float* dst = tc_opaque(l0.theta + off0);
#pragma unroll
                for (int k = 0; k < TC_HK; k++)
                    if (k < nk) dst[k * cs] = th[k];
            }
            if (h == 0) {
                l0.like[g] = like_c; l0.prior[g] = prior_f; l1.like[g] = like_f; l1.prior[g] = prior_f;
                l0.sv_like[1][g] = like_c; l0.sv_prior[1][g] = prior_f;
                l0.acc_sub[g] = 0;
                l0.n_acc[g] += nacc_c; l1.n_acc[g] += nacc_f;
                p.ucur[g] = ucur;
                if (adaptive) p.win_sum[g] = wsum;
            }
            // both halves must have read the pair's scalars before anyone starts the next pair's
            // loads only matters for p.ucur / like arrays of THIS pair, which are per chain: no hazard
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == TC_MMA_WARP) tc::tmem_dealloc(tbase, 512);
}

template <typename R>
struct DaTcState {
    std::string err;
    bool eligible(const tda_config&, const Params<R>&) const { return false; }
    int run(Params<R>&, const tda_config&, long long, int, cudaStream_t) { err = "tensor-core DA kernel is float32 only"; return -5; }
    void destroy() {}
};

template <>
struct DaTcState<float> {
    std::string err;
    float *dT = nullptr, *dGc = nullptr, *dF = nullptr;
    bool prepared = false;
    DaTcParams q{};
    std::vector<float> yc, yf, clp;

    bool eligible(const tda_config& c, const Params<float>& P) const {
        if (c.dtype != TDA_F32 || c.n_levels != 2 || c.aem || c.randomize_subchain || c.mtm_k) return false;
        if (c.adaptive && (c.period < 1 || !P.win || !P.win_sum)) return false;
        if (c.prop_kind != TDA_PROP_PCN && c.prop_kind != TDA_PROP_RWMH) return false;
        if (c.d != 16 && c.d != 32 && c.d != 48 && c.d != TC_K) return false;      // zero-padded to 64 on the host
        for (int l = 0; l < 2; l++) {
            if (c.level[l].model_kind != TDA_MODEL_LINEAR) return false;
            // diagonal and dense Gaussian likelihoods are folded into the operators on the host (prepare)
            const int lk = c.level[l].lik_kind;
            if (lk != TDA_LIK_ISO && lk != TDA_LIK_DIAG && lk != TDA_LIK_DENSE) return false;
            if (lk == TDA_LIK_DENSE && c.level[l].m > 1920) return false;      // host Cholesky of the precision
        }
        // observation counts are zero-padded on the host to multiples of 16 (coarse) / 64 (fine): a zero operator
        // column against a zero datum adds nothing to a residual sum
        if (c.level[0].m < 1 || c.level[0].m > TC_MAX_MC) return false;
        if (c.level[1].m < 1 || c.level[1].m > TC_MAX_MF) return false;
        // Link.prior (coarse) and Link.model_output (both levels) are not kept by the kernel: rebuilt from the stored
        // parameters when first fetched (engine: fill_lazy_history), so they need the parameters
        if ((c.level[0].store & (TDA_STORE_STATS | TDA_STORE_OUTPUT)) && !(c.level[0].store & TDA_STORE_THETA)) return false;
        if ((c.level[1].store & TDA_STORE_OUTPUT) && !(c.level[1].store & TDA_STORE_THETA)) return false;
        if ((P.Cs % 256) != 0) return false;
        return true;
    }

    void destroy() {
        if (dT) cudaFree(dT);
        if (dGc) cudaFree(dGc);
        if (dF) cudaFree(dF);
        dT = dGc = dF = nullptr;
    }

    // canonical K-major hi / lo images of B[n][k] = W[k][n] (W row-major [K][ldw])
    static void canon_split(const std::vector<float>& W, int ldw, int n0, int rows, float* hi, float* lo) {
        for (int n = 0; n < rows; n++)
            for (int k = 0; k < TC_K; k++) {
                float x = W[(size_t)k * ldw + n0 + n];
                float h = tc::tf32_hi(x);
                size_t o = tc::canon_offset_f32(n, k, TC_K) / 4;
                hi[o] = h;
                lo[o] = x - h;
            }
    }

    int prepare(const Params<float>& P, const tda_config& c) {
        auto fetch = [&](const float* dev, size_t n, std::vector<float>& h) {
            h.resize(n);
            return cudaMemcpy(h.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost);
        };
        const int mc0 = c.level[0].m, mf0 = c.level[1].m;
        const int mc = (mc0 + 15) / 16 * 16, mf = (mf0 + TC_CH - 1) / TC_CH * TC_CH;      // padded shapes the kernel sees
        std::vector<float> T, LP, Ac, Af, bc, bf, dc, df, mu;
        cudaError_t e = cudaSuccess;
        // matrices with d rows on the device -> 64 rows on the host (zero rows for the parameters that do not exist;
        // the columns beyond d of T and LP are zero in the device images already)
        const int d0 = c.d;
        auto fetch_rows = [&](const float* dev, size_t ld, std::vector<float>& h) {
            h.assign((size_t)TC_K * ld, 0.f);
            return cudaMemcpy(h.data(), dev, (size_t)d0 * ld * sizeof(float), cudaMemcpyDeviceToHost);
        };
        if (e == cudaSuccess) e = fetch_rows(P.T, P.ldD, T);
        if (e == cudaSuccess) e = fetch_rows(P.LP, P.ldD, LP);
        if (e == cudaSuccess) e = fetch_rows(P.lv[0].A, P.lv[0].ldA, Ac);
        if (e == cudaSuccess) e = fetch_rows(P.lv[1].A, P.lv[1].ldA, Af);
        if (e == cudaSuccess) e = fetch(P.lv[0].b, mc0, bc);
        if (e == cudaSuccess) e = fetch(P.lv[1].b, mf0, bf);
        if (e == cudaSuccess) e = fetch(P.lv[0].data, mc0, dc);
        if (e == cudaSuccess) e = fetch(P.lv[1].data, mf0, df);
        if (P.lv[0].ldA < mc || P.lv[1].ldA < mf) { err = "tc prepare: operator images narrower than the padded shapes"; return 1; }
        if (e == cudaSuccess) { e = fetch(P.prior_mean, d0, mu); mu.resize(TC_K, 0.f); }
        if (e != cudaSuccess) { err = std::string("tc prepare: ") + cudaGetErrorString(e); return -2; }
        // Diagonal and dense Gaussian likelihoods (distributions.py:304-315, :246-301) are folded into the operators:
        // with prec = L L^T (diagonal: L = diag(var^-1/2)) the log-likelihood -0.5 r^T prec r is -0.5 |L^T r|^2 -- an
        // isotropic unit-variance likelihood of the whitened model L^T G against the whitened data L^T (y - b).
        // (Link.model_output of the records is rebuilt from the engine's own operator, not from these images.)
        std::vector<double> rres[2];
        double lik_var[2] = {c.level[0].lik_var, c.level[1].lik_var};
        for (int l = 0; l < 2; l++) {
            const int kind = c.level[l].lik_kind, m0 = l ? mf0 : mc0, ld = P.lv[l].ldA;
            std::vector<float>& A = l ? Af : Ac;
            const std::vector<float>& bb = l ? bf : bc;
            const std::vector<float>& dd = l ? df : dc;
            std::vector<double>& r = rres[l];
            r.resize(m0);
            for (int j = 0; j < m0; j++) r[j] = (double)dd[j] - (double)bb[j];
            if (kind == TDA_LIK_ISO) continue;
            lik_var[l] = 1.0;
            if (kind == TDA_LIK_DIAG) {
                std::vector<float> var;
                e = fetch(P.lv[l].var, m0, var);
                if (e != cudaSuccess) { err = std::string("tc prepare: ") + cudaGetErrorString(e); return -2; }
                for (int n = 0; n < m0; n++) {
                    if (!(var[n] > 0.f)) { err = "tc: non-positive likelihood variance"; return 1; }
                    const double w = 1.0 / std::sqrt((double)var[n]);
                    for (int k = 0; k < d0; k++) A[(size_t)k * ld + n] = (float)(A[(size_t)k * ld + n] * w);
                    r[n] *= w;
                }
            } else {
                std::vector<float> pf;
                e = fetch(P.lv[l].prec, (size_t)m0 * m0, pf);
                if (e != cudaSuccess) { err = std::string("tc prepare: ") + cudaGetErrorString(e); return -2; }
                std::vector<double> Lc(pf.begin(), pf.end());
                for (int j = 0; j < m0; j++) {            // in-place lower Cholesky factor of the precision
                    double dj = Lc[(size_t)j * m0 + j];
                    for (int k = 0; k < j; k++) dj -= Lc[(size_t)j * m0 + k] * Lc[(size_t)j * m0 + k];
                    if (!(dj > 0.0)) { err = "tc: likelihood precision is not positive definite in float32"; return 1; }
                    dj = std::sqrt(dj);
                    Lc[(size_t)j * m0 + j] = dj;
                    for (int i = j + 1; i < m0; i++) {
                        double v = Lc[(size_t)i * m0 + j];
                        for (int k = 0; k < j; k++) v -= Lc[(size_t)i * m0 + k] * Lc[(size_t)j * m0 + k];
                        Lc[(size_t)i * m0 + j] = v / dj;
                    }
                }
                std::vector<double> row(m0), x(m0);
                for (int k = 0; k <= d0; k++) {            // rows of G^T, then the data residual: x <- x L
                    for (int n = 0; n < m0; n++) x[n] = (k < d0) ? (double)A[(size_t)k * ld + n] : r[n];
                    for (int n = 0; n < m0; n++) {
                        double v = 0;
                        for (int j = n; j < m0; j++) v += x[j] * Lc[(size_t)j * m0 + n];
                        row[n] = v;
                    }
                    for (int n = 0; n < m0; n++) {
                        if (k < d0) A[(size_t)k * ld + n] = (float)row[n];
                        else r[n] = row[n];
                    }
                }
            }
        }
        const int nfc = mf / TC_CH, nch = nfc + 1;
        std::vector<float> hT((size_t)2 * 64 * TC_K), hG((size_t)2 * mc * TC_K), hF((size_t)nch * 2 * TC_CH * TC_K);
        canon_split(T, P.ldD, 0, 64, hT.data(), hT.data() + 64 * TC_K);
        canon_split(Ac, P.lv[0].ldA, 0, mc, hG.data(), hG.data() + (size_t)mc * TC_K);
        for (int ch = 0; ch < nfc; ch++)
            canon_split(Af, P.lv[1].ldA, ch * TC_CH, TC_CH, hF.data() + (size_t)ch * 2 * TC_CH * TC_K,
                        hF.data() + (size_t)ch * 2 * TC_CH * TC_K + TC_CH * TC_K);
        canon_split(LP, P.ldD, 0, TC_CH, hF.data() + (size_t)nfc * 2 * TC_CH * TC_K, hF.data() + (size_t)nfc * 2 * TC_CH * TC_K + TC_CH * TC_K);
        destroy();
        if (e == cudaSuccess) e = cudaMalloc(&dT, hT.size() * 4);
        if (e == cudaSuccess) e = cudaMalloc(&dGc, hG.size() * 4);
        if (e == cudaSuccess) e = cudaMalloc(&dF, hF.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(dT, hT.data(), hT.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dGc, hG.data(), hG.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dF, hF.data(), hF.size() * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { err = std::string("tc prepare: ") + cudaGetErrorString(e); return -2; }
        yc.assign(TC_MAX_MC, 0.f); yf.assign(TC_MAX_MF, 0.f); clp.assign(TC_K, 0.f);
        for (int j = 0; j < mc0; j++) yc[j] = (float)rres[0][j];
        for (int j = 0; j < mf0; j++) yf[j] = (float)rres[1][j];
        for (int n = 0; n < TC_K; n++) {
            double a = 0;
            for (int k = 0; k < TC_K; k++) a += (double)mu[k] * (double)LP[(size_t)k * P.ldD + n];
            clp[n] = (float)a;
        }
        q.T_hl = dT; q.Gc_hl = dGc; q.F_chunks = dF;
        q.rwmh = (c.prop_kind == TDA_PROP_RWMH) ? 1 : 0;
        q.nst = q.rwmh ? 2 : TC_NST;
        q.LP_hl = dF + (size_t)nfc * 2 * TC_CH * TC_K;
        q.mc = mc; q.mf = mf; q.n_chunks = nch; q.J = c.subchain[0];
        q.var_c = (float)lik_var[0]; q.var_f = (float)lik_var[1];
        q.prior_logconst = (float)c.prior_logconst;
        prepared = true;
        return 0;
    }

    int run(Params<float>& P, const tda_config& c, long long iterations, int sm_count, cudaStream_t st) {
        if (!prepared) { int r = prepare(P, c); if (r) return r; }
        cudaError_t e;
        e = cudaMemcpyToSymbolAsync(c_yc, yc.data(), TC_MAX_MC * 4, 0, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_yf, yf.data(), TC_MAX_MF * 4, 0, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_lp, clp.data(), TC_K * 4, 0, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { err = std::string("tc constants: ") + cudaGetErrorString(e); return -2; }
        q.n_pairs = P.Cs / 256;
        const size_t smem = tc_smem_bytes(q.rwmh, q.nst);
        e = q.rwmh ? cudaFuncSetAttribute(da_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                   : cudaFuncSetAttribute(da_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { err = std::string("tc attr: ") + cudaGetErrorString(e); return -2; }
        P.mode = MODE_RUN;
        P.iterations = iterations;
        int grid = q.n_pairs < sm_count ? q.n_pairs : sm_count;
        if (q.rwmh) da_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(P, q);
        else da_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(P, q);
        e = cudaGetLastError();
        if (e != cudaSuccess) { err = std::string("tc launch: ") + cudaGetErrorString(e); return -2; }
        return 0;
    }
};

}  // namespace tda
