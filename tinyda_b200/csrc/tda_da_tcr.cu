// Tensor-core (tcgen05 / TMEM) Delayed-Acceptance kernel, fourth generation ("tcr": whitened state,
// output recursion): two-level DA, pCN proposal, linear forward operators, Gaussian likelihoods
// (BASELINE cfg2), float32 engine.  Reference semantics: chain.py:325-444 (DAChain), proposal.py:261-369
// (pCN), exactly as Tile::base_step / Tile::upper_step in tda_kernels.cuh and as tda_da_tc16.cuh.
//
// What changed against tda_da_tc16.cuh (same warp roles, same fp16-split arithmetic, same z16 normals):
//
//  * Whitened state.  The chain carries w = theta T^-1 (T = the proposal's covariance factor, xi = z T),
//    so pCN's theta' = a theta + b z T (proposal.py:351-355) is the ELEMENTWISE w' = a w + b z on the row
//    threads: no xi = z @ T product, and a and b are per-chain scalars, i.e. per-chain step sizes
//    (adaptive pCN, proposal.py:228-245) cost nothing.  Every operator is composed with T once on the
//    host: F_c = w @ (T G_c^T), F_f = w @ (T G_f^T), the prior's whitened residual w @ (T LP), theta = w @ T.
//  * Output recursion.  The model is linear, so the proposal's coarse output is
//        F_c(theta') = a F_c(theta) + b z @ (T G_c^T):
//    the only MMA of a coarse step has the NORMALS as its A operand (two products, shared memory), it does
//    not depend on the chain's accept/reject history and runs ahead of the row threads.  F_c of the current
//    state lives in TMEM (128 columns per tile); the row threads form a F + b D, the residual and its
//    square sum in one pass, and fold the accepted proposals back into F in a second pass.  The
//    MMA -> row threads -> MMA round trip of the older kernels (3 products + the xi products on the
//    critical path of every coarse step) is gone; executed tensor work drops from 1.40 to 0.75 MFLOP per
//    transition.
//  * F_c(theta) is refreshed from w by an MMA at the start of every fine iteration (recursion depth
//    <= J steps, so no drift), and the coarse log-likelihood of the current state is re-evaluated from
//    the refreshed output: results do not depend on where launches / work units are cut.
//  * The fine stage streams [T G_c^T (refresh) | T G_f^T | T LP | T] in 64-column chunks; theta itself is
//    only formed once per fine iteration (for the Link record and the state other kernels read).
//
// TMEM per tile (256 columns): [0,128) F_c(current) / fine accumulators 1,2; [128,256) the z products of
// a coarse step / during the fine stage the packed A operand split(w) [128,192) and fine accumulator 0.
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "tda_da_tcr.h"
#include "tda_tc_prims.cuh"

namespace tda {

namespace {

constexpr int TR_K = 64;                        // parameters (contraction length)
constexpr int TR_CH = 64;                       // columns per streamed operator chunk
constexpr int TR_NST = 4;                       // ring stages
constexpr int TR_MAX_MC = 128;
constexpr int TR_MAX_MF = 1920;
constexpr int TR_ROW_WARP0 = 0;                 // warps 0..15  (warpgroups 0-3)
constexpr int TR_RNG_WARP0 = 16;                // warps 16..23 (warpgroups 4-5)
constexpr int TR_RNG_WARPS = 8;
constexpr int TR_MMA_WARP0 = 24;                // warps 24, 25 (one per tile)
constexpr int TR_PROD_WARP = 26;                // warp 27 idles (completes warpgroup 6)
constexpr int TR_THREADS = 28 * 32;
constexpr int TR_HK = TR_K / 2;                 // state columns per row thread
constexpr int TR_IMG = 128 * TR_K * 2;          // one [128 x 64] fp16 image: 16 KB
constexpr int TR_TIMG = 64 * TR_K * 2;          // one [64 x 64] fp16 image: 8 KB
constexpr int TR_CHUNK_BYTES = 2 * TR_TIMG;     // hi + lo
constexpr int TR_REGS_ROW = 96, TR_REGS_RNG = 40, TR_REGS_AUX = 40;

constexpr int TR_OFF_M = 0;                                       // T G_c^T (hi | lo), resident
constexpr int TR_OFF_RING = TR_OFF_M + 2 * TR_IMG;
constexpr int TR_OFF_Z = TR_OFF_RING + TR_NST * TR_CHUNK_BYTES;   // [tile 2][buffer 2] images
constexpr int TR_OFF_PART = TR_OFF_Z + 4 * TR_IMG;                // [buf 2][tile 2][half 2][val 2][128] f32
constexpr int TR_OFF_U = TR_OFF_PART + 2 * 2 * 2 * 2 * 128 * 4;   // [buf 2][tile 2][128] f32
constexpr int TR_OFF_PF = TR_OFF_U + 2 * 2 * 128 * 4;             // [tile 2][half 2][val 2][128] f32
constexpr int TR_OFF_UB = TR_OFF_PF + 2 * 2 * 2 * 128 * 4;        // [tile 2][128] uint4: the chain's current Philox block of uniforms
constexpr int TR_OFF_NY = TR_OFF_UB + 2 * 128 * 16;               // f32: [-y_c 128 | -(mu @ LP) 64 | -y_f mf]
constexpr int TR_NY_FLOATS = TR_MAX_MC + TR_K + TR_MAX_MF;
constexpr int TR_OFF_BARS = TR_OFF_NY + TR_NY_FLOATS * 4;
constexpr size_t TR_SMEM_BYTES = TR_OFF_BARS + 512;
static_assert(TR_SMEM_BYTES <= 232448, "shared memory budget");

}  // namespace

struct DaTcrParams {
    const __half* M_hl;       // [hi mc x 64 | lo mc x 64] canonical K-major, T G_c^T * 2^sM
    const __half* chunks;     // n_chunks x [hi 64 x 64 | lo 64 x 64]: 2 refresh chunks (T G_c^T), then T G_f^T, T LP, T
    const float* ny;          // [-(y_c - b_c) 128 | -(mu @ LP) 64 | -(y_f - b_f) mf]
    float* wstate;            // [64][Cs] whitened fine-level state (unscaled)
    int mc, mf, nfc, n_f1, n_chunks, J;
    float var_c, var_f, prior_logconst;
    float bz;                 // 2^(s_w - 12): scaled increment per unit of scaled z product / per unit of 4096 z
    float sc_c, sc_f, sc_p, sc_t;   // accumulator -> true units: coarse output, fine output, whitened prior residual, theta
    float w_scale, w_unscale;
    int n_pairs;
    int ib, nb;               // work units: iterations per block, blocks per launch
    int* progress;            // [n_pairs]
};

namespace {

template <typename T>
__device__ __forceinline__ T* tr_opaque(T* ptr) {
    asm volatile("" : "+l"(ptr));
    return ptr;
}

// "my TMEM / shared-memory accesses are done" -> one arrival per warp
__device__ __forceinline__ void tr_warp_arrive(uint64_t* bar, int lane) {
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(bar);
}

// 3 products (A = split(w) hi | lo packed in TMEM, B hi | lo in shared memory), K = 16 nks
__device__ __forceinline__ void tr_issue_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, uint32_t idesc, int nks) {
    uint32_t accumulate = 0;
#pragma unroll
    for (int pass = 0; pass < 3; pass++) {
        const uint32_t a = a_tmem + (pass == 1 ? 32 : 0);
        const uint64_t b0 = tc::smem_desc_kmajor(pass == 2 ? b_lo : b_hi, 128, (TR_K / 8) * 128);
#pragma unroll
        for (int ks = 0; ks < TR_K / 16; ks++) {
            if (ks < nks) {
                tc::mma_f16_ts(d_tmem, a + ks * 8, b0 + (uint64_t)(ks * 16), idesc, accumulate);
                accumulate = 1;
            }
        }
    }
}

// z-operand products: z @ B_hi + z @ B_lo (+ z_lo @ B_hi when the stream is injected)
__device__ __forceinline__ void tr_issue_z(uint32_t d_tmem, uint32_t z_hi, uint32_t z_lo, bool with_lo, uint32_t b_hi, uint32_t b_lo,
                                           uint32_t idesc, int nks) {
    uint32_t accumulate = 0;
    const int npass = with_lo ? 3 : 2;
    for (int pass = 0; pass < npass; pass++) {
        const uint64_t a0 = tc::smem_desc_kmajor(pass == 2 ? z_lo : z_hi, 128, (TR_K / 8) * 128);
        const uint64_t b0 = tc::smem_desc_kmajor(pass == 1 ? b_lo : b_hi, 128, (TR_K / 8) * 128);
#pragma unroll
        for (int ks = 0; ks < TR_K / 16; ks++) {
            if (ks < nks) {
                tc::mma_f16_ss(d_tmem, a0 + (uint64_t)(ks * 16), b0 + (uint64_t)(ks * 16), idesc, accumulate);
                accumulate = 1;
            }
        }
    }
}

// One 16-normal group (z16 stream) of one chain, packed to fp16 and stored into the chain's row of a
// canonical K-major z image (row_ptr = image + row offset; columns 16 q4 .. 16 q4 + 15).
__device__ __forceinline__ void tr_z_group(unsigned long long seed, long long gchain, unsigned long long group, unsigned char* row_ptr, int q4) {
    const unsigned long long b3 = 3 * group;
    float s[16];
    z16_group_scaled(philox_block(seed, gchain, STREAM_Z, b3), philox_block(seed, gchain, STREAM_Z, b3 + 1),
                     philox_block(seed, gchain, STREAM_Z, b3 + 2), s);
    uint4 w0, w1;
    w0.x = tc::pack_f16x2(s[0], s[1]);   w0.y = tc::pack_f16x2(s[2], s[3]);
    w0.z = tc::pack_f16x2(s[4], s[5]);   w0.w = tc::pack_f16x2(s[6], s[7]);
    w1.x = tc::pack_f16x2(s[8], s[9]);   w1.y = tc::pack_f16x2(s[10], s[11]);
    w1.z = tc::pack_f16x2(s[12], s[13]); w1.w = tc::pack_f16x2(s[14], s[15]);
    *reinterpret_cast<uint4*>(row_ptr + (2 * q4) * 128) = w0;
    *reinterpret_cast<uint4*>(row_ptr + (2 * q4 + 1) * 128) = w1;
}

// Work distribution: as in tda_da_tc16.cuh (units = (tile pair, iteration block), dealt round-robin,
// block-major; chain state travels through global memory, `progress[pair]` orders the hand-off).
struct TrUnit {
    int pair, blk, it0, it1;
};
__device__ __forceinline__ bool tr_unit(const DaTcrParams& q, int iters, int k, TrUnit& u) {
    const int idx = (int)blockIdx.x + k * (int)gridDim.x;
    if (idx >= q.n_pairs * q.nb) return false;
    u.blk = idx / q.n_pairs;
    u.pair = idx - u.blk * q.n_pairs;
    u.it0 = u.blk * q.ib;
    u.it1 = min(u.it0 + q.ib, iters);
    return true;
}
__device__ __forceinline__ int tr_ld_acquire(const int* ptr) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void tr_red_release_add(int* ptr, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

// the two fp16 values of a packed pair, widened
__device__ __forceinline__ unsigned long long tr_h2_to_f2(uint32_t h2) { return f2pack(tc::f16lo_to_f32(h2), tc::f16hi_to_f32(h2)); }

// PAD = false: d == 64, the native shape; PAD = true: d in {16, 32, 48} (zero-padded operators)
template <bool PAD>
__global__ void __launch_bounds__(TR_THREADS, 1)
da_tcr_kernel(const __grid_constant__ Params<float> p, const __grid_constant__ DaTcrParams q) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sM = smem + TR_OFF_M;
    unsigned char* ring = smem + TR_OFF_RING;
    unsigned char* zbuf = smem + TR_OFF_Z;
    float* s_part = reinterpret_cast<float*>(smem + TR_OFF_PART);
    float* s_u = reinterpret_cast<float*>(smem + TR_OFF_U);
    float* s_pf = reinterpret_cast<float*>(smem + TR_OFF_PF);
    const float* s_ny = reinterpret_cast<const float*>(smem + TR_OFF_NY);
    uint4* s_ub = reinterpret_cast<uint4*>(smem + TR_OFF_UB);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TR_OFF_BARS);
    uint64_t* bar_res = bars;                       // resident operands landed
    uint64_t* bar_reqR = bars + 1;                  // [tile] rows -> MMA: split(w_current) stored, fine accumulators consumed
    uint64_t* bar_reqA = bars + 3;                  // [tile] rows -> MMA: split(w_proposal) stored, coarse outputs dead
    uint64_t* bar_dfree = bars + 5;                 // [tile] rows -> MMA: z products consumed
    uint64_t* bar_dfull = bars + 7;                 // [tile] MMA  -> rows: z products (and the refresh before them) complete
    uint64_t* bar_rdone = bars + 9;                 // [tile] MMA  -> MMA : refresh complete (its A operand may be overwritten)
    uint64_t* bar_reqF = bars + 11;                 // [tile][3] rows -> MMA: fine-chunk accumulator consumed
    uint64_t* bar_respF = bars + 17;                // [tile][3] MMA  -> rows: fine chunk complete
    uint64_t* bar_zfull = bars + 23;                // [tile][buffer] RNG  -> MMA, rows
    uint64_t* bar_zfree = bars + 27;                // [tile][buffer] MMA + rows -> RNG
    uint64_t* bar_full = bars + 31;                 // [NST]
    uint64_t* bar_empty = bars + 31 + TR_NST;       // [NST]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 31 + 2 * TR_NST);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (warp == TR_MMA_WARP0) tc::tmem_alloc(s_tmem, 512);
    if (tid == TR_PROD_WARP * 32) {
        tc::mbar_init(bar_res, 1);
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(bar_reqR + i, 8);
            tc::mbar_init(bar_reqA + i, 8);
            tc::mbar_init(bar_dfree + i, 8);
            tc::mbar_init(bar_dfull + i, 1);
            tc::mbar_init(bar_rdone + i, 1);
        }
        for (int i = 0; i < 6; i++) {
            tc::mbar_init(bar_reqF + i, 8);
            tc::mbar_init(bar_respF + i, 1);
        }
        for (int i = 0; i < 4; i++) {
            tc::mbar_init(bar_zfull + i, 4);
            tc::mbar_init(bar_zfree + i, 9);           // the MMA's commit + the tile's 8 row warps
        }
        for (int s = 0; s < TR_NST; s++) { tc::mbar_init(bar_full + s, 1); tc::mbar_init(bar_empty + s, 2); }
        tc::fence_mbar_init();
    }
    const int d = PAD ? p.d : TR_K;
    if (PAD) {
        uint4* zz = reinterpret_cast<uint4*>(zbuf);
        for (int i = tid; i < 4 * TR_IMG / 16; i += TR_THREADS) zz[i] = make_uint4(0u, 0u, 0u, 0u);
        tc::fence_proxy_async_smem();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = __shfl_sync(0xffffffffu, *s_tmem, 0);

    const int J = q.J, mc = q.mc, NF1 = q.n_f1, NCH = q.n_chunks;
    const int iters = (int)p.iterations;
    const bool inj = p.rng_mode == TDA_RNG_INJECTED;

    if (warp >= TR_RNG_WARP0 && warp < TR_RNG_WARP0 + TR_RNG_WARPS) {
        // =====================================================================================
        // RNG warps: thread = one chain of one tile; 64 normals per coarse step -> z image(s)
        // =====================================================================================
        tc::setmaxnreg_dec<TR_REGS_RNG>();
        const int t = (warp - TR_RNG_WARP0) >> 2;
        const int row = (warp & 3) * 32 + lane;
        unsigned char* zt = zbuf + (size_t)t * 2 * TR_IMG + (row >> 3) * ((TR_K / 8) * 128) + (row & 7) * 16;
        uint64_t* zfull = bar_zfull + t * 2;
        uint64_t* zfree = bar_zfree + t * 2;
        unsigned n = 0;                                    // coarse steps produced
        TrUnit un;
        for (int uk = 0; tr_unit(q, iters, uk, un); uk++) {
            const int g = un.pair * 256 + t * 128 + row;
            const long long gchain = p.chain_offset + g;
            const bool live = g < p.C;
            long long tb = p.t_base + (long long)un.it0 * J;
            const int nsteps = (un.it1 - un.it0) * J;
            for (int st = 0; st < nsteps; st++, n++, tb++) {
                const int b = inj ? 0 : (int)(n & 1);
                const unsigned use = inj ? n : (n >> 1);
                if (use >= 1) tc::mbar_wait(zfree + b, (uint32_t)((use - 1) & 1));
                unsigned char* dst = zt + (size_t)b * TR_IMG;
                if (!inj) {
                    const unsigned long long grp0 = (unsigned long long)(tb * (d >> 4));
#pragma unroll 1
                    for (int q4 = 0; q4 < (d >> 4); q4++) tr_z_group(p.seed, gchain, grp0 + q4, dst, q4);
                } else {
                    const long long z0 = tb * d;
                    for (int kg = 0; kg < (d >> 3); kg++) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const long long idx = z0 + kg * 8 + 2 * i;
                            const float x0 = (live && idx < p.zlen) ? p.zs[(size_t)g * p.zlen + idx] * Z16_SCALE : 0.0f;
                            const float x1 = (live && idx + 1 < p.zlen) ? p.zs[(size_t)g * p.zlen + idx + 1] * Z16_SCALE : 0.0f;
                            tc::split_f16x2(x0, x1, hi[i], lo[i]);
                        }
                        *reinterpret_cast<uint4*>(dst + kg * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(dst + TR_IMG + kg * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                tc::fence_proxy_async_smem();          // generic-proxy writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(zfull + b);
            }
        }
    } else if (warp >= TR_MMA_WARP0) {
        tc::setmaxnreg_dec<TR_REGS_AUX>();
        if (warp == TR_PROD_WARP) {
            // ===== producer: resident operands once, then the operator chunk ring (NCH chunks per fine iteration) =====
            if (lane == 0) {
                unsigned total_chunks = 0;
                {
                    TrUnit un;
                    for (int uk = 0; tr_unit(q, iters, uk, un); uk++) total_chunks += (unsigned)(un.it1 - un.it0) * NCH;
                }
                const uint32_t bI = (uint32_t)mc * TR_K * 2;       // one mc-row image
                const uint32_t bNY = (uint32_t)(TR_MAX_MC + TR_K + q.mf) * 4;
                tc::mbar_expect_tx(bar_res, 2 * bI + bNY);
                tc::bulk_g2s(smem + TR_OFF_NY, q.ny, bNY, bar_res);
                tc::bulk_g2s(sM, q.M_hl, bI, bar_res);
                tc::bulk_g2s(sM + TR_IMG, q.M_hl + (size_t)mc * TR_K, bI, bar_res);
                for (unsigned g = 0; g < total_chunks; g++) {
                    const int st = (int)(g % TR_NST);
                    if (g >= TR_NST) tc::mbar_wait(bar_empty + st, (uint32_t)(((g / TR_NST) - 1) & 1));
                    const int c = (int)(g % NCH);
                    tc::mbar_expect_tx(bar_full + st, TR_CHUNK_BYTES);
                    tc::bulk_g2s(ring + (size_t)st * TR_CHUNK_BYTES, q.chunks + (size_t)c * (TR_CHUNK_BYTES / 2), TR_CHUNK_BYTES, bar_full + st);
                }
            }
        } else if (warp < TR_PROD_WARP) {
            // ===== MMA issuer of tile t (warp-uniform control flow, one elected lane issues) =====
            const int t = warp - TR_MMA_WARP0;
            tc::mbar_wait(bar_res, 0);
            const uint32_t tF = tbase + t * 256, tD = tF + 128;        // A operand (packed split(w)): tD + [0, 64)
            const uint32_t sM_hi = tc::smem_u32(sM), sM_lo = sM_hi + TR_IMG;
            const uint32_t z_base = tc::smem_u32(zbuf) + t * 2 * TR_IMG;
            const uint32_t ring_base = tc::smem_u32(ring);
            const uint32_t idesc_c = tc::idesc_f16(128, mc), idesc_64 = tc::idesc_f16(128, 64);
            uint64_t* reqR = bar_reqR + t;
            uint64_t* reqA = bar_reqA + t;
            uint64_t* dfree = bar_dfree + t;
            uint64_t* dfull = bar_dfull + t;
            uint64_t* rdone = bar_rdone + t;
            uint64_t* reqF = bar_reqF + t * 3;
            uint64_t* respF = bar_respF + t * 3;
            uint32_t pr = 0, pa = 0, pd = 0, prd = 0, pf = 0;        // pf: bit b = parity of reqF[b]
            const int nks = PAD ? (d >> 4) : TR_K / 16;
            unsigned n = 0, gch = 0;
            int total_it = 0;
            {
                TrUnit un;
                for (int uk = 0; tr_unit(q, iters, uk, un); uk++) total_it += un.it1 - un.it0;
            }
            // accumulator b of the fine stage: 0 -> tD + 64, 1 -> tF, 2 -> tF + 64
            for (int itg = 0; itg < total_it; itg++) {
                // ---- refresh: F_c(current) = w @ (T G_c^T) into tF[0:128), two 64-column chunks ----
                tc::mbar_wait(reqR, pr); pr ^= 1;
                for (int c = 0; c < 2; c++, gch++) {
                    const int st = (int)(gch % TR_NST);
                    tc::mbar_wait(bar_full + st, (uint32_t)((gch / TR_NST) & 1));
                    tc::fence_after_sync();
                    if (tc::elect_one()) {
                        const uint32_t b_hi = ring_base + st * TR_CHUNK_BYTES, b_lo = b_hi + TR_TIMG;
                        tr_issue_w(tF + c * TR_CH, tD, b_hi, b_lo, idesc_64, nks);
                        tc::mma_commit(bar_empty + st);
                        if (c == 1) tc::mma_commit(rdone);
                    }
                    __syncwarp();
                }
                // the z products of the first coarse step overwrite the A operand of the refresh
                tc::mbar_wait(rdone, prd); prd ^= 1;
                // ---- coarse steps: D = z @ (T G_c^T), ahead of the row threads ----
                for (int j = 0; j < J; j++, n++) {
                    const int b = inj ? 0 : (int)(n & 1);
                    const unsigned use = inj ? n : (n >> 1);
                    const uint32_t z_hi = z_base + b * TR_IMG, z_lo = z_base + TR_IMG;
                    tc::mbar_wait(bar_zfull + t * 2 + b, (uint32_t)(use & 1));
                    tc::mbar_wait(dfree, pd); pd ^= 1;
                    tc::fence_after_sync();
                    if (tc::elect_one()) {
                        tr_issue_z(tD, z_hi, z_lo, inj, sM_hi, sM_lo, idesc_c, nks);
                        tc::mma_commit(dfull);
                        tc::mma_commit(bar_zfree + t * 2 + b);
                    }
                    __syncwarp();
                }
                // ---- fine stage: w_proposal @ [T G_f^T | T LP | T], three accumulators in flight ----
                tc::mbar_wait(reqA, pa); pa ^= 1;
                for (int c = 0, b = 0; c < NF1; c++, gch++, b = (b == 2 ? 0 : b + 1)) {
                    const int st = (int)(gch % TR_NST);
                    if (c >= 3) {
                        tc::mbar_wait(reqF + b, (pf >> b) & 1u);
                        pf ^= 1u << b;
                    }
                    tc::mbar_wait(bar_full + st, (uint32_t)((gch / TR_NST) & 1));
                    tc::fence_after_sync();
                    if (tc::elect_one()) {
                        const uint32_t b_hi = ring_base + st * TR_CHUNK_BYTES, b_lo = b_hi + TR_TIMG;
                        const uint32_t acc = (b == 0) ? tD + TR_CH : tF + (b - 1) * TR_CH;
                        tr_issue_w(acc, tD, b_hi, b_lo, idesc_64, nks);
                        tc::mma_commit(respF + b);
                        tc::mma_commit(bar_empty + st);
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else {
        // =====================================================================================
        // row threads: two per chain (column halves)
        // =====================================================================================
        tc::setmaxnreg_inc<TR_REGS_ROW>();
        const int rw = warp - TR_ROW_WARP0;
        const int t = rw >> 3;                         // tile within the pair
        const int h = (rw >> 2) & 1;                   // column half
        const int wq = rw & 3;                         // TMEM lane quarter (= warp % 4)
        const int cl = wq * 32 + lane;                 // chain within the tile
        const int col0 = h * TR_HK;                    // first state column of this thread
        const int nk = PAD ? min(TR_HK, max(0, d - col0)) : TR_HK;
        const uint32_t tF = tbase + ((uint32_t)(wq * 32) << 16) + t * 256;
        const uint32_t tD = tF + 128;
        uint64_t* reqR = bar_reqR + t;
        uint64_t* reqA = bar_reqA + t;
        uint64_t* dfree = bar_dfree + t;
        uint64_t* dfull = bar_dfull + t;
        uint64_t* reqF = bar_reqF + t * 3;
        uint64_t* respF = bar_respF + t * 3;
        uint64_t* zfull = bar_zfull + t * 2;
        uint64_t* zfree = bar_zfree + t * 2;
        uint32_t phD = 0, phF = 0;                    // phF: bit b = parity of respF[b]
        unsigned n = 0;                                // coarse steps consumed (z buffer / parity bookkeeping)
        int sbuf = 0;
        const LevelP<float>& l0 = p.lv[0];
        const LevelP<float>& l1 = p.lv[1];
        const float inv2vc = -0.5f / q.var_c, inv2vf = -0.5f / q.var_f;
        const float sc_c = q.sc_c, sc_f = q.sc_f, sc_p = q.sc_p, sc_t = q.sc_t;
        const unsigned long long sc_c2 = f2pack(sc_c, sc_c);
        const float w_scale = q.w_scale, w_unscale = q.w_unscale;
        const int ngc = mc >> 4, gc0 = h ? (ngc + 1) / 2 : 0, gc1 = h ? ngc : (ngc + 1) / 2;
        const unsigned char* zrow = zbuf + (size_t)t * 2 * TR_IMG + (cl >> 3) * ((TR_K / 8) * 128) + (cl & 7) * 16 + (h * 4) * 128;
        const bool leader = (h == 0 && wq == 0);
        tc::mbar_wait(bar_res, 0);                     // the data vector arrives with the resident operands

        TrUnit un;
        for (int uk = 0; tr_unit(q, iters, uk, un); uk++) {
            const int pair = un.pair;
            const int g = pair * 256 + t * 128 + cl;                   // chain slot (padded arrays)
            if (un.blk > 0) {
                if (leader) while (tr_ld_acquire(q.progress + pair) < 2 * un.blk) __nanosleep(200);
                tc::named_bar_sync(3 + t, 256);
            }
            const long long gchain = p.chain_offset + g;
            const bool live = g < p.C;
            const size_t cs = (size_t)p.Cs;
            const size_t off0 = (size_t)col0 * cs + g;
            float w[TR_HK];                                  // current coarse state, whitened, scaled by 2^s_w
            {
                const float* src = tr_opaque(q.wstate + off0);
#pragma unroll
                for (int k = 0; k < TR_HK; k++) w[k] = (k < nk) ? __ldcg(src + k * cs) * w_scale : 0.0f;
            }
            float like_c = 0.0f, like_f = __ldcg(l1.like + g), prior_f = __ldcg(l1.prior + g);
            float like_cs = 0.0f, like_cur = __ldcg(l0.like + g);       // like_cur: coarse Link of the aligned state
            long long ucur = __ldcg(p.ucur + g);
            int nacc_c = 0, nacc_f = 0, acc_any = 0;
            // per-chain pCN step: a = sqrt(1 - s^2), increment b = s 2^(s_w - 12) per unit of 4096 z
            float ca, cb;
            {
                const float s = __ldcg(p.scaling + g);
                ca = sqrtf(1.0f - s * s);
                cb = s * q.bz;
            }
            // adaptive global scaling (proposal.py:228-245): the `accepted` window of the last `period` entries --
            // coarse decisions and the alignment entry of every fine iteration -- as a ring
            // in global memory, shared with the generic kernel; kept by the h == 0 thread of the chain
            const bool adaptive = p.adaptive != 0;
            int wsum = (adaptive && h == 0) ? __ldcg(p.win_sum + g) : 0;
            long long wc = p.wcount + (long long)un.it0 * (J + 1);
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

            auto store_A = [&]() {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; i++) tc::split_f16x2(w[2 * i], w[2 * i + 1], hi[i], lo[i]);
                tc::tmem_st16(tD + h * 16, hi);
                tc::tmem_st16(tD + 32 + h * 16, lo);
                tc::tmem_wait_st();
            };
            // accept-test uniforms: one Philox block serves four draws (kept in shared memory, not in registers)
            uint4* my_ub = s_ub + t * 128 + cl;
            if (h == 0 && !inj && (ucur & 3)) *my_ub = philox_block(p.seed, gchain, STREAM_U, (unsigned long long)ucur >> 2);
            auto draw_u = [&]() -> float {
                if (inj) return (live && ucur < p.ulen) ? p.us[(size_t)g * p.ulen + ucur] : 0.5f;
                const unsigned o = (unsigned)ucur & 3u;
                if (o == 0) *my_ub = philox_block(p.seed, gchain, STREAM_U, (unsigned long long)ucur >> 2);
                return u01<float>(reinterpret_cast<const uint32_t*>(my_ub)[o]);
            };
            auto wait_bar = [&](uint64_t* bar, uint32_t parity) {
                if (leader) tc::mbar_wait(bar, parity);
                tc::named_bar_sync(3 + t, 256);
                tc::fence_after_sync();
            };

            for (int it = un.it0; it < un.it1; it++) {
                // ---- iteration start: the MMA refreshes F_c(current) from split(w) ----
                store_A();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) { tc::mbar_arrive(reqR); tc::mbar_arrive(dfree); }

                for (int j = 0; j < J; j++, n++) {
                    const int zb = inj ? 0 : (int)(n & 1);
                    const unsigned use = inj ? n : (n >> 1);
                    float u_mine = 0.0f;
                    if (h == 0) u_mine = draw_u();
                    ucur++;
                    // z products (and, for j == 0, the refresh issued before them) are complete; the
                    // normals themselves are visible to this thread through their own barrier
                    if (leader) { tc::mbar_wait(dfull, phD); tc::mbar_wait(zfull + zb, (uint32_t)(use & 1)); }
                    phD ^= 1;
                    tc::named_bar_sync(3 + t, 256);
                    tc::fence_after_sync();
                    // ---- pass 1: F' = a F + b D, residual, square sum (j == 0: also of the current state) ----
                    const unsigned long long ca2 = f2pack(ca, ca), cb2 = f2pack(cb, cb);
                    unsigned long long ssq2a = 0ull, ssq2b = 0ull;
                    const bool first = (j == 0);
                    float ssq_cur = 0.0f;
                    if (first) {
                        // the coarse Link of the current state from the refreshed output (once per fine iteration)
                        unsigned long long cur2a = 0ull, cur2b = 0ull;
                        for (int g8 = gc0 * 2; g8 < gc1 * 2; g8++) {
                            uint32_t fv[8];
                            tc::tmem_ld8(tF + g8 * 8, fv);
                            const float4* ny4 = reinterpret_cast<const float4*>(s_ny + g8 * 8);
                            const float4 ya = ny4[0], yb = ny4[1];
                            tc::tmem_wait_ld();
                            const unsigned long long c0 = f2fma(f2pack(__uint_as_float(fv[0]), __uint_as_float(fv[1])), sc_c2, f2pack(ya.x, ya.y));
                            const unsigned long long c1 = f2fma(f2pack(__uint_as_float(fv[2]), __uint_as_float(fv[3])), sc_c2, f2pack(ya.z, ya.w));
                            const unsigned long long c2 = f2fma(f2pack(__uint_as_float(fv[4]), __uint_as_float(fv[5])), sc_c2, f2pack(yb.x, yb.y));
                            const unsigned long long c3 = f2fma(f2pack(__uint_as_float(fv[6]), __uint_as_float(fv[7])), sc_c2, f2pack(yb.z, yb.w));
                            cur2a = f2fma(c0, c0, cur2a); cur2b = f2fma(c1, c1, cur2b);
                            cur2a = f2fma(c2, c2, cur2a); cur2b = f2fma(c3, c3, cur2b);
                        }
                        float e0, e1, e2, e3;
                        f2unpack(cur2a, e0, e1); f2unpack(cur2b, e2, e3);
                        ssq_cur = (e0 + e1) + (e2 + e3);
                    }
                    for (int g8 = gc0 * 2; g8 < gc1 * 2; g8++) {
                        uint32_t fv[8], dv[8];
                        tc::tmem_ld8(tF + g8 * 8, fv);
                        tc::tmem_ld8(tD + g8 * 8, dv);
                        const float4* ny4 = reinterpret_cast<const float4*>(s_ny + g8 * 8);
                        const float4 ya = ny4[0], yb = ny4[1];
                        tc::tmem_wait_ld();
                        const unsigned long long p0 = f2fma(f2pack(__uint_as_float(fv[0]), __uint_as_float(fv[1])), ca2, f2mul(f2pack(__uint_as_float(dv[0]), __uint_as_float(dv[1])), cb2));
                        const unsigned long long p1 = f2fma(f2pack(__uint_as_float(fv[2]), __uint_as_float(fv[3])), ca2, f2mul(f2pack(__uint_as_float(dv[2]), __uint_as_float(dv[3])), cb2));
                        const unsigned long long p2 = f2fma(f2pack(__uint_as_float(fv[4]), __uint_as_float(fv[5])), ca2, f2mul(f2pack(__uint_as_float(dv[4]), __uint_as_float(dv[5])), cb2));
                        const unsigned long long p3 = f2fma(f2pack(__uint_as_float(fv[6]), __uint_as_float(fv[7])), ca2, f2mul(f2pack(__uint_as_float(dv[6]), __uint_as_float(dv[7])), cb2));
                        const unsigned long long r0 = f2fma(p0, sc_c2, f2pack(ya.x, ya.y)), r1 = f2fma(p1, sc_c2, f2pack(ya.z, ya.w));
                        const unsigned long long r2 = f2fma(p2, sc_c2, f2pack(yb.x, yb.y)), r3 = f2fma(p3, sc_c2, f2pack(yb.z, yb.w));
                        ssq2a = f2fma(r0, r0, ssq2a); ssq2b = f2fma(r1, r1, ssq2b);
                        ssq2a = f2fma(r2, r2, ssq2a); ssq2b = f2fma(r3, r3, ssq2b);
                    }
                    float ssq;
                    {
                        float e0, e1, e2, e3;
                        f2unpack(ssq2a, e0, e1); f2unpack(ssq2b, e2, e3);
                        ssq = (e0 + e1) + (e2 + e3);
                    }
                    float* sp = s_part + ((sbuf * 2 + t) * 2) * 256;       // [half][val][128]
                    float* su = s_u + (sbuf * 2 + t) * 128;
                    sp[h * 256 + cl] = ssq;
                    if (first) sp[h * 256 + 128 + cl] = ssq_cur;
                    if (h == 0) su[cl] = u_mine;
                    sbuf ^= 1;
                    tc::named_bar_sync(1 + t, 256);
                    if (first) {
                        // coarse Link of the current state, evaluated from the refreshed output (posterior.py:78-110)
                        like_c = inv2vc * (sp[128 + cl] + sp[256 + 128 + cl]);
                        like_cs = like_c;
                    }
                    const float like_p = inv2vc * (sp[cl] + sp[256 + cl]);
                    const float u = su[cl];
                    const float alpha = isnan(like_p) ? 0.0f : expf(like_p - like_c);
                    const bool acc = u < alpha;
                    const bool any = __any_sync(0xffffffffu, acc);
                    const float c1 = acc ? ca : 1.0f, c2 = acc ? cb : 0.0f;
                    const unsigned long long c12 = f2pack(c1, c1), c22 = f2pack(c2, c2);
                    // ---- pass 2: accepted proposals become the current output ----
                    if (any) {
                        for (int g8 = gc0 * 2; g8 < gc1 * 2; g8++) {
                            uint32_t fv[8], dv[8];
                            tc::tmem_ld8(tF + g8 * 8, fv);
                            tc::tmem_ld8(tD + g8 * 8, dv);
                            tc::tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const unsigned long long f0 = f2pack(__uint_as_float(fv[2 * i]), __uint_as_float(fv[2 * i + 1]));
                                const unsigned long long d0 = f2pack(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1]));
                                float x0, x1;
                                f2unpack(f2fma(f0, c12, f2mul(d0, c22)), x0, x1);
                                fv[2 * i] = __float_as_uint(x0); fv[2 * i + 1] = __float_as_uint(x1);
                            }
                            tc::tmem_st8(tF + g8 * 8, fv);
                        }
                        tc::tmem_wait_st();
                    }
                    if (j + 1 < J) tr_warp_arrive(dfree, lane);           // z products consumed
                    // ---- the state itself: w' = a w + b z, elementwise ----
                    if (any) {
                        const unsigned char* zp = zrow + (size_t)zb * TR_IMG;
#pragma unroll
                        for (int kg = 0; kg < 4; kg++) {
                            const uint4 zh = *reinterpret_cast<const uint4*>(zp + kg * 128);
                            unsigned long long z0 = tr_h2_to_f2(zh.x), z1 = tr_h2_to_f2(zh.y), z2 = tr_h2_to_f2(zh.z), z3 = tr_h2_to_f2(zh.w);
                            if (inj) {
                                const uint4 zl = *reinterpret_cast<const uint4*>(zp + TR_IMG + kg * 128);
                                const unsigned long long one2 = f2pack(1.0f, 1.0f);
                                z0 = f2fma(tr_h2_to_f2(zl.x), one2, z0); z1 = f2fma(tr_h2_to_f2(zl.y), one2, z1);
                                z2 = f2fma(tr_h2_to_f2(zl.z), one2, z2); z3 = f2fma(tr_h2_to_f2(zl.w), one2, z3);
                            }
                            float* wk = w + kg * 8;
                            f2unpack(f2fma(f2pack(wk[0], wk[1]), c12, f2mul(z0, c22)), wk[0], wk[1]);
                            f2unpack(f2fma(f2pack(wk[2], wk[3]), c12, f2mul(z1, c22)), wk[2], wk[3]);
                            f2unpack(f2fma(f2pack(wk[4], wk[5]), c12, f2mul(z2, c22)), wk[4], wk[5]);
                            f2unpack(f2fma(f2pack(wk[6], wk[7]), c12, f2mul(z3, c22)), wk[6], wk[7]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(zfree + zb);            // normals consumed (shared-memory reads only)
                    if (acc) { like_c = like_p; acc_any = 1; nacc_c++; }
                    // ---- coarse-level record (chain_coarse_i, sampler.py:421-436): whitened parameters, log-likelihood,
                    // accept flag; theta = w T, Link.prior and Link.model_output are filled in when first fetched ----
                    if (l0.store) {
                        const long long r0 = p.rec[0] + (long long)it * J + j;
                        if (r0 < l0.hist_cap) {
                            if (l0.store & TDA_STORE_THETA) {
                                float* dst = tr_opaque(l0.h_theta + (size_t)r0 * d * cs + off0);
#pragma unroll
                                for (int k = 0; k < TR_HK; k++)
                                    if (k < nk) __stcs(dst + k * cs, w[k] * w_unscale);
                            }
                            if (h == 0) {
                                if (l0.store & TDA_STORE_STATS) __stcs(l0.h_like + (size_t)r0 * cs + g, like_c);
                                if (l0.store & TDA_STORE_ACCEPT) l0.h_acc[(size_t)r0 * cs + g] = (uint8_t)acc;
                            }
                        }
                    }
                    if (adaptive) {
                        window_append(acc ? 1 : 0);
                        const long long tnow = p.t_base + (long long)it * J + j + 1;       // proposal.t after this step
                        if ((tnow % p.period) == 0) {
                            if (h == 0) {
                                const long long kk = tnow / p.period - 1;
                                const float rate = (float)wsum / (float)p.period;
                                const float s_old = __ldcg(p.scaling + g);
                                p.scaling[g] = expf(logf(s_old) + powf(p.gamma, (float)(-(double)kk)) * (rate - p.alpha_star));
                            }
                            tc::named_bar_sync(1 + t, 256);          // the other half-thread of the chain reads the new step
                            const float s_new = __ldcg(p.scaling + g);
                            ca = sqrtf(1.0f - s_new * s_new);
                            cb = s_new * q.bz;
                        }
                    }
                }
                // ---- fine level: w @ [T G_f^T | T LP | T] streamed in 64-column chunks ----
                // the A operand overlays the z-product columns the OTHER half-thread of this chain may still be
                // reading in its pass 2: both halves meet first
                tc::fence_before_sync();
                tc::named_bar_sync(1 + t, 256);
                tc::fence_after_sync();
                store_A();
                tr_warp_arrive(reqA, lane);                // split(w_proposal) stored; coarse outputs and z products dead
                float u2 = 0.0f;
                if (h == 0) u2 = draw_u();
                float ssq_f = 0.0f, ssq_p = 0.0f;
                int accf = 0;
                float like_fp = 0.0f, prior_p = 0.0f;
                for (int c = 0, b = 0; c < NF1; c++, b = (b == 2 ? 0 : b + 1)) {
                    wait_bar(respF + b, (phF >> b) & 1u);
                    phF ^= 1u << b;
                    const uint32_t accb = (b == 0) ? tD + TR_CH : tF + (b - 1) * TR_CH;
                    uint32_t v0[16], v1[16];
                    tc::tmem_ld16(accb + col0, v0);
                    tc::tmem_ld16(accb + col0 + 16, v1);
                    tc::tmem_wait_ld();
                    if (c + 3 < NF1) tr_warp_arrive(reqF + b, lane);
                    if (c <= q.nfc) {
                        const bool prior_chunk = (c == q.nfc);
                        const float4* ny4 = reinterpret_cast<const float4*>(prior_chunk ? s_ny + TR_MAX_MC + col0
                                                                                        : s_ny + TR_MAX_MC + TR_K + c * TR_CH + col0);
                        const float scl = prior_chunk ? sc_p : sc_f;
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
                        if (prior_chunk) ssq_p += acc2;
                        else ssq_f += acc2;
                        if (prior_chunk) {
                            // ---- second-stage accept / reject (chain.py:475-483) ----
                            float* sf = s_pf + (t * 2) * 256;
                            sf[h * 256 + cl] = ssq_f;
                            sf[h * 256 + 128 + cl] = ssq_p;
                            float* su = s_u + (sbuf * 2 + t) * 128;
                            if (h == 0) su[cl] = u2;
                            sbuf ^= 1;
                            tc::named_bar_sync(1 + t, 256);
                            like_fp = inv2vf * (sf[cl] + sf[256 + cl]);
                            prior_p = -0.5f * (q.prior_logconst + (sf[128 + cl] + sf[256 + 128 + cl]));
                            if (acc_any) {
                                const float alpha2 = expf(like_fp - like_f + like_cs - like_c);
                                accf = (su[cl] < alpha2) ? 1 : 0;
                                ucur++;
                            }
                        }
                    } else {
                        // ---- theta chunk: the Link record of this iteration and the state other kernels read ----
                        float* th_state = tr_opaque(l1.theta + off0);
                        const long long r = p.rec[1] + it;
                        const bool rec_on = r < l1.hist_cap;
                        float* th_hist = (rec_on && (l1.store & TDA_STORE_THETA)) ? tr_opaque(l1.h_theta + (size_t)r * d * cs + off0) : nullptr;
                        float* s1 = tr_opaque(p.sum1 + off0);
                        float* s2 = tr_opaque(p.sum2 + off0);
#pragma unroll
                        for (int k0 = 0; k0 < TR_HK; k0 += 8) {
                            float x[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const int k = k0 + i;
                                x[i] = 0.0f;
                                if (k < nk) {
                                    const float prop = __uint_as_float(k < 16 ? v0[k & 15] : v1[k & 15]) * sc_t;
                                    x[i] = accf ? prop : __ldcg(th_state + k * cs);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const int k = k0 + i;
                                if (k < nk) {
                                    if (accf) th_state[k * cs] = x[i];
                                    if (th_hist) th_hist[k * cs] = x[i];
                                    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(s1 + k * cs), "f"(x[i]) : "memory");
                                    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(s2 + k * cs), "f"(x[i] * x[i]) : "memory");
                                }
                            }
                        }
                        if (accf) { like_f = like_fp; prior_f = prior_p; nacc_f++; }
                        like_cur = accf ? like_c : like_cs;
                        if (adaptive) window_append(accf);           // the alignment entry (chain.py:391, :397)
                        if (rec_on && h == 0) {
                            if (l1.store & TDA_STORE_STATS) { l1.h_prior[(size_t)r * cs + g] = prior_f; l1.h_like[(size_t)r * cs + g] = like_f; }
                            if (l1.store & TDA_STORE_ACCEPT) l1.h_acc[(size_t)r * cs + g] = (uint8_t)accf;
                        }
                    }
                }
                // ---- the whitened state: promote or rewind (chain.py:385-398) ----
                {
                    float* ws = tr_opaque(q.wstate + off0);
                    if (accf) {
#pragma unroll
                        for (int k = 0; k < TR_HK; k++)
                            if (k < nk) ws[k * cs] = w[k] * w_unscale;
                    } else {
#pragma unroll
                        for (int k = 0; k < TR_HK; k++)
                            if (k < nk) w[k] = __ldcg(ws + k * cs) * w_scale;
                    }
                }
                acc_any = 0;
                // every TMEM read of the fine accumulators is done: the next refresh may overwrite them
                // (signalled together with the stored A operand at the top of the loop)
            }
            // ---- write the chain state back (layout shared with the other kernels) ----
            {
                // the coarse level's state equals the fine one after the alignment step
                const float* src = tr_opaque(l1.theta + off0);
                float* dst = tr_opaque(l0.theta + off0);
#pragma unroll
                for (int k = 0; k < TR_HK; k++)
                    if (k < nk) dst[k * cs] = __ldcg(src + k * cs);
            }
            if (h == 0) {
                // like_c / like_cs are re-evaluated from the refreshed output when the chain continues on this kernel;
                // other kernels find the last coarse value of the current state
                const float lc = like_cur;
                l0.like[g] = lc; l0.prior[g] = prior_f; l1.like[g] = like_f; l1.prior[g] = prior_f;
                l0.sv_like[1][g] = lc; l0.sv_prior[1][g] = prior_f;
                l0.acc_sub[g] = 0;
                l0.n_acc[g] = __ldcg(l0.n_acc + g) + nacc_c;
                l1.n_acc[g] = __ldcg(l1.n_acc + g) + nacc_f;
                p.ucur[g] = ucur;
                if (adaptive) p.win_sum[g] = wsum;
            }
            if (q.nb > 1) {
                __threadfence();
                tc::named_bar_sync(3 + t, 256);
                if (leader && lane == 0) tr_red_release_add(q.progress + pair, 1);
            }
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == TR_MMA_WARP0) tc::tmem_dealloc(tbase, 512);
}

// w = theta @ T^-1 in float64 (thread per chain), and the largest |w| (as the bits of a non-negative float)
__global__ void __launch_bounds__(128) tcr_whiten_kernel(const float* __restrict__ theta, const double* __restrict__ Tinv, float* __restrict__ wstate,
                                                         int d, int Cs, unsigned int* __restrict__ maxabs_bits) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= Cs) return;
    double th[TR_K];
#pragma unroll
    for (int j = 0; j < TR_K; j++) th[j] = (j < d) ? (double)theta[(size_t)j * Cs + c] : 0.0;
    float m = 0.0f;
    for (int k = 0; k < TR_K; k++) {
        double a = 0.0;
        if (k < d) {
#pragma unroll
            for (int j = 0; j < TR_K; j++) a = fma(th[j], __ldg(Tinv + (size_t)j * TR_K + k), a);
        }
        const float wf = (float)a;
        wstate[(size_t)k * Cs + c] = wf;
        m = fmaxf(m, fabsf(wf));
    }
    if (!(m <= 3.0e38f)) m = 3.0e38f;                     // NaN / inf -> "does not fit"
    atomicMax(maxabs_bits, __float_as_uint(m));
}

// power-of-two scale that brings `maxabs` into [2^(top-1), 2^top)
inline int tr_pow2_scale_for(double maxabs, int top) {
    if (!(maxabs > 0.0) || !std::isfinite(maxabs)) return 0;
    int e;
    std::frexp(maxabs, &e);
    return top - e;
}

// canonical K-major fp16 hi / lo images of B[n][k] = W[k][n0 + n] * 2^s   (W row-major [K][ldw], K = 64)
inline void tr_canon_split16(const std::vector<double>& W, int ldw, int n0, int rows, int s, __half* hi, __half* lo) {
    for (int n = 0; n < rows; n++)
        for (int k = 0; k < TR_K; k++) {
            const float x = (float)std::ldexp(W[(size_t)k * ldw + n0 + n], s);
            const __half h = __float2half_rn(x);
            const __half l = __float2half_rn(x - __half2float(h));
            const size_t o = tc::canon_offset_f16(n, k, TR_K) / 2;
            hi[o] = h;
            lo[o] = l;
        }
}

inline double tr_maxabs(const std::vector<double>& W) {
    double m = 0;
    for (double x : W) m = std::fmax(m, std::fabs(x));
    return m;
}

}  // namespace

struct DaTcrImpl {
    __half *dM = nullptr, *dChunks = nullptr;
    float *dNY = nullptr, *dW = nullptr;
    double* dTinv = nullptr;
    unsigned int* dMax = nullptr;
    int* dProgress = nullptr;
    int progress_len = 0;
    float w_limit = 0.0f;         // |w| must stay below this for the fp16 images
    DaTcrParams q{};
    void free_all() {
        if (dM) cudaFree(dM);
        if (dChunks) cudaFree(dChunks);
        if (dNY) cudaFree(dNY);
        if (dW) cudaFree(dW);
        if (dTinv) cudaFree(dTinv);
        if (dMax) cudaFree(dMax);
        if (dProgress) cudaFree(dProgress);
        dM = dChunks = nullptr;
        dNY = dW = nullptr;
        dTinv = nullptr; dMax = nullptr; dProgress = nullptr;
        progress_len = 0;
    }
};

bool DaTcrState::eligible(const tda_config& c, const Params<float>& P) const {
    if (c.dtype != TDA_F32 || c.n_levels != 2 || c.aem || c.randomize_subchain || c.mtm_k || c.prop_kind != TDA_PROP_PCN) return false;
    if (c.d > TR_K || c.d < 16 || (c.d % 16) != 0) return false;
    for (int l = 0; l < 2; l++)
        if (c.level[l].model_kind != TDA_MODEL_LINEAR || c.level[l].lik_kind > TDA_LIK_DENSE) return false;
    if (c.level[0].m < 1 || c.level[0].m > TR_MAX_MC) return false;
    if (c.level[1].m < 1 || c.level[1].m > TR_MAX_MF) return false;
    // Coarse Links are recorded as (whitened parameters, log-likelihood, accept flag); the engine turns the
    // parameters into theta = w T and rebuilds Link.prior / Link.model_output from them when they are first
    // fetched (EngineT::fill_lazy_history), so those fields need the parameters
    if ((c.level[0].store & (TDA_STORE_STATS | TDA_STORE_OUTPUT)) && !(c.level[0].store & TDA_STORE_THETA)) return false;
    if ((c.level[1].store & TDA_STORE_OUTPUT) && !(c.level[1].store & TDA_STORE_THETA)) return false;
    if ((P.Cs % 256) != 0) return false;
    if (!(c.scaling > 0.0 && c.scaling < 1.0)) return false;
    return true;
}

void DaTcrState::destroy() {
    if (impl) {
        impl->free_all();
        delete impl;
        impl = nullptr;
    }
    prepared = false;
    w_valid = false;
}

int DaTcrState::prepare(const Params<float>& P, const tda_config& c) {
    auto fetch = [&](const float* dev, size_t n, std::vector<double>& h) {
        std::vector<float> f(n);
        cudaError_t e = cudaMemcpy(f.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost);
        h.assign(f.begin(), f.end());
        return e;
    };
    const int d0 = c.d, mc0 = c.level[0].m, mf0 = c.level[1].m;
    const int mc = (mc0 + 15) / 16 * 16, mf = (mf0 + TR_CH - 1) / TR_CH * TR_CH;
    std::vector<double> T, LP, Ac, Af, bc, bf, dc, df, mu, sc;
    cudaError_t e = cudaSuccess;
    if (P.ldD < TR_K || P.lv[0].ldA < mc || P.lv[1].ldA < mf) { err = "tcr: operand leading dimensions"; return 1; }
    if (e == cudaSuccess) e = fetch(P.T, (size_t)d0 * P.ldD, T);
    if (e == cudaSuccess) e = fetch(P.LP, (size_t)d0 * P.ldD, LP);
    if (e == cudaSuccess) e = fetch(P.lv[0].A, (size_t)d0 * P.lv[0].ldA, Ac);
    if (e == cudaSuccess) e = fetch(P.lv[1].A, (size_t)d0 * P.lv[1].ldA, Af);
    if (e == cudaSuccess) e = fetch(P.lv[0].b, mc0, bc);
    if (e == cudaSuccess) e = fetch(P.lv[1].b, mf0, bf);
    if (e == cudaSuccess) e = fetch(P.lv[0].data, mc0, dc);
    if (e == cudaSuccess) e = fetch(P.lv[1].data, mf0, df);
    if (e == cudaSuccess) e = fetch(P.prior_mean, d0, mu);
    if (e == cudaSuccess) e = fetch(P.scaling, (size_t)P.Cs, sc);
    if (e != cudaSuccess) { err = std::string("tcr prepare: ") + cudaGetErrorString(e); return -2; }
    for (int i = 0; i < P.C; i++)
        if (!(sc[i] > 0.0 && sc[i] < 1.0)) { err = "tcr: pCN step outside (0, 1)"; return 1; }
    if (c.adaptive && (!P.win || !P.win_sum)) { err = "tcr: adaptation window missing"; return 1; }
    const int ldD = P.ldD, ldc = P.lv[0].ldA, ldf = P.lv[1].ldA;
    // Diagonal and dense Gaussian likelihoods are folded into the operators (see tda_da_tc16.cuh):
    // -0.5 r^T prec r = -0.5 |L^T r|^2 with prec = L L^T
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
            if (e != cudaSuccess) { err = std::string("tcr prepare: ") + cudaGetErrorString(e); return -2; }
            for (int n = 0; n < m0; n++) {
                if (!(var[n] > 0.0)) { err = "tcr: non-positive likelihood variance"; return 1; }
                const double wgt = 1.0 / std::sqrt(var[n]);
                for (int k = 0; k < d0; k++) A[(size_t)k * ld + n] *= wgt;
                r[n] *= wgt;
            }
        } else {
            std::vector<double> Lc;
            e = fetch(P.lv[l].prec, (size_t)m0 * m0, Lc);
            if (e != cudaSuccess) { err = std::string("tcr prepare: ") + cudaGetErrorString(e); return -2; }
            for (int j = 0; j < m0; j++) {            // in-place lower Cholesky factor of the precision
                double dj = Lc[(size_t)j * m0 + j];
                for (int k = 0; k < j; k++) dj -= Lc[(size_t)j * m0 + k] * Lc[(size_t)j * m0 + k];
                if (!(dj > 0.0)) { err = "tcr: likelihood precision is not positive definite in float32"; return 1; }
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
    // ---- T^-1 (Gauss-Jordan with partial pivoting, float64) ----
    std::vector<double> Tm((size_t)d0 * d0), Ti((size_t)d0 * d0, 0.0);
    for (int i = 0; i < d0; i++)
        for (int j = 0; j < d0; j++) Tm[(size_t)i * d0 + j] = T[(size_t)i * ldD + j];
    for (int i = 0; i < d0; i++) Ti[(size_t)i * d0 + i] = 1.0;
    const double tnorm = tr_maxabs(Tm);
    for (int col = 0; col < d0; col++) {
        int piv = col;
        for (int r = col + 1; r < d0; r++)
            if (std::fabs(Tm[(size_t)r * d0 + col]) > std::fabs(Tm[(size_t)piv * d0 + col])) piv = r;
        if (!(std::fabs(Tm[(size_t)piv * d0 + col]) > 1e-12 * tnorm)) { err = "tcr: the proposal covariance factor is singular"; return 1; }
        if (piv != col)
            for (int j = 0; j < d0; j++) { std::swap(Tm[(size_t)piv * d0 + j], Tm[(size_t)col * d0 + j]); std::swap(Ti[(size_t)piv * d0 + j], Ti[(size_t)col * d0 + j]); }
        const double pv = 1.0 / Tm[(size_t)col * d0 + col];
        for (int j = 0; j < d0; j++) { Tm[(size_t)col * d0 + j] *= pv; Ti[(size_t)col * d0 + j] *= pv; }
        for (int r = 0; r < d0; r++) {
            if (r == col) continue;
            const double f = Tm[(size_t)r * d0 + col];
            if (f == 0.0) continue;
            for (int j = 0; j < d0; j++) { Tm[(size_t)r * d0 + j] -= f * Tm[(size_t)col * d0 + j]; Ti[(size_t)r * d0 + j] -= f * Ti[(size_t)col * d0 + j]; }
        }
    }
    std::vector<double> Tinv64((size_t)TR_K * TR_K, 0.0);
    for (int i = 0; i < d0; i++)
        for (int j = 0; j < d0; j++) Tinv64[(size_t)i * TR_K + j] = Ti[(size_t)i * d0 + j];
    // ---- operators composed with T: X[k][n] = sum_j T[k][j] W[j][n]  (k, j < d0; rows d0..63 zero) ----
    auto compose = [&](const std::vector<double>& W, int ldw, int cols, int cols_pad) {
        std::vector<double> X((size_t)TR_K * cols_pad, 0.0);
        for (int k = 0; k < d0; k++)
            for (int n = 0; n < cols; n++) {
                double s = 0;
                for (int j = 0; j < d0; j++) s += T[(size_t)k * ldD + j] * W[(size_t)j * ldw + n];
                X[(size_t)k * cols_pad + n] = s;
            }
        return X;
    };
    std::vector<double> Mc = compose(Ac, ldc, mc0, TR_MAX_MC);       // refresh chunks always span 128 columns
    std::vector<double> Mf = compose(Af, ldf, mf0, mf);
    std::vector<double> Mp = compose(LP, ldD, d0, TR_K);
    std::vector<double> Mt((size_t)TR_K * TR_K, 0.0);
    for (int k = 0; k < d0; k++)
        for (int n = 0; n < d0; n++) Mt[(size_t)k * TR_K + n] = T[(size_t)k * ldD + n];
    // ---- scales: |w| < 32 (a chain whose whitened state is larger than that is 32 prior standard
    // deviations out), operators to [2^13, 2^14) ----
    const int s_w = 10;
    const int s_M = tr_pow2_scale_for(tr_maxabs(Mc), 14);
    const int s_Mf = tr_pow2_scale_for(tr_maxabs(Mf), 14);
    const int s_Mp = tr_pow2_scale_for(tr_maxabs(Mp), 14);
    const int s_Mt = tr_pow2_scale_for(tr_maxabs(Mt), 14);
    auto bad = [](int s) { return s < -20 || s > 40; };
    if (bad(s_M) || bad(s_Mf) || bad(s_Mp) || bad(s_Mt)) { err = "tcr: operands do not fit the fp16 range"; return 1; }

    const int nfc = mf / TR_CH, nf1 = nfc + 2, nch = nf1 + 2;
    const size_t chunk_halves = (size_t)2 * TR_CH * TR_K;
    std::vector<__half> hM((size_t)2 * mc * TR_K), hC((size_t)nch * chunk_halves);
    tr_canon_split16(Mc, TR_MAX_MC, 0, mc, s_M, hM.data(), hM.data() + (size_t)mc * TR_K);
    auto put_chunk = [&](int idx, const std::vector<double>& W, int ldw, int n0, int s) {
        tr_canon_split16(W, ldw, n0, TR_CH, s, hC.data() + (size_t)idx * chunk_halves, hC.data() + (size_t)idx * chunk_halves + (size_t)TR_CH * TR_K);
    };
    put_chunk(0, Mc, TR_MAX_MC, 0, s_M);
    put_chunk(1, Mc, TR_MAX_MC, TR_CH, s_M);
    for (int ch = 0; ch < nfc; ch++) put_chunk(2 + ch, Mf, mf, ch * TR_CH, s_Mf);
    put_chunk(2 + nfc, Mp, TR_K, 0, s_Mp);
    put_chunk(3 + nfc, Mt, TR_K, 0, s_Mt);

    const bool had_w = impl && impl->dW;
    if (!impl) impl = new DaTcrImpl();
    float* keepW = impl->dW;                  // the whitened state survives a re-prepare (step sizes changed)
    impl->dW = nullptr;
    impl->free_all();
    impl->dW = keepW;
    if (e == cudaSuccess) e = cudaMalloc(&impl->dM, hM.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&impl->dChunks, hC.size() * 2);
    if (e == cudaSuccess && !impl->dW) e = cudaMalloc(&impl->dW, (size_t)TR_K * P.Cs * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&impl->dTinv, Tinv64.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&impl->dMax, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemcpy(impl->dM, hM.data(), hM.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(impl->dChunks, hC.data(), hC.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(impl->dTinv, Tinv64.data(), Tinv64.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { err = std::string("tcr prepare: ") + cudaGetErrorString(e); return -2; }
    std::vector<float> ny((size_t)TR_MAX_MC + TR_K + mf, 0.f);
    for (int j = 0; j < mc0; j++) ny[j] = -(float)rc[j];
    for (int n = 0; n < d0; n++) {
        double s = 0;
        for (int k = 0; k < d0; k++) s += mu[k] * LP[(size_t)k * ldD + n];
        ny[TR_MAX_MC + n] = -(float)s;
    }
    for (int j = 0; j < mf0; j++) ny[TR_MAX_MC + TR_K + j] = -(float)rf[j];
    e = cudaMalloc(&impl->dNY, ny.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(impl->dNY, ny.data(), ny.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { err = std::string("tcr prepare: ") + cudaGetErrorString(e); return -2; }
    DaTcrParams& q = impl->q;
    q.M_hl = impl->dM; q.chunks = impl->dChunks; q.ny = impl->dNY; q.wstate = impl->dW;
    q.mc = mc; q.mf = mf; q.nfc = nfc; q.n_f1 = nf1; q.n_chunks = nch; q.J = c.subchain[0];
    q.var_c = (float)lik_var[0]; q.var_f = (float)lik_var[1];
    q.prior_logconst = (float)c.prior_logconst;
    q.bz = (float)std::ldexp(1.0, s_w - 12);
    q.sc_c = (float)std::ldexp(1.0, -(s_w + s_M));
    q.sc_f = (float)std::ldexp(1.0, -(s_w + s_Mf));
    q.sc_p = (float)std::ldexp(1.0, -(s_w + s_Mp));
    q.sc_t = (float)std::ldexp(1.0, -(s_w + s_Mt));
    q.w_scale = (float)std::ldexp(1.0, s_w);
    q.w_unscale = (float)std::ldexp(1.0, -s_w);
    impl->w_limit = 32.0f;
    prepared = true;
    if (!had_w) w_valid = false;
    return 0;
}

int DaTcrState::ready(const Params<float>& P, const tda_config& c, cudaStream_t st) {
    if (!prepared) { int r = prepare(P, c); if (r) return r; }
    cudaError_t e = cudaSuccess;
    if (!w_valid) {
        // the finest level's theta (written by init or by another kernel) -> whitened state, checked against
        // the range of the fp16 images
        e = cudaMemsetAsync(impl->dMax, 0, sizeof(unsigned int), st);
        if (e == cudaSuccess) {
            tcr_whiten_kernel<<<(P.Cs + 127) / 128, 128, 0, st>>>(P.lv[1].theta, impl->dTinv, impl->dW, P.d, P.Cs, impl->dMax);
            e = cudaGetLastError();
        }
        unsigned int bits = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&bits, impl->dMax, sizeof(bits), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { err = std::string("tcr whiten: ") + cudaGetErrorString(e); return -2; }
        float m;
        memcpy(&m, &bits, sizeof(m));
        if (!(m < impl->w_limit)) { err = "tcr: a chain's whitened state lies outside the fp16 operand range (|w| >= 32)"; return 1; }
        w_valid = true;
    }
    return 0;
}

int DaTcrState::run(Params<float>& P, const tda_config& c, long long iterations, int sm_count, cudaStream_t st) {
    { int r = ready(P, c, st); if (r) return r; }
    DaTcrParams& q = impl->q;
    cudaError_t e = cudaSuccess;
    q.n_pairs = P.Cs / 256;
    const int grid = q.n_pairs < sm_count ? q.n_pairs : sm_count;
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
        if (impl->progress_len < q.n_pairs) {
            if (impl->dProgress) cudaFree(impl->dProgress);
            impl->dProgress = nullptr;
            if (cudaMalloc(&impl->dProgress, (size_t)q.n_pairs * sizeof(int)) != cudaSuccess) { err = "tcr: cudaMalloc progress"; return -3; }
            impl->progress_len = q.n_pairs;
        }
        e = cudaMemsetAsync(impl->dProgress, 0, (size_t)q.n_pairs * sizeof(int), st);
        if (e != cudaSuccess) { err = std::string("tcr progress: ") + cudaGetErrorString(e); return -2; }
    }
    q.progress = impl->dProgress;
    const size_t smem = TR_SMEM_BYTES;
    auto kern = (P.d == TR_K) ? da_tcr_kernel<false> : da_tcr_kernel<true>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { err = std::string("tcr attr: ") + cudaGetErrorString(e); return -2; }
    P.mode = MODE_RUN;
    P.iterations = iterations;
    P.z_round = 1;
    kern<<<grid, TR_THREADS, smem, st>>>(P, q);
    e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("tcr launch: ") + cudaGetErrorString(e); return -2; }
    return 0;
}

}  // namespace tda
