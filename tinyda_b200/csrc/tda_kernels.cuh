// Generic lock-step chain kernel: one launch advances every chain of the engine by
// `iterations` finest-level iterations of MH / Delayed Acceptance / MLDA, nested coarse
// subchains included, with no host round trip.  A CTA owns a tile of TC chains for the whole
// run; per-chain state lives in structure-of-arrays buffers (chain index fastest, so a warp
// reads 32 consecutive chains = one 128/256-byte line); the dense contractions (proposal
// factor, prior whitening, linear forward operator, Poisson log-conductivity field) are
// tile contractions [TC chains x K] @ [K x N] with the shared operand staged through shared
// memory by cp.async double buffering and the per-chain reductions (|w|^2, |F-data|^2)
// fused into the epilogue.
//
// Reference map (file:line into the reference's tinyDA/ package):
//   Tile::base_step        chain.py:101-125, :415-444; proposal.py:1583-1613
//   Tile::upper_step       chain.py:353-402, :708-765; proposal.py:1511-1578
//   Tile::align            proposal.py:1469-1493 (object identity -> state ids + saved versions)
//   Tile::eval_level       posterior.py:78-110 (create_link)
//   Tile::push_bias        chain.py:485-523, :740-765; proposal.py:1442-1467, :1548-1578;
//                          distributions.py:385-425
//   Tile::propose_*        proposal.py:247-251 (RWMH/AM), :349-355 (pCN), :811-852 (DREAMZ),
//                          :948-959 (MALA)
//   Tile::adapt            proposal.py:228-245, :502-512, :790-795; utils.py:113-124
//   Tile::sd_terms / aem_update_sd   state-dependent error model chain.py:446-473, :501-522;
//                          utils.py:190-201; distributions.py:427-446; proposal.py:364-369
//   Tile::draw_promoted / snapshot_promoted   randomize_subchain_length chain.py:369, :525-527
//   Tile::mtm_step / mtm_q / mtm_logsumexp    MultipleTry ray.py:279-354 (k candidate Links, the
//                          choice, k-1 reference Links, ratio of summed weights)
//   Tile::propose_gaussian (OWPCN branches)   OperatorWeightedCrankNicolson proposal.py:575-598
//   Tile::propose_dream + the crossover part of Tile::adapt   DREAMZ proposal.py:790-852
//   Tile::loglike_adaptive_tile, the cooperative factorisation in Tile::push_bias
//                          AdaptiveGaussianLogLike distributions.py:385-446
#pragma once
#include "tda_common.cuh"

namespace tda {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <typename R> struct VecB;
template <> struct VecB<float> {
    static constexpr int W = 4;
    typedef float4 T;
    static __device__ __forceinline__ void unpack(const T& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
template <> struct VecB<double> {
    static constexpr int W = 2;
    typedef double2 T;
    static __device__ __forceinline__ void unpack(const T& v, double* o) { o[0] = v.x; o[1] = v.y; }
};

// -------------------------------------------------------------------------------------------
// Tile contraction: out[c][n] = sum_k (As[k][c] - sub[k]) * Bg[k][n],  c < TC, n < N.
// Thread (lane, warp): chains c0 = rg*64+lane, c1 = c0+32 (rg = warp/4), columns
// n0 + cw*RN + j (cw = warp%4).  `epi(r, c, n, value)` is called once per output.
// -------------------------------------------------------------------------------------------
template <typename R, typename Epi>
__device__ __forceinline__ void tile_gemm(const R* __restrict__ As, const R* __restrict__ sub,
                                          const R* __restrict__ Bg, int K, int N, int ldb,
                                          R* __restrict__ bs, int KB, Epi& epi) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int cw = w & (CW - 1), rg = w >> 2;
    const int c0 = rg * 64 + lane, c1 = c0 + 32;
    const int nchunks = (N + NB - 1) / NB;
    constexpr int VW = VecB<R>::W;
    constexpr int VPR = NB / VW;          // 16-byte vectors per staged row
    const int total = K * VPR;
    // prefetch chunk 0
    for (int v = tid; v < total; v += NT) {
        int k = v / VPR, j = v - k * VPR;
        cp_async16(bs + k * NB + j * VW, Bg + (size_t)k * ldb + j * VW);
    }
    cp_async_commit();
    for (int ch = 0; ch < nchunks; ch++) {
        if (ch + 1 < nchunks) {
            R* dst = bs + ((ch + 1) & 1) * KB * NB;
            const R* src = Bg + (size_t)(ch + 1) * NB;
            for (int v = tid; v < total; v += NT) {
                int k = v / VPR, j = v - k * VPR;
                cp_async16(dst + k * NB + j * VW, src + (size_t)k * ldb + j * VW);
            }
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        R acc0[RN], acc1[RN];
#pragma unroll
        for (int j = 0; j < RN; j++) { acc0[j] = (R)0; acc1[j] = (R)0; }
        const R* bsc = bs + (ch & 1) * KB * NB + cw * RN;
#pragma unroll 2
        for (int k = 0; k < K; k++) {
            R a0 = As[k * TC + c0], a1 = As[k * TC + c1];
            if (sub != nullptr) { R s = sub[k]; a0 -= s; a1 -= s; }
            R b[RN];
            const typename VecB<R>::T* bv = reinterpret_cast<const typename VecB<R>::T*>(bsc + k * NB);
#pragma unroll
            for (int j = 0; j < RN / VW; j++) VecB<R>::unpack(bv[j], b + j * VW);
#pragma unroll
            for (int j = 0; j < RN; j++) {
                acc0[j] = fma(a0, b[j], acc0[j]);
                acc1[j] = fma(a1, b[j], acc1[j]);
            }
        }
        const int nbase = ch * NB + cw * RN;
#pragma unroll
        for (int j = 0; j < RN; j++) {
            int n = nbase + j;
            if (n < N) {
                epi(0, c0, n, acc0[j]);
                epi(1, c1, n, acc1[j]);
            }
        }
        __syncthreads();
    }
}

// ---- epilogues -----------------------------------------------------------------------------
template <typename R>
struct EpiSsq {            // sum of squares per chain (prior whitening)
    R ps[RM];
    __device__ EpiSsq() { ps[0] = ps[1] = (R)0; }
    __device__ __forceinline__ void operator()(int r, int, int, R v) { ps[r] = fma(v, v, ps[r]); }
};

template <typename R>
struct EpiCombine {        // theta' = a_c * theta_c + b_c * xi   -> proposal tile
    const R* theta; int Cs; int chain0; R* pt; const R* ca; const R* cb;
    __device__ __forceinline__ void operator()(int, int c, int n, R v) {
        pt[n * TC + c] = ca[c] * theta[(size_t)n * Cs + chain0 + c] + cb[c] * v;
    }
};

template <typename R>
struct EpiAffine {         // independence proposal: pt[n][c] = mu[n] + v
    R* pt; const R* mu;
    __device__ __forceinline__ void operator()(int, int c, int n, R v) { pt[n * TC + c] = mu[n] + v; }
};

template <typename R>
struct EpiLinear {         // F = v + b; optional store; residual reduction for ISO / DIAG
    const R* b; const R* data; const R* var; R* Fp; int Cs; int chain0; int lik_kind; int need_F;
    R ps[RM];
    __device__ __forceinline__ void operator()(int r, int c, int n, R v) {
        R F = v + b[n];
        if (need_F) Fp[(size_t)n * Cs + chain0 + c] = F;
        R res = F - data[n];
        if (lik_kind == TDA_LIK_ISO) ps[r] = fma(res, res, ps[r]);
        else if (lik_kind == TDA_LIK_DIAG) ps[r] += res * res / var[n];
    }
};

template <typename R>
struct EpiExpStore {       // Poisson: inverse conductivity field 1 / k = exp(-Phi theta) -> scratch
    R* kf; int Cs; int chain0;
    __device__ __forceinline__ void operator()(int, int c, int n, R v) {
        kf[(size_t)n * Cs + chain0 + c] = texp(-v);
    }
};

template <typename R>
struct EpiStoreNeg {       // out[n][c] (+)= -v   (MALA prior gradient P (mu - x))
    R* out; int Cs; int chain0; int accumulate;
    __device__ __forceinline__ void operator()(int, int c, int n, R v) {
        size_t i = (size_t)n * Cs + chain0 + c;
        out[i] = accumulate ? out[i] - v : -v;
    }
};

template <typename R>
struct EpiTile {           // shared-memory tile: t[n][c] (+)= v
    R* t; int accumulate;
    __device__ __forceinline__ void operator()(int, int c, int n, R v) {
        t[n * TC + c] = accumulate ? t[n * TC + c] + v : v;
    }
};

template <typename R>
struct EpiOw {             // adaptive OWPCN: t[n][c] = sqrt(s_c lam_n) t[n][c] + sqrt(1 - s_c lam_n) v
    R* t; const R* s; const R* lam;
    __device__ __forceinline__ void operator()(int, int c, int n, R v) {
        const R sl = s[c] * lam[n];
        t[n * TC + c] = tsqrt(sl) * t[n * TC + c] + tsqrt((R)1 - sl) * v;
    }
};

template <typename R>
struct EpiStore {          // out[n][c] = v
    R* out; int Cs; int chain0;
    __device__ __forceinline__ void operator()(int, int c, int n, R v) {
        out[(size_t)n * Cs + chain0 + c] = v;
    }
};

// -------------------------------------------------------------------------------------------
template <typename R>
struct Tile {
    const Params<R>& p;
    R *zt, *pt, *bs, *red, *s_prior, *s_like, *s_ca, *s_cb;
    int* s_acc;
    int* s_inv;        // [TC] chains whose bias covariance must be re-factorised (push_bias)
    R* aem_w;          // [m*m][aem_G] workspace of the cooperative Cholesky (adaptive error model)
    R* s_r;            // [m][TC] residual tile of the adaptive likelihood
    int KB;            // rows of a staged operand chunk
    int chain0;        // first chain of this tile
    int tid;
    // run-local counters (uniform across chains)
    long long t_base, wcount, rec[MAXL], lvl_steps[MAXL], slots;
    long long mtm_zoff;          // MultipleTry: normals drawn so far within the current step
    const R* prop_src;           // state the Gaussian proposals start from (default: level 0's current state)

    __device__ Tile(const Params<R>& p_, unsigned char* smem, int kt) : p(p_) {
        tid = threadIdx.x;
        KB = kt;
        zt = reinterpret_cast<R*>(smem);
        pt = zt + kt * TC;
        bs = pt + kt * TC;
        red = bs + 2 * kt * NB;
        s_prior = red + CW * TC;
        s_like = s_prior + TC;
        s_ca = s_like + TC;
        s_cb = s_ca + TC;
        s_acc = reinterpret_cast<int*>(s_cb + TC);
        s_inv = s_acc + TC;
        s_r = reinterpret_cast<R*>(s_inv + TC);
        aem_w = s_r + p_.m_adapt * TC;
    }

    __device__ __forceinline__ size_t gi(int k, int c) const { return (size_t)k * p.Cs + chain0 + c; }

    // ---- random streams --------------------------------------------------------------------
    __device__ __forceinline__ R normal_at(int c, long long idx) const {
        if (p.rng_mode == TDA_RNG_INJECTED) {
            int ch = chain0 + c;
            return (ch < p.C && idx < p.zlen) ? p.zs[(size_t)ch * p.zlen + idx] : (R)0;
        }
        return philox_normal<R>(p.seed, p.chain_offset + chain0 + c, idx, p.z_round);
    }
    __device__ __forceinline__ R uniform_at(int c, long long idx) const {
        if (p.rng_mode == TDA_RNG_INJECTED) {
            int ch = chain0 + c;
            return (ch < p.C && idx < p.ulen) ? p.us[(size_t)ch * p.ulen + idx] : (R)0.5;
        }
        return philox_uniform<R>(p.seed, p.chain_offset + chain0 + c, idx);
    }

    // d normals per chain for this base step -> zt[k][c]
    __device__ void fill_normals() {
        const int d = p.d;
        if (p.mtm_k) {      // per-chain cursor: offset by the draws this chain has consumed so far
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                zt[k * TC + c] = normal_at(c, p.zcur[chain0 + c] + mtm_zoff + k);
            }
            return;
        }
        const long long z0 = t_base * d;
        if (p.rng_mode == TDA_RNG_PHILOX && (z0 & 3) == 0 && !p.z_round) {
            const int nb4 = (d + 3) >> 2;
            for (int e = tid; e < nb4 * TC; e += NT) {
                int q = e / TC, c = e - q * TC;
                R v[4];
                normals4<R>(philox_block(p.seed, p.chain_offset + chain0 + c, STREAM_Z, (unsigned long long)(z0 >> 2) + q), v);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (4 * q + i < d) zt[(4 * q + i) * TC + c] = v[i];
            }
        } else {
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                zt[k * TC + c] = normal_at(c, z0 + k);
            }
        }
    }

    // cross-column-warp reduction of per-thread partial sums; result valid for tid < TC
    __device__ __forceinline__ R reduce_cols(R ps0, R ps1) {
        const int lane = tid & 31, w = tid >> 5, cw = w & (CW - 1), rg = w >> 2;
        red[cw * TC + rg * 64 + lane] = ps0;
        red[cw * TC + rg * 64 + 32 + lane] = ps1;
        __syncthreads();
        R tot = (R)0;
        if (tid < TC) tot = (red[tid] + red[TC + tid]) + (red[2 * TC + tid] + red[3 * TC + tid]);
        __syncthreads();
        return tot;
    }

    // ---- likelihoods needing the whole output vector (thread per chain) ----------------------
    __device__ R loglike_from_F(int l, const R* Fsrc, int c) const {
        const LevelP<R>& v = p.lv[l];
        const int m = v.m;
        if (v.lik_kind == TDA_LIK_ISO || v.lik_kind == TDA_LIK_DIAG) {
            R s = (R)0;
            for (int j = 0; j < m; j++) {
                R r = Fsrc[gi(j, c)] - v.data[j];
                s += (v.lik_kind == TDA_LIK_ISO) ? r * r : r * r / v.var[j];
            }
            return (v.lik_kind == TDA_LIK_ISO) ? (R)-0.5 * s / v.lik_var : (R)-0.5 * s;
        }
        R q = (R)0;
        if (v.lik_kind == TDA_LIK_DENSE) {
            for (int i = 0; i < m; i++) {
                R ri = Fsrc[gi(i, c)] - v.data[i];
                R t = (R)0;
                for (int j = 0; j < m; j++) t = fma(v.prec[i * m + j], Fsrc[gi(j, c)] - v.data[j], t);
                q = fma(ri, t, q);
            }
        } else {   // ADAPTIVE: r = F + bias - data; r^T inv(cov + cov_bias) r = |Li r|^2 with the
                   // per-chain inverse Cholesky factor Li (lower triangular) kept in lik_prec
            for (int i = 0; i < m; i++) {
                R t = (R)0;
                for (int j = 0; j <= i; j++)
                    t = fma(v.lik_prec[gi(i * m + j, c)], Fsrc[gi(j, c)] + v.lik_bias[gi(j, c)] - v.data[j], t);
                q = fma(t, t, q);
            }
        }
        return (R)-0.5 * q;
    }

    // AdaptiveGaussianLogLike for the whole tile, all NT threads: r = F + bias - data staged in
    // shared memory, then |Li r|^2 with the per-chain triangular factor, the rows of Li dealt
    // alternately to the two threads of a chain.  out[c] (shared) valid after the call.
    __device__ void loglike_adaptive_tile(int l, const R* Fsrc, R* out) {
        const LevelP<R>& v = p.lv[l];
        const int m = v.m;
        for (int e = tid; e < m * TC; e += NT) {
            int j = e / TC, c = e - j * TC;
            s_r[e] = Fsrc[gi(j, c)] + v.lik_bias[gi(j, c)] - v.data[j];
        }
        __syncthreads();
        const int c = tid & (TC - 1), h = tid / TC;
        const R* __restrict__ Li = v.lik_prec + chain0 + c;
        const size_t Cs = (size_t)p.Cs;
        R q = (R)0;
        for (int i = h; i < m; i += NT / TC) {
            const R* row = Li + (size_t)i * m * Cs;
            R t0 = (R)0, t1 = (R)0;
            int j = 0;
            for (; j + 1 <= i; j += 2) {
                t0 = fma(row[(size_t)j * Cs], s_r[j * TC + c], t0);
                t1 = fma(row[(size_t)(j + 1) * Cs], s_r[(j + 1) * TC + c], t1);
            }
            if (j <= i) t0 = fma(row[(size_t)j * Cs], s_r[j * TC + c], t0);
            const R t = t0 + t1;
            q = fma(t, t, q);
        }
        red[h * TC + c] = q;
        __syncthreads();
        if (tid < TC) {
            R tot = (R)0;
            for (int k = 0; k < NT / TC; k++) tot += red[k * TC + tid];
            out[tid] = (R)-0.5 * tot;
        }
        __syncthreads();
    }

    // ---- create_link: log-prior + forward model + log-likelihood for the tile in pt ----------
    // results: s_prior[c], s_like[c] (shared), lv[l].Fp (global) when need_F
    __device__ void eval_level(int l) {
        const LevelP<R>& v = p.lv[l];
        const int d = p.d;
        {   // prior: -0.5*(logconst + |(x-mu) LP|^2)
            EpiSsq<R> e;
            tile_gemm<R>(pt, p.prior_mean, p.LP, d, d, p.ldD, bs, KB, e);
            R tot = reduce_cols(e.ps[0], e.ps[1]);
            if (tid < TC) s_prior[tid] = (R)-0.5 * (p.prior_logconst + tot);
        }
        if (v.model_kind == TDA_MODEL_LINEAR) {
            EpiLinear<R> e;
            e.b = v.b; e.data = v.data; e.var = v.var; e.Fp = v.Fp; e.Cs = p.Cs; e.chain0 = chain0;
            e.lik_kind = v.lik_kind; e.need_F = v.need_F; e.ps[0] = e.ps[1] = (R)0;
            tile_gemm<R>(pt, (const R*)nullptr, v.A, d, v.m, v.ldA, bs, KB, e);
            R tot = reduce_cols(e.ps[0], e.ps[1]);
            if (tid < TC) {
                if (v.lik_kind == TDA_LIK_ISO) s_like[tid] = (R)-0.5 * tot / v.lik_var;
                else if (v.lik_kind == TDA_LIK_DIAG) s_like[tid] = (R)-0.5 * tot;
            }
            if (v.lik_kind >= TDA_LIK_DENSE) {
                __syncthreads();     // Fp visible to the owning thread (block-scope global writes)
                if (v.lik_kind == TDA_LIK_ADAPTIVE) loglike_adaptive_tile(l, v.Fp, s_like);
                else if (tid < TC) s_like[tid] = loglike_from_F(l, v.Fp, tid);
            }
        } else if (v.model_kind == TDA_MODEL_ROSENBROCK) {
            if (tid < TC) {
                R x = pt[tid], y = pt[TC + tid];
                R a = v.sc0 - x, b = y - x * x;
                R F = a * a + v.sc1 * b * b;
                v.Fp[gi(0, tid)] = F;
                if (v.lik_kind != TDA_LIK_ADAPTIVE) s_like[tid] = loglike_from_F(l, v.Fp, tid);
            }
            if (v.lik_kind == TDA_LIK_ADAPTIVE) { __syncthreads(); loglike_adaptive_tile(l, v.Fp, s_like); }
        } else {   // POISSON1D
            const int n = v.n_grid;
            R* kf = p.scratch;
            R* cp = p.scratch + (size_t)p.n_max * p.Cs;
            EpiExpStore<R> e;
            e.kf = kf; e.Cs = p.Cs; e.chain0 = chain0;
            tile_gemm<R>(pt, (const R*)nullptr, v.A, d, n, v.ldA, bs, KB, e);
            __syncthreads();
            if (tid < TC) {
                // -(k u')' = 1, u(0) = u(1) = 0 in one dimension: the flux in cell i is q_i = q_0 - i h^2, so u at node j
                // is sum_{i<j} q_i / k_i and u(1) = 0 fixes q_0 = h^2 sum(i / k_i) / sum(1 / k_i).  Algebraically the
                // solution of the tridiagonal system models.py solves with the Thomas algorithm, but two running sums
                // instead of a division recurrence whose pivots cancel: in float32 the Thomas sweeps lose 3-4 digits
                // on a 512-cell grid (2.6e-4 relative against float64), these sums stay at rounding level.  The cell
                // index is centred (i - n/2) to keep the two terms of u small.
                const int c = tid;
                const size_t Cs = (size_t)p.Cs;
                const R* __restrict__ rfc = kf + chain0 + c;          // 1 / k_i
                R* __restrict__ p1c = cp + chain0 + c;                // prefix sums of (i - n/2) / k_i at the sensors
                const R h2 = (R)1 / ((R)n * (R)n), c0 = (R)(n / 2);
                const int stride = v.stride;
                R P0 = (R)0, P1 = (R)0;
                int s = 0, next = stride;
#pragma unroll 4
                for (int i = 0; i < n; i++) {
                    const R r = rfc[(size_t)i * Cs];
                    P0 += r;
                    P1 = fma((R)i - c0, r, P1);
                    if (i + 1 == next) {
                        if (s < v.m) { v.Fp[gi(s, c)] = P0; p1c[(size_t)s * Cs] = P1; }
                        s++;
                        next += stride;
                    }
                }
                const R q = P1 / P0;
                for (int j = 0; j < v.m; j++) v.Fp[gi(j, c)] = h2 * (q * v.Fp[gi(j, c)] - p1c[(size_t)j * Cs]);
                if (v.lik_kind != TDA_LIK_ADAPTIVE) s_like[tid] = loglike_from_F(l, v.Fp, tid);
            }
            if (v.lik_kind == TDA_LIK_ADAPTIVE) { __syncthreads(); loglike_adaptive_tile(l, v.Fp, s_like); }
        }
        __syncthreads();
    }

    // ---- adaptive error model ------------------------------------------------------------
    // level l (>=1) pushes bias moments into level l-1's likelihood and re-scores level l-1's
    // current link.  Thread per chain.
    __device__ void push_bias(int l) {
        const int L = p.L;
        const LevelP<R>& lo = p.lv[l - 1];
        const int m = lo.m;
        const int kend = (l == L - 1) ? l : L - 1;
        // (1) bias mean; does any entry of the summed bias covariance reach 1e-9 (distributions.py:399)?
        for (int e = tid; e < m * TC; e += NT) {
            int i = e / TC, c = e - i * TC;
            R mu = (R)0;
            if (p.aem == 2) mu = p.lv[l].model_diff[gi(i, c)];     // chain.py:296-300, :519-522
            else for (int k = l; k <= kend; k++) mu += p.lv[k].bias_mu[gi(i, c)];
            lo.lik_bias[gi(i, c)] = mu;
        }
        if (tid < TC) s_inv[tid] = 0;
        __syncthreads();
        // (2) W = cov + cov_bias = Lc Lc^T, Li = inv(Lc) -> lik_prec.  G chains at a time in shared
        // memory, NT/G threads per chain: the rows of W, the trailing rows of a Cholesky column and
        // the columns of the triangular inverse are dealt round-robin to a chain's threads.  The
        // "is any entry of cov_bias >= 1e-9" test of set_bias (distributions.py:399) rides on the
        // assembly of W (the matrices are symmetric: the lower triangle decides).
        const int G = p.aem_G, NW = NT / G;
        const int cg = tid % G, wk = tid / G;
        R* Wm = aem_w;                 // [m*m][G]
        R* dg = aem_w + m * m * G;     // [m][G] diagonal of Lc
        for (int g0 = 0; g0 < TC; g0 += G) {
            const int c = g0 + cg;
            {
                int big = 0;
                for (int i = wk; i < m; i += NW)
                    for (int j = 0; j <= i; j++) {
                        R sg = (R)0;
                        for (int k = l; k <= kend; k++) sg += p.lv[k].bias_sigma[gi(i * m + j, c)];
                        if (!(sg < (R)1e-9)) big = 1;
                        Wm[(i * m + j) * G + cg] = lo.cov[i * m + j] + sg;
                    }
                if (big) s_inv[c] = 1;
            }
            __syncthreads();
            const bool on = s_inv[c] != 0;
            for (int j = 0; j < m; j++) {
                if (on) {
                    R djj = Wm[(j * m + j) * G + cg];
                    for (int k = 0; k < j; k++) { R t = Wm[(j * m + k) * G + cg]; djj -= t * t; }
                    // a non-positive pivot (near-singular bias covariance in this dtype) is clamped and
                    // flagged; the reference's np.linalg.inv does not fail there (distributions.py:402)
                    if (!(djj > (R)0)) { djj = (R)1e-30; if (wk == 0) *p.error_flag = 1; }
                    djj = tsqrt(djj);
                    const R inv = (R)1 / djj;
                    if (wk == 0) dg[j * G + cg] = djj;
                    for (int i = j + 1 + wk; i < m; i += NW) {
                        R sv = Wm[(i * m + j) * G + cg];
                        for (int k = 0; k < j; k++) sv -= Wm[(i * m + k) * G + cg] * Wm[(j * m + k) * G + cg];
                        Wm[(i * m + j) * G + cg] = sv * inv;
                    }
                }
                __syncthreads();
            }
            if (on) {
                R* Li = lo.lik_prec;
                for (int j = wk; j < m; j += NW) {     // column j of inv(Lc); a thread re-reads only its own writes
                    Li[gi(j * m + j, c)] = (R)1 / dg[j * G + cg];
                    for (int i = j + 1; i < m; i++) {
                        R sv = (R)0;
                        for (int k = j; k < i; k++) sv -= Wm[(i * m + k) * G + cg] * Li[gi(k * m + j, c)];
                        Li[gi(i * m + j, c)] = sv / dg[i * G + cg];
                    }
                }
            }
            __syncthreads();
        }
        // (3) re-score level l-1's current link (posterior.py:112-134)
        loglike_adaptive_tile(l - 1, lo.F, s_like);
        if (tid < TC) {
            const int c = tid;
            R nl = s_like[c];
            lo.like[chain0 + c] = nl;
            const int sid = lo.sid[chain0 + c];
            for (int a = l; a < L; a++)
                if (p.lv[a].sid[chain0 + c] == sid) lo.sv_like[a][chain0 + c] = nl;
        }
        __syncthreads();
    }

    __device__ void aem_update(int l, long long tcount) {
        // bias.update(model_diff)  (utils.py:113-124), t = tcount (starts at 1); element-parallel
        const LevelP<R>& v = p.lv[l];
        const int m = v.m;
        for (int e = tid; e < m * TC; e += NT) {
            int j = e / TC, c = e - j * TC;
            if (s_acc[c]) v.model_diff[gi(j, c)] = v.F[gi(j, c)] - p.lv[l - 1].F[gi(j, c)];
        }
        __syncthreads();
        const R t = (R)tcount;
        const R f1 = (t - (R)1) / t, f2 = (R)1 / t, g1 = (R)1 / (t + (R)1);
        {
            const R* __restrict__ md = v.model_diff;
            const R* __restrict__ bm = v.bias_mu;
            R* __restrict__ sg = v.bias_sigma;
            for (int e = tid; e < m * m * TC; e += NT) {
                int c = e & (TC - 1), ij = e / TC, i = ij / m, j = ij - i * m;
                R xi = md[gi(i, c)], xj = md[gi(j, c)];
                R mpi = bm[gi(i, c)], mpj = bm[gi(j, c)];
                R mni = g1 * (t * mpi + xi), mnj = g1 * (t * mpj + xj);
                size_t o = gi(ij, c);
                sg[o] = f1 * sg[o] + f2 * (t * mpi * mpj - (t + (R)1) * mni * mnj + xi * xj);
            }
        }
        __syncthreads();
        for (int e = tid; e < m * TC; e += NT) {
            int i = e / TC, c = e - i * TC;
            v.bias_mu[gi(i, c)] = g1 * (t * v.bias_mu[gi(i, c)] + v.model_diff[gi(i, c)]);
        }
        __syncthreads();
    }

    // state-dependent error model (two levels): ZeroMeanRecursiveSampleMoments.update with the
    // difference corrected by the previous offset, then the offset itself (chain.py:501-522)
    __device__ void aem_update_sd(int l, long long tcount) {
        if (tid < TC) {
            const int c = tid;
            const LevelP<R>& v = p.lv[l];
            const LevelP<R>& lo = p.lv[l - 1];
            const int m = v.m;
            R* corr = v.bias_mu;      // scratch: the mean is not used by this error model
            for (int j = 0; j < m; j++) {
                R ff = v.F[gi(j, c)], fc = lo.F[gi(j, c)];
                corr[gi(j, c)] = ff - (fc + v.model_diff[gi(j, c)]);
                v.model_diff[gi(j, c)] = ff - fc;
            }
            const R t = (R)tcount;
            const R f1 = (t - (R)1) / t, f2 = (R)1 / t;
            for (int i = 0; i < m; i++) {
                R xi = corr[gi(i, c)];
                for (int j = 0; j < m; j++) {
                    size_t e = gi(i * m + j, c);
                    v.bias_sigma[e] = f1 * v.bias_sigma[e] + f2 * (xi * corr[gi(j, c)]);
                }
            }
        }
        __syncthreads();
    }

    // state-dependent second stage: q(x,y) = log N(y; sqrt(1-s^2) x, s^2 C) up to the constant
    // that every term of chain.py:463-471 shares -> s_ca = q_x_y, s_cb = q_y_x (0 if symmetric).
    // x = the level's current state, y = the promoted state in pt.
    __device__ void sd_terms(int l) {
        const int d = p.d;
        const LevelP<R>& v = p.lv[l];
        if (p.prop_kind != TDA_PROP_PCN) {
            if (tid < TC) { s_ca[tid] = (R)0; s_cb[tid] = (R)0; }
            __syncthreads();
            return;
        }
        for (int pass = 0; pass < 2; pass++) {
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                R s = p.scaling[chain0 + c];
                R a = tsqrt((R)1 - s * s);
                R x = v.theta[gi(k, c)], y = pt[e];
                zt[e] = pass == 0 ? y - a * x : x - a * y;
            }
            __syncthreads();
            EpiSsq<R> e;
            tile_gemm<R>(zt, (const R*)nullptr, p.LP, d, d, p.ldD, bs, KB, e);
            R tot = reduce_cols(e.ps[0], e.ps[1]);
            if (tid < TC) {
                R s = p.scaling[chain0 + tid];
                R q = (R)-0.5 * tot / (s * s);
                if (pass == 0) s_ca[tid] = q; else s_cb[tid] = q;
            }
            __syncthreads();
        }
    }

    // ---- randomize_subchain_length (two levels) -------------------------------------------------
    // np.random.randint(-J, 0) is drawn after the subchain (chain.py:369) from the uniform that
    // follows the J accept-test uniforms; every base step of the proposals allowed here consumes
    // exactly one uniform, so that value can be read before the subchain starts and the chosen
    // link snapshotted when the subchain passes it.
    __device__ void draw_promoted() {
        if (tid < TC) {
            const int g = chain0 + tid, J = p.J[0];
            R u = uniform_at(tid, p.ucur[g] + J);
            int k = (int)tfloor(u * (R)J);
            if (k > J - 1) k = J - 1;
            p.promo_j[g] = k + 1;          // index -J+k = the link after coarse step k+1
        }
        __syncthreads();
    }
    __device__ void snapshot_promoted(int j) {
        const int d = p.d;
        const LevelP<R>& lo = p.lv[0];
        for (int e = tid; e < d * TC; e += NT) {
            int k = e / TC, c = e - k * TC;
            if (p.promo_j[chain0 + c] == j) p.pm_theta[gi(k, c)] = lo.theta[gi(k, c)];
        }
        if (lo.need_F)
            for (int e = tid; e < lo.m * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                if (p.promo_j[chain0 + c] == j) p.pm_F[gi(k, c)] = lo.F[gi(k, c)];
            }
        if (tid < TC) {
            const int g = chain0 + tid;
            if (p.promo_j[g] == j) { p.pm_prior[g] = lo.prior[g]; p.pm_like[g] = lo.like[g]; p.pm_sid[g] = lo.sid[g]; }
        }
        __syncthreads();
    }

    // ---- history ------------------------------------------------------------------------------
    __device__ void record(int l) {
        const LevelP<R>& v = p.lv[l];
        const long long r = rec[l];
        if (r < v.hist_cap) {
            const int d = p.d;
            if (v.store & TDA_STORE_THETA)
                for (int e = tid; e < d * TC; e += NT) {
                    int k = e / TC, c = e - k * TC;
                    v.h_theta[((size_t)r * d + k) * p.Cs + chain0 + c] = v.theta[gi(k, c)];
                }
            if ((v.store & TDA_STORE_OUTPUT) && v.need_F)
                for (int e = tid; e < v.m * TC; e += NT) {
                    int k = e / TC, c = e - k * TC;
                    v.h_F[((size_t)r * v.m + k) * p.Cs + chain0 + c] = v.F[gi(k, c)];
                }
            if (tid < TC) {
                size_t o = (size_t)r * p.Cs + chain0 + tid;
                if (v.store & TDA_STORE_STATS) { v.h_prior[o] = v.prior[chain0 + tid]; v.h_like[o] = v.like[chain0 + tid]; }
                if (v.store & TDA_STORE_ACCEPT) v.h_acc[o] = (uint8_t)s_acc[tid];
            }
        }
        if (l == p.L - 1) {
            // running moments of the finest level; the three arrays never alias, which lets the
            // read-modify-writes of different elements overlap instead of serialising on each load
            const int d = p.d;
            const R* __restrict__ th = v.theta;
            R* __restrict__ s1 = p.sum1;
            R* __restrict__ s2 = p.sum2;
#pragma unroll 4
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                const size_t o = gi(k, c);
                R x = th[o];
                s1[o] += x;
                s2[o] += x * x;
            }
        }
        rec[l] = r + 1;
    }

    // append one entry to the level-0 `accepted` window (thread per chain, tid < TC)
    __device__ __forceinline__ void window_append(int c, int acc) {
        if (!p.adaptive) return;
        const int pos = (int)(wcount % p.period);
        size_t o = (size_t)pos * p.Cs + chain0 + c;
        int old = (wcount >= p.period) ? (int)p.win[o] : 0;
        p.win[o] = (uint8_t)acc;
        p.win_sum[chain0 + c] += acc - old;
    }

    // ---- proposals ---------------------------------------------------------------------------
    __device__ void propose_gaussian() {
        // RWMH / pCN / AM:  theta' = a*theta + b*(z @ T)
        const int d = p.d;
        fill_normals();
        if (tid < TC) {
            R s = p.scaling[chain0 + tid];
            if (p.prop_kind == TDA_PROP_PCN) { s_ca[tid] = tsqrt((R)1 - s * s); s_cb[tid] = s; }
            else { s_ca[tid] = (R)1; s_cb[tid] = s; }
        }
        __syncthreads();
        if (p.prop_kind == TDA_PROP_INDEP) {
            // q.rvs(): mu_q + z @ T_q, independent of the current state (proposal.py:117-119)
            EpiAffine<R> e; e.pt = pt; e.mu = p.ow_lambda;
            tile_gemm<R>(zt, (const R*)nullptr, p.T, d, d, p.ldD, bs, KB, e);
        } else if (p.prop_kind == TDA_PROP_OWPCN && p.adaptive) {
            // per-chain step size s: S = V diag(sqrt(1 - s lam)) V^T, N = V diag(sqrt(s lam)) V^T
            // (proposal.py:578-591)  ->  theta' = [ sqrt(s lam) (V^T xi) + sqrt(1 - s lam) (V^T theta) ] V^T
            EpiTile<R> e1; e1.t = pt; e1.accumulate = 0;
            tile_gemm<R>(zt, (const R*)nullptr, p.T, d, d, p.ldD, bs, KB, e1);        // V^T xi
            __syncthreads();
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                zt[e] = prop_src[gi(k, c)];
            }
            __syncthreads();
            EpiOw<R> e2; e2.t = pt; e2.s = s_cb; e2.lam = p.ow_lambda;
            tile_gemm<R>(zt, (const R*)nullptr, p.Sop, d, d, p.ldD, bs, KB, e2);      // combine with V^T theta
            __syncthreads();
            for (int e = tid; e < d * TC; e += NT) zt[e] = pt[e];
            __syncthreads();
            EpiTile<R> e3; e3.t = pt; e3.accumulate = 0;
            tile_gemm<R>(zt, (const R*)nullptr, p.Sop2, d, d, p.ldD, bs, KB, e3);     // back: @ V^T
        } else if (p.prop_kind == TDA_PROP_OWPCN) {
            // theta' = S theta + N xi  (proposal.py:593-598) = theta @ S^T + z @ (T N^T)
            EpiTile<R> e1; e1.t = pt; e1.accumulate = 0;
            tile_gemm<R>(zt, (const R*)nullptr, p.T, d, d, p.ldD, bs, KB, e1);
            __syncthreads();
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                zt[e] = prop_src[gi(k, c)];
            }
            __syncthreads();
            EpiTile<R> e2; e2.t = pt; e2.accumulate = 1;
            tile_gemm<R>(zt, (const R*)nullptr, p.Sop, d, d, p.ldD, bs, KB, e2);
        } else if (p.prop_kind == TDA_PROP_AM) {
            // per-chain factor: thread (c, half) computes half of the output columns
            const int c = tid & (TC - 1), half = tid / TC;
            const int j0 = half * ((d + 1) / 2), j1 = min(d, j0 + (d + 1) / 2);
            for (int j = j0; j < j1; j++) {
                R s = (R)0;
                for (int k = 0; k < d; k++) s = fma(zt[k * TC + c], p.am_T[gi(k * d + j, c)], s);
                pt[j * TC + c] = s_ca[c] * prop_src[gi(j, c)] + s_cb[c] * s;
            }
        } else {
            EpiCombine<R> e;
            e.theta = prop_src; e.Cs = p.Cs; e.chain0 = chain0; e.pt = pt; e.ca = s_ca; e.cb = s_cb;
            tile_gemm<R>(zt, (const R*)nullptr, p.T, d, d, p.ldD, bs, KB, e);
        }
        __syncthreads();
    }

    __device__ void propose_mala() {
        const int d = p.d;
        fill_normals();
        __syncthreads();
        for (int e = tid; e < d * TC; e += NT) {
            int k = e / TC, c = e - k * TC;
            R s = p.scaling[chain0 + c];
            pt[e] = p.lv[0].theta[gi(k, c)] + (R)0.5 * s * s * p.grad[gi(k, c)] + s * zt[e];
        }
        __syncthreads();
    }

    // gradient of the log-posterior at the tile in pt (F in lv[0].Fp) -> p.gradp   (MALA)
    __device__ void mala_gradient(R* out) {
        const LevelP<R>& v = p.lv[0];
        const int d = p.d;
        if (v.model_kind == TDA_MODEL_LINEAR) {
            // sensitivity tile (data - F)/var -> zt[j][c], then  sens @ G  ([m] x [m][d])
            for (int e = tid; e < v.m * TC; e += NT) {
                int j = e / TC, c = e - j * TC;
                R r = v.data[j] - v.Fp[gi(j, c)];
                zt[e] = (v.lik_kind == TDA_LIK_ISO) ? r / v.lik_var : r / v.var[j];
            }
            __syncthreads();
            EpiStore<R> e1; e1.out = out; e1.Cs = p.Cs; e1.chain0 = chain0;
            tile_gemm<R>(zt, (const R*)nullptr, v.A2, v.m, d, p.ldD, bs, KB, e1);
        } else {   // ROSENBROCK
            if (tid < TC) {
                const int c = tid;
                R x = pt[c], y = pt[TC + c];
                R sens = (v.data[0] - v.Fp[gi(0, c)]) / ((v.lik_kind == TDA_LIK_ISO) ? v.lik_var : v.var[0]);
                R dFdx = (R)-2 * (v.sc0 - x) - (R)4 * v.sc1 * x * (y - x * x);
                R dFdy = (R)2 * v.sc1 * (y - x * x);
                out[gi(0, c)] = sens * dFdx;
                out[gi(1, c)] = sens * dFdy;
            }
        }
        __syncthreads();
        EpiStoreNeg<R> e2; e2.out = out; e2.Cs = p.Cs; e2.chain0 = chain0; e2.accumulate = 1;
        tile_gemm<R>(pt, p.prior_mean, p.Pprec, d, d, p.ldD, bs, KB, e2);
        __syncthreads();
    }

    __device__ __forceinline__ const R* archive_row(long long r, int c, long long nslots) const {
        long long g, slot;
        if (p.prop_kind == TDA_PROP_DREAM) { g = r / nslots; slot = r - g * nslots; }
        else { g = chain0 + c; slot = r; }     // DREAMZ: own (local) archive column
        return p.archive + ((size_t)slot * p.Cg + g) * p.d;
    }

    __device__ void propose_dream() {
        // proposal.py:811-852; integer draws derive from uniforms as documented in DESIGN.md.
        // The ~2*delta + 2 + 2d uniforms and d normals of a proposal are consecutive stream entries:
        // sequential cursors reuse each Philox block for 4 draws.
        const int d = p.d;
        if (tid < TC) {
            const int c = tid, delta = p.dream_delta;
            const long long nslots = slots;        // rows per chain visible now (advances inside a persistent launch)
            const long long M = (p.prop_kind == TDA_PROP_DREAM) ? nslots * p.Cg : nslots;
            const bool seq_z = !p.z_round;
            ChainStreams<R> rs(p, chain0 + c);
            rs.seek_uniform(p.ucur[chain0 + c]);
            const long long z0 = t_base * d;
            if (seq_z) rs.seek_normal(z0);
            long long r1[MAX_DELTA], r2[MAX_DELTA];
            for (int i = 0; i < delta; i++) {
                R u1 = rs.uniform(), u2 = rs.uniform();
                long long a = (long long)tfloor(u1 * (R)M); if (a > M - 1) a = M - 1;
                long long b = (long long)tfloor(u2 * (R)(M - 1)); if (b > M - 2) b = M - 2;
                if (b >= a) b++;
                r1[i] = a; r2[i] = b;
            }
            R ucr = rs.uniform();
            int mCR = 0;
            {
                double cs = 0.0;
                for (int i = 0; i < p.dream_nCR; i++) {
                    cs += p.adaptive ? (double)p.dream_pCR[gi(i, c)] : 1.0 / p.dream_nCR;
                    if (cs <= (double)ucr) mCR++;
                }
                if (mCR > p.dream_nCR - 1) mCR = p.dream_nCR - 1;
            }
            if (p.adaptive) p.dream_mCR[chain0 + c] = mCR;
            const R CR = (R)(mCR + 1) / (R)p.dream_nCR;
            unsigned long long mask = 0ull;
            int card = 0;
            for (int k = 0; k < d; k++) if (rs.uniform() < CR) { mask |= 1ull << k; card++; }
            if (card == 0) {
                int k = (int)tfloor(rs.uniform() * (R)d); if (k > d - 1) k = d - 1;
                mask = 1ull << k; card = 1;
            }
            const R gam = p.scaling[chain0 + c] * (R)2.38 / tsqrt((R)(2 * delta * card));
            // jump direction sum(Z[r1]) - sum(Z[r2]) first, on its own: the two gathered archive rows of
            // a chain are scattered 4d-byte reads, and a loop with nothing but loads lets them overlap
            // (sums kept separately, zt = sum Z[r1], pt = sum Z[r2], like Z_r1 / Z_r2 in proposal.py:818-826)
            for (int k = 0; k < d; k++) { zt[k * TC + c] = (R)0; pt[k * TC + c] = (R)0; }
            for (int i = 0; i < delta; i++) {
                const R* __restrict__ ra = archive_row(r1[i], c, nslots);
                const R* __restrict__ rb = archive_row(r2[i], c, nslots);
#pragma unroll 8
                for (int k = 0; k < d; k++) { zt[k * TC + c] += __ldcg(ra + k); pt[k * TC + c] += __ldcg(rb + k); }   // L2: rows written by other CTAs / GPUs during this launch
            }
            for (int k = 0; k < d; k++) {
                R e = -p.dream_b + (p.dream_b + p.dream_b) * rs.uniform();
                R eps = p.dream_b_star * (seq_z ? rs.normal() : normal_at(c, z0 + k));
                R dz = zt[k * TC + c] - pt[k * TC + c];
                R th = p.lv[0].theta[gi(k, c)];
                pt[k * TC + c] = ((mask >> k) & 1ull) ? th + (((R)1 + e) * gam * dz + eps) : th;
            }
            p.ucur[chain0 + c] = (long long)rs.ui;
        }
        __syncthreads();
    }

    // log q(pt) of the IndependenceSampler's normal q, up to its constant -> s_cb (tid < TC)
    __device__ void indep_logq() {
        EpiSsq<R> e;
        tile_gemm<R>(pt, p.ow_lambda, p.Sop, p.d, p.d, p.ldD, bs, KB, e);
        R tot = reduce_cols(e.ps[0], e.ps[1]);
        if (tid < TC) s_cb[tid] = (R)-0.5 * tot;
        __syncthreads();
    }

    // ---- MultipleTry (ray.py:213-354) ---------------------------------------------------------------
    // log-density of the pCN kernel up to its constant between the tile in pt and the state src
    // -> s_cb (valid for tid < TC): q(src -> pt), the direction the reference uses (ray.py:292-296,
    // :337-342), or q(pt -> src) when `reverse` (the weights w(y, x) = pi(y) T(y, x) of Liu et al.'s
    // MTM(I), used by the detailed-balance mode).  Zero for symmetric kernels (MTM(II)).
    __device__ void mtm_q(const R* src, bool reverse) {
        const int d = p.d;
        if (p.prop_kind != TDA_PROP_PCN) {
            if (tid < TC) s_cb[tid] = (R)0;
            __syncthreads();
            return;
        }
        for (int e = tid; e < d * TC; e += NT) {
            int k = e / TC, c = e - k * TC;
            R s = p.scaling[chain0 + c];
            const R a = tsqrt((R)1 - s * s);
            zt[e] = reverse ? src[gi(k, c)] - a * pt[e] : pt[e] - a * src[gi(k, c)];
        }
        __syncthreads();
        EpiSsq<R> e;
        tile_gemm<R>(zt, (const R*)nullptr, p.LP, d, d, p.ldD, bs, KB, e);
        R tot = reduce_cols(e.ps[0], e.ps[1]);
        if (tid < TC) { R s = p.scaling[chain0 + tid]; s_cb[tid] = (R)-0.5 * tot / (s * s); }
        __syncthreads();
    }

    // scipy.special.logsumexp of w[0..n) (stride Cs) for chain c
    __device__ R mtm_logsumexp(const R* w, int n, int c) const {
        if (n == 0) return -INFINITY;
        R a_max = w[gi(0, c)];
        for (int i = 1; i < n; i++) a_max = fmax(a_max, w[gi(i, c)]);
        if (!isfinite(a_max)) a_max = (R)0;
        R s = (R)0;
        for (int i = 0; i < n; i++) s += texp(w[gi(i, c)] - a_max);
        return tlog(s) + a_max;
    }

    // make_proposal + get_acceptance of MultipleTry: leaves the chosen candidate in pt / s_prior /
    // s_like / Fp and the acceptance probability in s_ca
    __device__ void mtm_step() {
        const int d = p.d, K = p.mtm_k;
        const LevelP<R>& v = p.lv[0];
        mtm_zoff = 0;
        prop_src = v.theta;
        for (int i = 0; i < K; i++) {                  // ray.py:281-303: k candidates and their weights
            propose_gaussian();
            mtm_zoff += d;
            eval_level(0);
            mtm_q(v.theta, p.mtm_include_current != 0);
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                p.mt_theta[((size_t)i * d + k) * p.Cs + chain0 + c] = pt[e];
            }
            if (v.need_F)
                for (int e = tid; e < v.m * TC; e += NT) {
                    int k = e / TC, c = e - k * TC;
                    p.mt_F[((size_t)i * v.m + k) * p.Cs + chain0 + c] = v.Fp[gi(k, c)];
                }
            if (tid < TC) {
                const int c = tid;
                R w = (s_prior[c] + s_like[c]) + s_cb[c];
                if (tisnan(w)) w = -INFINITY;
                p.mt_prior[gi(i, c)] = s_prior[c];
                p.mt_like[gi(i, c)] = s_like[c];
                p.mt_w[gi(i, c)] = w;
            }
            __syncthreads();
        }
        if (tid < TC) {                                // ray.py:305-317: choose one
            const int c = tid, g = chain0 + c;
            bool all_inf = true;
            for (int i = 0; i < K; i++) if (!isinf(p.mt_w[gi(i, c)])) all_inf = false;
            long long uc = p.ucur[g];
            const R u = uniform_at(c, uc);
            p.ucur[g] = uc + 1;
            int idx;
            const R lse = mtm_logsumexp(p.mt_w, K, c);
            if (all_inf) {
                idx = (int)tfloor(u * (R)K);
                if (idx > K - 1) idx = K - 1;
            } else {
                R cs = (R)0;
                idx = 0;
                for (int i = 0; i < K; i++) { cs += texp(p.mt_w[gi(i, c)] - lse); if (cs <= u) idx++; }
                if (idx > K - 1) idx = K - 1;
            }
            const R post = p.mt_prior[gi(idx, c)] + p.mt_like[gi(idx, c)];
            const bool dead = all_inf || tisnan(post);          // ray.py:321: acceptance 0, no reference draws
            p.mt_sel[g] = idx;
            s_acc[c] = dead ? 1 : 0;
            p.mt_lse[g] = lse;
        }
        __syncthreads();
        for (int e = tid; e < d * TC; e += NT) {
            int k = e / TC, c = e - k * TC;
            p.mt_y[gi(k, c)] = p.mt_theta[((size_t)p.mt_sel[chain0 + c] * d + k) * p.Cs + chain0 + c];
        }
        __syncthreads();
        prop_src = p.mt_y;
        for (int i = 0; i < K - 1; i++) {              // ray.py:325-347: k-1 reference points
            propose_gaussian();
            mtm_zoff += d;
            eval_level(0);
            mtm_q(p.mt_y, p.mtm_include_current != 0);
            if (tid < TC) {
                const int c = tid;
                R w = (s_prior[c] + s_like[c]) + s_cb[c];
                if (tisnan(w)) w = -INFINITY;
                p.mt_w[gi(i, c)] = w;
            }
            __syncthreads();
        }
        prop_src = v.theta;
        int n_ref = K - 1;
        if (p.mtm_include_current) {
            // Liu et al. (2000): the current state x is the k-th reference point, weight pi(x) q(y, x)
            for (int e = tid; e < d * TC; e += NT) pt[e] = v.theta[gi(e / TC, e % TC)];
            __syncthreads();
            mtm_q(p.mt_y, true);          // pt = x, src = y: q(x -> y)
            if (tid < TC) {
                const int c = tid, g = chain0 + c;
                R w = (v.prior[g] + v.like[g]) + s_cb[c];
                if (tisnan(w)) w = -INFINITY;
                p.mt_w[gi(K - 1, c)] = w;
            }
            n_ref = K;
            __syncthreads();
        }
        // the chosen candidate becomes the proposal link (chain.py:105)
        for (int e = tid; e < d * TC; e += NT) pt[e] = p.mt_y[gi(e / TC, e % TC)];
        if (v.need_F)
            for (int e = tid; e < v.m * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                v.Fp[gi(k, c)] = p.mt_F[((size_t)p.mt_sel[chain0 + c] * v.m + k) * p.Cs + chain0 + c];
            }
        if (tid < TC) {
            const int c = tid, g = chain0 + c, idx = p.mt_sel[g];
            s_prior[c] = p.mt_prior[gi(idx, c)];
            s_like[c] = p.mt_like[gi(idx, c)];
            s_ca[c] = s_acc[c] ? (R)0 : texp(p.mt_lse[g] - mtm_logsumexp(p.mt_w, n_ref, c));   // ray.py:350-353
            // a chain whose candidates were all invalid draws no reference points (ray.py:321-323)
            p.zcur[g] += s_acc[c] ? (long long)K * d : (long long)(2 * K - 1) * d;
        }
        __syncthreads();
    }

    // ---- base level step -----------------------------------------------------------------------
    __device__ void base_step() {
        const int d = p.d;
        const LevelP<R>& v = p.lv[0];
        if (p.mtm_k) mtm_step();
        else {
            if (p.prop_kind == TDA_PROP_MALA) propose_mala();
            else if (is_dream(p.prop_kind)) propose_dream();
            else propose_gaussian();
            eval_level(0);
        }
        if (p.prop_kind == TDA_PROP_MALA) mala_gradient(p.gradp);
        if (p.prop_kind == TDA_PROP_INDEP) indep_logq();
        if (tid < TC) {
            const int c = tid, g = chain0 + c;
            R pr = s_prior[c], lk = s_like[c];
            R pr0 = v.prior[g], lk0 = v.like[g];
            R x;
            if (is_pcn_like(p.prop_kind)) x = lk - lk0;
            else x = (pr + lk) - (pr0 + lk0);
            if (p.prop_kind == TDA_PROP_INDEP) x = x + p.qcur[g] - s_cb[c];     // proposal.py:121-127
            if (p.prop_kind == TDA_PROP_MALA) {
                R s = p.scaling[g];
                R qxy = (R)0, qyx = (R)0;
                for (int k = 0; k < d; k++) {
                    R th = v.theta[gi(k, c)], tp = pt[k * TC + c];
                    R a = th - tp - (R)0.5 * s * s * p.gradp[gi(k, c)];
                    R b = tp - th - (R)0.5 * s * s * p.grad[gi(k, c)];
                    qxy = fma(a, a, qxy);
                    qyx = fma(b, b, qyx);
                }
                const R f = (R)-0.5 / (s * s);
                x = x + f * qxy - f * qyx;
            }
            R alpha = tisnan(pr + lk) ? (R)0 : texp(x);
            if (p.mtm_k) alpha = s_ca[c];
            long long uc = p.ucur[g];
            R u = uniform_at(c, uc);
            p.ucur[g] = uc + 1;
            int acc = (u < alpha) ? 1 : 0;
            s_acc[c] = acc;
            if (acc) {
                v.prior[g] = pr; v.like[g] = lk;
                if (p.prop_kind == TDA_PROP_INDEP) p.qcur[g] = s_cb[c];
                v.sid[g] = (int)(t_base + 1);
                v.n_acc[g] += 1;
                v.acc_sub[g] += 1;
            }
            window_append(c, acc);
        }
        __syncthreads();
        // DREAM(Z) crossover adaptation needs the squared jump of this step (proposal.py:799)
        const bool cr_adapt = is_dream(p.prop_kind) && p.adaptive && ((t_base + 1) % p.period) == 0;
        if (cr_adapt)
            for (int e = tid; e < d * TC; e += NT) {
                int c = e % TC, k = e / TC;
                R jd = s_acc[c] ? pt[e] - v.theta[gi(k, c)] : (R)0;
                zt[e] = jd * jd;
            }
        // accepted chains: proposal -> current state
        for (int e = tid; e < d * TC; e += NT) {
            int c = e % TC;
            if (s_acc[c]) { int k = e / TC; v.theta[gi(k, c)] = pt[e]; }
        }
        if (v.need_F)
            for (int e = tid; e < v.m * TC; e += NT) {
                int c = e % TC;
                if (s_acc[c]) { int k = e / TC; v.F[gi(k, c)] = v.Fp[gi(k, c)]; }
            }
        if (p.prop_kind == TDA_PROP_MALA)
            for (int e = tid; e < d * TC; e += NT) {
                int c = e % TC;
                if (s_acc[c]) { int k = e / TC; p.grad[gi(k, c)] = p.gradp[gi(k, c)]; }
            }
        __syncthreads();
        wcount += 1;
        t_base += 1;
        record(0);
        adapt();
        __syncthreads();
    }

    // proposal.adapt(): global scaling, AM moments / refactor, DREAM archive append
    __device__ void adapt() {
        const int d = p.d;
        const long long t = t_base;     // already incremented: proposal.t after this call
        if (p.adaptive && (t % p.period) == 0 && tid < TC) {
            const int g = chain0 + tid;
            const long long k = t / p.period - 1;
            R rate = (R)p.win_sum[g] / (R)p.period;
            R s = p.scaling[g];
            p.scaling[g] = texp(tlog(s) + tpow(p.gamma, (R)(-(double)k)) * (rate - p.alpha_star));
        }
        if (p.prop_kind == TDA_PROP_AM) {
            // RecursiveSampleMoments.update(theta_cur), recursor.t = t (starts at 1)
            const R tt = (R)t;
            const R f1 = (tt - (R)1) / tt, f2 = p.am_sd / tt;
            for (int e = tid; e < d * d * TC; e += NT) {
                int c = e % TC, ij = e / TC, i = ij / d, j = ij - i * d;
                R xi = p.lv[0].theta[gi(i, c)], xj = p.lv[0].theta[gi(j, c)];
                R mpi = p.am_mu[gi(i, c)], mpj = p.am_mu[gi(j, c)];
                R mni = ((R)1 / (tt + (R)1)) * (tt * mpi + xi), mnj = ((R)1 / (tt + (R)1)) * (tt * mpj + xj);
                size_t o = gi(ij, c);
                p.am_sigma[o] = f1 * p.am_sigma[o] + f2 * (tt * mpi * mpj - (tt + (R)1) * mni * mnj + xi * xj + ((i == j) ? p.am_eps : (R)0));
            }
            __syncthreads();
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                p.am_mu[gi(k, c)] = ((R)1 / (tt + (R)1)) * (tt * p.am_mu[gi(k, c)] + p.lv[0].theta[gi(k, c)]);
            }
            if (p.am_device_refactor && t >= p.am_t0 && (t % p.period) == 0) {
                __syncthreads();
                if (tid < TC) {   // T = chol(sigma)^T so that z @ T = L z
                    const int c = tid;
                    R* T = p.am_T;
                    for (int j = 0; j < d; j++) {
                        R djj = p.am_sigma[gi(j * d + j, c)];
                        for (int k = 0; k < j; k++) { R t2 = T[gi(k * d + j, c)]; djj -= t2 * t2; }
                        if (!(djj > (R)0)) { djj = (R)1e-30; *p.error_flag = 1; }     // the reference factors by SVD
                        djj = tsqrt(djj);
                        T[gi(j * d + j, c)] = djj;
                        for (int i = j + 1; i < d; i++) {
                            R s = p.am_sigma[gi(i * d + j, c)];
                            for (int k = 0; k < j; k++) s -= T[gi(k * d + i, c)] * T[gi(k * d + j, c)];
                            T[gi(j * d + i, c)] = s / djj;     // L[i][j] stored at T[j][i]
                            T[gi(i * d + j, c)] = (R)0;
                        }
                    }
                }
            }
        }
        if (is_dream(p.prop_kind)) {
            // archive append of the current state (proposal.py:794 / :1652)
            if (slots < p.dream_cap)
                for (int e = tid; e < d * TC; e += NT) {
                    int c = e / d, k = e - c * d;
                    if (chain0 + c < p.C) {
                        const size_t o = ((size_t)slots * p.Cg + p.arch_off + chain0 + c) * d + k;
                        const R x = p.lv[0].theta[gi(k, c)];
                        p.archive[o] = x;
                        // the other GPUs' replicas, straight over NVLink (ray.py:372-376: update_archive)
                        for (int r = 0; r < p.n_peers; r++)
                            if (r != p.my_rank) p.peer_archive[r][o] = x;
                    }
                }
            slots += 1;
            if (p.adaptive) {
                // local archive moments (the shared-archive variant keeps a local Z too, proposal.py:794)
                for (int e = tid; e < d * TC; e += NT) {
                    int k = e / TC, c = e - k * TC;
                    R x = p.lv[0].theta[gi(k, c)];
                    p.arch_s1[gi(k, c)] += x;
                    p.arch_s2[gi(k, c)] += x * x;
                }
                if ((t % p.period) == 0) {
                    __syncthreads();
                    if (tid < TC) {
                        const int c = tid, nCR = p.dream_nCR;
                        const R M = (R)(p.dream_M0 + t);          // rows of the local archive
                        R acc = (R)0;
                        for (int k = 0; k < d; k++) {
                            R mean = p.arch_s1[gi(k, c)] / M;
                            R var = p.arch_s2[gi(k, c)] / M - mean * mean;      // np.var(Z, axis=0)
                            acc += zt[k * TC + c] / var;
                        }
                        const int mCR = p.dream_mCR[chain0 + c];
                        p.dream_DeltaCR[gi(mCR, c)] += acc;
                        p.dream_LCR[gi(mCR, c)] += (R)1;
                        bool all_pos = true;
                        R tot = (R)0;
                        for (int i = 0; i < nCR; i++) {
                            if (!(p.dream_LCR[gi(i, c)] > (R)0)) all_pos = false;
                            else tot += p.dream_DeltaCR[gi(i, c)] / p.dream_LCR[gi(i, c)];
                        }
                        if (all_pos)
                            for (int i = 0; i < nCR; i++)
                                p.dream_pCR[gi(i, c)] = p.dream_DeltaCR[gi(i, c)] / p.dream_LCR[gi(i, c)] / tot;
                    }
                }
            }
        }
    }

    // ---- alignment of the lower levels after a step of level l --------------------------------
    __device__ void align(int l) {
        const int d = p.d;
        for (int j = l - 1; j >= 0; j--) {
            const LevelP<R>& lo = p.lv[j];
            if (tid < TC) {
                const int c = tid, g = chain0 + c;
                if (s_acc[c]) {
                    lo.sv_prior[l][g] = lo.prior[g];
                    lo.sv_like[l][g] = lo.like[g];
                } else {
                    lo.prior[g] = lo.sv_prior[l][g];
                    lo.like[g] = lo.sv_like[l][g];
                    lo.sid[g] = p.lv[l].sid[g];
                    for (int a = j + 1; a < l; a++) { lo.sv_prior[a][g] = lo.sv_prior[l][g]; lo.sv_like[a][g] = lo.sv_like[l][g]; }
                }
                lo.acc_sub[g] = 0;
                if (j == 0) window_append(c, s_acc[c]);
            }
            for (int e = tid; e < d * TC; e += NT) {
                int c = e % TC;
                if (!s_acc[c]) { int k = e / TC; lo.theta[gi(k, c)] = p.lv[l].theta[gi(k, c)]; }
            }
            if (lo.need_F)
                for (int e = tid; e < lo.m * TC; e += NT) {
                    int c = e % TC, k = e / TC;
                    if (s_acc[c]) lo.sv_F[l][gi(k, c)] = lo.F[gi(k, c)];
                    else {
                        R f = lo.sv_F[l][gi(k, c)];
                        lo.F[gi(k, c)] = f;
                        for (int a = j + 1; a < l; a++) lo.sv_F[a][gi(k, c)] = f;
                    }
                }
        }
        wcount += 1;
        __syncthreads();
    }

    // ---- step of level l >= 1 --------------------------------------------------------------------
    __device__ void upper_step(int l) {
        const int d = p.d;
        const LevelP<R>& v = p.lv[l];
        const LevelP<R>& lo = p.lv[l - 1];
        const bool rnd = p.randomize != 0;
        // the promoted link: the subchain's last one, or the randomly chosen one
        const R* b_theta = rnd ? p.pm_theta : lo.theta;
        const R* b_prior = rnd ? p.pm_prior : lo.prior;
        const R* b_like = rnd ? p.pm_like : lo.like;
        const R* b_F = rnd ? p.pm_F : lo.F;
        const int* b_sid = rnd ? p.pm_sid : lo.sid;
        for (int e = tid; e < d * TC; e += NT) {
            int k = e / TC, c = e - k * TC;
            pt[e] = b_theta[gi(k, c)];
        }
        __syncthreads();
        eval_level(l);
        if (p.aem == 2) sd_terms(l);
        if (tid < TC) {
            const int c = tid, g = chain0 + c;
            int acc = 0;
            if (lo.acc_sub[g] > 0) {
                R pr = s_prior[c], lk = s_like[c];
                const R post_new = pr + lk, post_cur = v.prior[g] + v.like[g];
                const R post_below = b_prior[g] + b_like[g];
                R x;
                if (p.aem == 2) {
                    // chain.py:446-473: the subchain start re-scored with the bias at the proposal
                    const int m = lo.m;
                    R q = (R)0;
                    for (int i = 0; i < m; i++) {
                        R t = (R)0;
                        for (int j = 0; j <= i; j++)
                            t = fma(lo.lik_prec[gi(i * m + j, c)],
                                    lo.sv_F[l][gi(j, c)] + (v.Fp[gi(j, c)] - b_F[gi(j, c)]) - lo.data[j], t);
                        q = fma(t, t, q);
                    }
                    const R post_biased = lo.sv_prior[l][g] + (R)-0.5 * q;
                    const R qxy = s_ca[c], qyx = s_cb[c];
                    // Python's min(a, b): a unless b < a (so a NaN in `a` survives, like the reference)
                    const R n1 = post_new + qyx, n2 = post_biased + qxy;
                    const R d1 = post_cur + qxy, d2 = post_below + qyx;
                    x = ((n2 < n1) ? n2 : n1) - ((d2 < d1) ? d2 : d1);
                } else {
                    x = post_new - post_cur + (lo.sv_prior[l][g] + lo.sv_like[l][g]) - post_below;
                }
                R alpha = texp(x);
                long long uc = p.ucur[g];
                if (rnd) uc += 1;          // the index draw (read ahead by draw_promoted)
                R u = uniform_at(c, uc);
                p.ucur[g] = uc + 1;
                acc = (u < alpha) ? 1 : 0;
                if (acc) {
                    v.prior[g] = pr; v.like[g] = lk;
                    v.sid[g] = b_sid[g];
                    v.n_acc[g] += 1;
                    v.acc_sub[g] += 1;
                    if (rnd) { lo.prior[g] = b_prior[g]; lo.like[g] = b_like[g]; lo.sid[g] = b_sid[g]; }
                }
            }
            s_acc[c] = acc;
        }
        __syncthreads();
        for (int e = tid; e < d * TC; e += NT) {
            int c = e % TC;
            if (s_acc[c]) {
                int k = e / TC;
                v.theta[gi(k, c)] = pt[e];
                if (rnd) lo.theta[gi(k, c)] = pt[e];      // chain.py:383: the promoted link re-enters the coarse chain
            }
        }
        if (v.need_F)
            for (int e = tid; e < v.m * TC; e += NT) {
                int c = e % TC;
                if (s_acc[c]) { int k = e / TC; v.F[gi(k, c)] = v.Fp[gi(k, c)]; }
            }
        if (rnd && lo.need_F)
            for (int e = tid; e < lo.m * TC; e += NT) {
                int c = e % TC;
                if (s_acc[c]) { int k = e / TC; lo.F[gi(k, c)] = b_F[gi(k, c)]; }
            }
        __syncthreads();
        record(l);
        align(l);
        if (p.aem) {
            lvl_steps[l] += 1;
            if (p.aem == 2) aem_update_sd(l, lvl_steps[l]);
            else aem_update(l, lvl_steps[l]);
            push_bias(l);
        }
    }

    // ---- initial links, AEM set-up --------------------------------------------------------------
    __device__ void init() {
        const int d = p.d, L = p.L;
        // a fresh proposal object per sample() call (sampler.py:176 deep-copies it): step size,
        // adaptation window and stream cursors start over
        if (tid < TC) {
            const int g = chain0 + tid;
            p.scaling[g] = p.scaling0;
            p.ucur[g] = 0;
            if (p.mtm_k) p.zcur[g] = 0;
            if (p.adaptive) p.win_sum[g] = 0;
        }
        for (int l = 0; l < L; l++) {
            const LevelP<R>& v = p.lv[l];
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                pt[e] = v.theta[gi(k, c)];
            }
            if (v.lik_kind == TDA_LIK_ADAPTIVE)    // per-chain factor starts as inv(chol(cov)) (uploaded)
                for (int e = tid; e < v.m * v.m * TC; e += NT) {
                    int c = e % TC, ij = e / TC;
                    v.lik_prec[gi(ij, c)] = v.prec[ij];
                }
            __syncthreads();
            eval_level(l);
            if (tid < TC) {
                const int g = chain0 + tid;
                v.prior[g] = s_prior[tid]; v.like[g] = s_like[tid]; v.sid[g] = 0;
                v.n_acc[g] = 0; v.acc_sub[g] = 0;
                for (int a = l + 1; a < L; a++) { v.sv_prior[a][g] = s_prior[tid]; v.sv_like[a][g] = s_like[tid]; }
                s_acc[tid] = 1;
            }
            if (v.need_F)
                for (int e = tid; e < v.m * TC; e += NT) {
                    int c = e % TC, k = e / TC;
                    R f = v.Fp[gi(k, c)];
                    v.F[gi(k, c)] = f;
                    for (int a = l + 1; a < L; a++) v.sv_F[a][gi(k, c)] = f;
                }
            if (l == 0 && p.prop_kind == TDA_PROP_MALA) { __syncthreads(); mala_gradient(p.grad); }
            if (l == 0 && p.prop_kind == TDA_PROP_INDEP) {
                __syncthreads();
                indep_logq();
                if (tid < TC) p.qcur[chain0 + tid] = s_cb[tid];
            }
            __syncthreads();
        }
        if (p.prop_kind == TDA_PROP_AM)
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                p.am_mu[gi(k, c)] = p.lv[0].theta[gi(k, c)];
            }
        if (is_dream(p.prop_kind) && p.adaptive) {
            if (tid < TC)
                for (int i = 0; i < p.dream_nCR; i++) {
                    p.dream_pCR[gi(i, tid)] = (R)(1.0 / p.dream_nCR);
                    p.dream_DeltaCR[gi(i, tid)] = (R)1;
                    p.dream_LCR[gi(i, tid)] = (R)0;
                }
            for (int e = tid; e < d * TC; e += NT) {
                int k = e / TC, c = e - k * TC;
                R a1 = (R)0, a2 = (R)0;
                if (chain0 + c < p.C)
                    for (int sl = 0; sl < p.dream_M0; sl++) {
                        R x = p.archive[((size_t)sl * p.Cg + p.arch_off + chain0 + c) * d + k];
                        a1 += x; a2 += x * x;
                    }
                p.arch_s1[gi(k, c)] = a1;
                p.arch_s2[gi(k, c)] = a2;
            }
        }
        if (p.aem && L > 1) {
            for (int l = 1; l < L; l++) {
                const LevelP<R>& v = p.lv[l];
                for (int e = tid; e < v.m * TC; e += NT) {
                    int c = e % TC, k = e / TC;
                    R df = v.F[gi(k, c)] - p.lv[l - 1].F[gi(k, c)];
                    v.model_diff[gi(k, c)] = df;
                    v.bias_mu[gi(k, c)] = df;
                }
                for (int e = tid; e < v.m * v.m * TC; e += NT) {
                    int c = e % TC, ij = e / TC;
                    v.bias_sigma[gi(ij, c)] = (R)0;
                }
            }
            __syncthreads();
            for (int l = L - 1; l >= 1; l--) push_bias(l);
        }
        __syncthreads();
        record(L - 1);
    }

    // End of a lock-step step of the shared-archive ensemble: every chain of every GPU has appended its row
    // (proposal.py:1652) before any chain reads the archive again (proposal.py:1655).  Local CTAs meet at a
    // counter in global memory (all co-resident: one tile per CTA, grid <= resident CTAs); across GPUs rank r
    // announces "my rows of step t are in your replica" by a system-scope release store into slot r of every
    // peer's flag array, and waits for the same from everybody.
    __device__ void step_barrier(unsigned step) {
        __syncthreads();
        if (tid == 0) {
            if (p.n_peers > 1) __threadfence_system();
            else __threadfence();
            atomicAdd(p.grid_bar, 1u);
            const unsigned target = (step + 1u) * gridDim.x;
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.grid_bar) : "memory");
                if (v < target) __nanosleep(40);
            } while (v < target);
            if (p.n_peers > 1) {
                const unsigned flag = p.flag_base + step + 1u;
                if (blockIdx.x == 0) {
                    __threadfence_system();          // one fence, then the flag stores back to back (not a release per peer)
                    for (int r = 0; r < p.n_peers; r++)
                        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.peer_flags[r] + p.my_rank), "r"(flag) : "memory");
                }
                for (int r = 0; r < p.n_peers; r++) {
                    const unsigned* f = p.peer_flags[p.my_rank] + r;
                    do {
                        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                        if ((int)(v - flag) < 0) __nanosleep(100);
                    } while ((int)(v - flag) < 0);
                }
            }
        }
        __syncthreads();
    }

    __device__ void run() {
        const int L = p.L;
        int cnt[MAXL];
        for (int l = 0; l < MAXL; l++) cnt[l] = 0;
        long long it = 0;
        while (it < p.iterations) {
            if (p.randomize && cnt[0] == 0) draw_promoted();
            base_step();
            if (p.grid_sync) step_barrier((unsigned)it);
            if (L == 1) { it++; continue; }
            cnt[0]++;
            if (p.randomize) snapshot_promoted(cnt[0]);
            int l = 0;
            while (l < L - 1 && cnt[l] == p.J[l]) {
                cnt[l] = 0;
                upper_step(l + 1);
                l++;
                if (l < L - 1) cnt[l]++;
                else it++;
            }
        }
    }
};

template <typename R>
__global__ void __launch_bounds__(NT, (sizeof(R) == 4 ? 2 : 1))
chain_kernel(const __grid_constant__ Params<R> p, int kt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        Tile<R> T(p, smem_raw, kt);
        T.chain0 = tile * TC;
        T.t_base = p.t_base;
        T.wcount = p.wcount;
        T.slots = p.dream_slots;
        T.mtm_zoff = 0;
        T.prop_src = p.lv[0].theta;
        for (int l = 0; l < MAXL; l++) { T.rec[l] = p.rec[l]; T.lvl_steps[l] = p.lvl_steps[l]; }
        if (p.mode == MODE_INIT) T.init();
        else T.run();
        __syncthreads();
    }
}

// exports the Philox streams (what TDA_RNG_PHILOX consumes) for the CPU oracle
template <typename R>
__global__ void fill_streams_kernel(unsigned long long seed, long long chain_offset, int C,
                                    double* z, long long nz, double* u, long long nu, int z_round) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total_z = (long long)C * nz, total_u = (long long)C * nu;
    if (i < total_z) {
        long long c = i / nz, k = i - c * nz;
        z[i] = (double)philox_normal<R>(seed, chain_offset + c, k, z_round);
    }
    if (i < total_u) {
        long long c = i / nu, k = i - c * nu;
        u[i] = (double)philox_uniform<R>(seed, chain_offset + c, k);
    }
}

}  // namespace tda
