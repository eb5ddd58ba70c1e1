// MH / Delayed Acceptance / MLDA with the state-independent adaptive error model, one WARP per chain (BASELINE cfg4
// class and the small problems of the reference's notebooks): up to 4 levels, the 1-D Poisson model or a linear
// operator per level, random-walk or pCN base proposal (fixed or adaptively scaled step), Gaussian likelihoods
// (isotropic / diagonal on any level, AdaptiveGaussianLogLike below the finest level), at most 31 outputs per level
// and 32 parameters.
// Reference semantics: chain.py:680-769 + proposal.py:1502-1613 (the recursive MLDA proposal), chain.py:485-499 and
// proposal.py:1442-1467 (state-independent error model), utils.py:113-124 (RecursiveSampleMoments),
// distributions.py:332-449 (AdaptiveGaussianLogLike.set_bias) -- step for step what Tile::base_step /
// Tile::upper_step / Tile::align / Tile::aem_update / Tile::push_bias of the lock-step kernel (tda_kernels.cuh) do,
// which the golden fixtures pin to the reference; same chain-state buffers, same stream positions.
//
// Why a second kernel: the lock-step kernel deals one chain to one thread for everything that is not a tile
// contraction -- the tridiagonal sweeps, and above all the per-chain m x m work of the error model (assembly of
// cov + bias covariance, Cholesky, triangular inverse, |Li r|^2 per likelihood, the moment recursion), all of it
// streamed through global memory with one or two loads in flight per thread: 16 % issue utilisation, 570 GB/s,
// 76 ms per fine iteration of cfg4 at 32768 chains.  Here lane i of a warp owns row i of every per-chain matrix:
//   * the Cholesky factorisation keeps the row in registers and reads the pivot rows as 16-byte shared-memory
//     broadcasts (no block-wide barrier, no global traffic); the triangular inverse has lane j solve column j;
//   * the inverse factor of level 0 (250 of the 281 likelihoods of a cfg4 iteration) lives in shared memory, the
//     factors of the higher levels and the bias covariances in a warp-major global image ([chain][column][lane]:
//     every access one 128-byte line, 31 independent lines in flight per lane);
//   * the forward model needs no tridiagonal solve at all: -(k u')' = 1 with u(0) = u(1) = 0 in one dimension has
//     the flux q_c = q_0 - c h^2 in cell c, so u at node j is sum_{c<j} q_c / k_c with q_0 fixed by u(1) = 0 --
//     two warp scans over the per-lane segment sums of 1/k_c and c/k_c (lane s owns the cells between sensors
//     s-1 and s).  Algebraically the solution of the same tridiagonal system the reference model solves with the
//     Thomas algorithm (models.py Poisson1D); agreement to rounding is what the parity tests check.
#include <cstdlib>
#include <string>

#include "tda_mlda_warp.h"

namespace tda {

namespace {

constexpr int MW_LIS = 33;     // row stride of an inverse-factor image: element (i, j) at j * 33 + i
constexpr int MW_M = 31;       // matrices are padded to 31 x 31 (identity outside m x m)

enum { V_TH = 0, V_F, V_BIAS, V_MD, V_BMU, V_SVF, NVEC = V_SVF + MAXL - 1 };     // V_SVF + a - 1: saved output of level a >= 1
enum { S_PRIOR = 0, S_LIKE, S_SVP, S_SVL = S_SVP + MAXL, NSC = S_SVL + MAXL };
enum { I_SID = 0, I_ACCSUB, I_NACC, I_REC, NIS };     // I_REC: records written / steps taken by the level in this launch

template <typename R>
__host__ __device__ constexpr int mw_warp_elems() {
    // vec | scalars | prop, xb, Fp, (spare) | li0 | acol   (all in units of R; ints live in the scalar block's tail)
    return MAXL * NVEC * 32 + 64 + 4 * 32 + MW_MATW + MW_MATW;
}

template <typename R> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&o)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <> struct Vec4<double> {
    static __device__ __forceinline__ void ld(const double* p, double (&o)[4]) {
        const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
};

__device__ __forceinline__ float mw_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double mw_rsqrt(double x) { return rsqrt(x); }

template <typename R>
__device__ __forceinline__ R mw_sum(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename R>
__device__ __forceinline__ R mw_scan(R v, int lane) {      // inclusive prefix sum over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const R t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

template <typename R>
struct Mw {
    const Params<R>& p;
    const int lane, g;
    const long long gchain;
    const size_t Cs;
    R *vec, *sc, *prop, *xb, *Fp, *dinv, *li0, *acol;
    int* isc;
    const R *Ts, *LPs;      // CTA-shared images [j][32] of the proposal factor and the prior's whitening matrix
    const MwParams* mpp;    // Phi images [k][t][32] per level: shared memory for the coarse levels, global otherwise
    const R* cta;
    bool diagT, diagLP;
    R dT, dLP;
    R* sigw;      // this chain's bias covariances, [level][k * 32 + i]
    R* liw;       // this chain's inverse factors,  [level][j * 33 + i]
    long long t_base, ucur, wcount;
    R scal;                 // proposal step of this chain (adapted when the proposal is adaptive, proposal.py:228-245)
    int win_sum, win_old;   // sum over the accept window; ring entry that the latest append replaced (not subtracted yet)

    __device__ Mw(const Params<R>& p_, int lane_, int g_) : p(p_), lane(lane_), g(g_), gchain(p_.chain_offset + g_), Cs((size_t)p_.Cs) {}

    __device__ __forceinline__ R& V(int l, int item) { return vec[(l * NVEC + item) * 32 + lane]; }
    __device__ __forceinline__ R* Vp(int l, int item) { return vec + (l * NVEC + item) * 32; }
    __device__ __forceinline__ R S(int l, int item) const { return sc[l * NSC + item]; }
    __device__ __forceinline__ void setS(int l, int item, R v) { if (lane == 0) sc[l * NSC + item] = v; }
    __device__ __forceinline__ int I(int l, int item) const { return isc[l * NIS + item]; }
    __device__ __forceinline__ void setI(int l, int item, int v) { if (lane == 0) isc[l * NIS + item] = v; }
    __device__ __forceinline__ R* li_of(int l) { return l == 0 ? li0 : liw + (size_t)l * MW_MATW; }

    __device__ __forceinline__ R uniform_at(long long idx) const {
        if (p.rng_mode == TDA_RNG_INJECTED) return (idx < p.ulen) ? p.us[(size_t)g * p.ulen + idx] : (R)0.5;
        return philox_uniform<R>(p.seed, gchain, idx);
    }
    __device__ __forceinline__ R normal_at(long long idx) const {
        if (p.rng_mode == TDA_RNG_INJECTED) return (idx < p.zlen) ? p.zs[(size_t)g * p.zlen + idx] : (R)0;
        return philox_normal<R>(p.seed, gchain, idx, 0);
    }

    // y[lane] = sum_j x[j] M[j][lane] (x in shared memory, zero beyond d; M: CTA-shared image [j][32], zero padded).
    // A diagonal M (identity-covariance priors, GaussianRandomWalk(C = c I)) is an elementwise product.
    __device__ __forceinline__ R matvec_d(const R* Ms, bool diag, R mdiag, const R* x) const {
        if (diag) return x[lane] * mdiag;
        const int d = p.d;
        R y0 = (R)0, y1 = (R)0;
        for (int j0 = 0; j0 < d; j0 += 4) {
            R xv[4];
            Vec4<R>::ld(x + j0, xv);
            const R* row = Ms + j0 * 32 + lane;
            y0 = fma(xv[0], row[0], y0);
            y1 = fma(xv[1], row[32], y1);
            y0 = fma(xv[2], row[64], y0);
            y1 = fma(xv[3], row[96], y1);
        }
        return y0 + y1;
    }

    // segment sums of 1/k_c and (c - c0)/k_c over the STRIDE cells of this lane; Phi image [k][t][32] (shared or global)
    template <int STRIDE>
    __device__ __forceinline__ void flux_sums(const R* __restrict__ Phi, int d, R c0, R& s0, R& s1) const {
        R acc[STRIDE];
#pragma unroll
        for (int t = 0; t < STRIDE; t++) acc[t] = (R)0;
        const R* col = Phi + lane;
        for (int k0 = 0; k0 < d; k0 += 4) {
            R th[4];
            Vec4<R>::ld(prop + k0, th);
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int t = 0; t < STRIDE; t++) acc[t] = fma(th[u], col[((k0 + u) * STRIDE + t) * 32], acc[t]);
        }
        s0 = (R)0; s1 = (R)0;
#pragma unroll
        for (int t = 0; t < STRIDE; t++) {
            const R r = texp(-acc[t]);              // 1 / k_c
            s0 += r;
            s1 = fma((R)(lane * STRIDE + t) - c0, r, s1);
        }
    }

    // |Li r|^2 with r = F + bias - data in xb (zero beyond m): lane i forms row i of Li r
    __device__ __forceinline__ R quad_form(const R* li) const {
        R t0 = (R)0, t1 = (R)0;
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 4) {
            R rv[4];
            Vec4<R>::ld(xb + j0, rv);
            t0 = fma(li[(j0 + 0) * MW_LIS + lane], rv[0], t0);
            t1 = fma(li[(j0 + 1) * MW_LIS + lane], rv[1], t1);
            t0 = fma(li[(j0 + 2) * MW_LIS + lane], rv[2], t0);
            if (j0 + 3 < MW_M) t1 = fma(li[(j0 + 3) * MW_LIS + lane], rv[3], t1);
        }
        const R t = (lane < MW_M) ? t0 + t1 : (R)0;
        return mw_sum<R>(t * t);
    }

    // One entry joins the level-0 `accepted` window (ring [period][Cs] in global memory, shared with the lock-step
    // kernel).  The entry it replaces is loaded here and subtracted at the NEXT append (or before the window is
    // read): the load's latency stays off the step's dependency chain.
    __device__ __forceinline__ void window_append(int acc) {
        if (!p.adaptive) return;
        win_sum += acc - win_old;
        const int pos = (int)(wcount % p.period);
        uint8_t* w = p.win + (size_t)pos * Cs + g;
        win_old = (wcount >= p.period) ? (int)__ldcg(w) : 0;
        __syncwarp();
        if (lane == 0) *w = (uint8_t)acc;
    }
    // proposal.adapt() after a coarsest-level step (proposal.py:228-245)
    __device__ __forceinline__ void adapt() {
        if (!p.adaptive || (t_base % p.period) != 0) return;
        win_sum -= win_old;
        win_old = 0;
        const long long k = t_base / p.period - 1;
        const R rate = (R)win_sum / (R)p.period;
        scal = texp(tlog(scal) + tpow(p.gamma, (R)(-(double)k)) * (rate - p.alpha_star));
    }

    // Link of the parameters in prop[] on level l (posterior.py:78-110): log-prior, forward model -> Fp, log-likelihood
    __device__ __forceinline__ void eval(int l, R& pr, R& lk) {
        const LevelP<R>& v = p.lv[l];
        const int d = p.d, m = v.m;
        __syncwarp();
        xb[lane] = (lane < d) ? prop[lane] - __ldg(p.prior_mean + lane) : (R)0;
        __syncwarp();
        {
            const R y = matvec_d(LPs, diagLP, dLP, xb);
            pr = (R)-0.5 * (p.prior_logconst + mw_sum<R>(y * y));
        }
        // forward model: a linear operator (F = theta @ A + b, lane n forms output n), or u at the sensors of the
        // Poisson problem from the cell fluxes (see the header)
        R F;
        if (v.model_kind == TDA_MODEL_LINEAR) {
            R a0 = (R)0, a1 = (R)0;
            if (lane < m) {
                const R* __restrict__ col = v.A + lane;
                for (int k0 = 0; k0 < d; k0 += 4) {
                    R th[4];
                    Vec4<R>::ld(prop + k0, th);
                    a0 = fma(th[0], __ldg(col + (size_t)k0 * v.ldA), a0);
                    if (k0 + 1 < d) a1 = fma(th[1], __ldg(col + (size_t)(k0 + 1) * v.ldA), a1);
                    if (k0 + 2 < d) a0 = fma(th[2], __ldg(col + (size_t)(k0 + 2) * v.ldA), a0);
                    if (k0 + 3 < d) a1 = fma(th[3], __ldg(col + (size_t)(k0 + 3) * v.ldA), a1);
                }
            }
            F = (lane < m) ? (a0 + a1) + __ldg(v.b + lane) : (R)0;
        } else {
            const int n = v.n_grid, stride = v.stride;
            const R c0 = (R)(n / 2);
            R s0 = (R)0, s1 = (R)0;
            const R* Phi = mpp->phi_smem[l] >= 0 ? cta + mpp->phi_smem[l] : reinterpret_cast<const R*>(mpp->phiw) + mpp->phi_off[l];
            if (stride == 2) flux_sums<2>(Phi, d, c0, s0, s1);
            else if (stride == 4) flux_sums<4>(Phi, d, c0, s0, s1);
            else if (stride == 8) flux_sums<8>(Phi, d, c0, s0, s1);
            else if (stride == 16) flux_sums<16>(Phi, d, c0, s0, s1);
            else if (stride == 1) flux_sums<1>(Phi, d, c0, s0, s1);
            else {
                for (int t = 0; t < stride; t++) {
                    R a0 = (R)0;
                    for (int k = 0; k < d; k++) a0 = fma(prop[k], Phi[(k * stride + t) * 32 + lane], a0);
                    const R r = texp(-a0);
                    s0 += r;
                    s1 = fma((R)(lane * stride + t) - c0, r, s1);
                }
            }
            if (lane > m) { s0 = (R)0; s1 = (R)0; }
            const R P0 = mw_scan<R>(s0, lane), P1 = mw_scan<R>(s1, lane);
            const R S0 = __shfl_sync(0xffffffffu, P0, m), S1 = __shfl_sync(0xffffffffu, P1, m);
            const R h2 = (R)1 / ((R)n * (R)n);
            F = (lane < m) ? h2 * ((S1 / S0) * P0 - P1) : (R)0;
        }
        __syncwarp();
        Fp[lane] = F;
        if (v.lik_kind == TDA_LIK_ADAPTIVE) {
            xb[lane] = (lane < m) ? F + V(l, V_BIAS) - __ldg(v.data + lane) : (R)0;
            __syncwarp();
            lk = (R)-0.5 * quad_form(li_of(l));
        } else {
            R q = (R)0;
            if (lane < m) {
                const R r = F - __ldg(v.data + lane);
                q = (v.lik_kind == TDA_LIK_ISO) ? r * r : r * r / __ldg(v.var + lane);
            }
            q = mw_sum<R>(q);
            lk = (v.lik_kind == TDA_LIK_ISO) ? (R)-0.5 * q / v.lik_var : (R)-0.5 * q;
        }
    }

    __device__ __forceinline__ void record(int l, int acc) {
        const LevelP<R>& v = p.lv[l];
        const long long r = p.rec[l] + I(l, I_REC);
        if (r < v.hist_cap) {
            if ((v.store & TDA_STORE_THETA) && lane < p.d) v.h_theta[((size_t)r * p.d + lane) * Cs + g] = V(l, V_TH);
            if ((v.store & TDA_STORE_OUTPUT) && v.need_F && lane < v.m) v.h_F[((size_t)r * v.m + lane) * Cs + g] = V(l, V_F);
            if (lane == 0) {
                if (v.store & TDA_STORE_STATS) { v.h_prior[(size_t)r * Cs + g] = S(l, S_PRIOR); v.h_like[(size_t)r * Cs + g] = S(l, S_LIKE); }
                if (v.store & TDA_STORE_ACCEPT) v.h_acc[(size_t)r * Cs + g] = (uint8_t)acc;
            }
        }
        __syncwarp();
        if (lane == 0) isc[l * NIS + I_REC] += 1;
        __syncwarp();
    }

    // chain.py:101-125 on the coarsest level
    __device__ __forceinline__ void base_proposal() {
        const int d = p.d;
        __syncwarp();
        xb[lane] = (lane < d) ? normal_at(t_base * d + lane) : (R)0;
        __syncwarp();
        {
            const R xi = matvec_d(Ts, diagT, dT, xb);
            const R s = scal;
            const R ca = (p.prop_kind == TDA_PROP_PCN) ? tsqrt((R)1 - s * s) : (R)1;
            __syncwarp();
            prop[lane] = (lane < d) ? ca * V(0, V_TH) + s * xi : (R)0;
        }
    }
    __device__ __forceinline__ void base_accept(R pr, R lk, R& s1, R& s2) {
        const R pr0 = S(0, S_PRIOR), lk0 = S(0, S_LIKE);
        const R x = (p.prop_kind == TDA_PROP_PCN) ? lk - lk0 : (pr + lk) - (pr0 + lk0);
        const R alpha = tisnan(pr + lk) ? (R)0 : texp(x);
        const R u = uniform_at(ucur);
        ucur += 1;
        const int acc = (u < alpha) ? 1 : 0;
        __syncwarp();
        if (acc) {
            V(0, V_TH) = prop[lane];
            V(0, V_F) = Fp[lane];
            if (lane == 0) {
                sc[S_PRIOR] = pr; sc[S_LIKE] = lk;
                isc[I_SID] = (int)(t_base + 1);
                isc[I_NACC] += 1;
                isc[I_ACCSUB] += 1;
            }
        }
        __syncwarp();
        window_append(acc);
        wcount += 1;
        t_base += 1;
        record(0, acc);
        if (p.L == 1) { const R th = V(0, V_TH); s1 += th; s2 += th * th; }
        adapt();
    }

    // after a step of level l: the lower levels follow it (proposal.py:1583-1613)
    __device__ __forceinline__ void align(int l, int acc) {
        for (int j = l - 1; j >= 0; j--) {
            if (acc) {
                setS(j, S_SVP + l, S(j, S_PRIOR));
                setS(j, S_SVL + l, S(j, S_LIKE));
                V(j, V_SVF + l - 1) = V(j, V_F);
            } else {
                const R sp = S(j, S_SVP + l), sl = S(j, S_SVL + l);
                const int sid = I(l, I_SID);
                __syncwarp();
                setS(j, S_PRIOR, sp); setS(j, S_LIKE, sl);
                setI(j, I_SID, sid);
                const R f = V(j, V_SVF + l - 1);
                V(j, V_F) = f;
                V(j, V_TH) = V(l, V_TH);
                for (int a = j + 1; a < l; a++) { setS(j, S_SVP + a, sp); setS(j, S_SVL + a, sl); V(j, V_SVF + a - 1) = f; }
            }
            setI(j, I_ACCSUB, 0);
            if (j == 0) window_append(acc);
        }
        wcount += 1;
        __syncwarp();
    }

    // bias.update(model_diff)  (utils.py:113-124), t = steps of level l so far
    __device__ __forceinline__ void aem_update(int l, long long tcount, int acc) {
        const int m = p.lv[l].m;
        if (acc) V(l, V_MD) = (lane < m) ? V(l, V_F) - V(l - 1, V_F) : (R)0;
        __syncwarp();
        const R t = (R)tcount;
        const R f1 = (t - (R)1) / t, f2 = (R)1 / t, g1 = (R)1 / (t + (R)1);
        const R xi = V(l, V_MD), mpi = V(l, V_BMU);
        const R mni = g1 * (t * mpi + xi);
        R* sg = sigw + (size_t)l * MW_MATW + lane;
        const R* md = Vp(l, V_MD);
        const R* bm = Vp(l, V_BMU);
#pragma unroll 2
        for (int k0 = 0; k0 < 32; k0 += 4) {
            R o[4], xv[4], mv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) o[u] = (k0 + u < MW_M) ? sg[(k0 + u) * 32] : (R)0;
            Vec4<R>::ld(md + k0, xv);
            Vec4<R>::ld(bm + k0, mv);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const R xj = xv[u], mpj = mv[u];
                const R mnj = g1 * (t * mpj + xj);
                if (k0 + u < MW_M) sg[(k0 + u) * 32] = f1 * o[u] + f2 * (t * mpi * mpj - (t + (R)1) * mni * mnj + xi * xj);
            }
        }
        __syncwarp();
        V(l, V_BMU) = g1 * (t * mpi + xi);
        __syncwarp();
    }

    // level l hands its bias moments to level l-1's likelihood (distributions.py:385-402) and level l-1's
    // current link is scored again (posterior.py:112-134)
    __device__ __forceinline__ void push_bias(int l) {
        const int L = p.L;
        const LevelP<R>& lo = p.lv[l - 1];
        const int m = lo.m;
        const int kend = (l == L - 1) ? l : L - 1;
        {
            R mu = (R)0;
            for (int k = l; k <= kend; k++) mu += V(k, V_BMU);
            V(l - 1, V_BIAS) = mu;
        }
        // W = cov + sum of the bias covariances, column-major in shared memory: element (i, k) at k * 32 + i, lane i
        // owns row i (conflict-free); padded with the identity outside m x m
        int big = 0;
        {
            const bool row = lane < m;
            const R* sg0 = sigw + (size_t)l * MW_MATW + lane;
            const int nlev = kend - l + 1;
#pragma unroll 8
            for (int k = 0; k < MW_M; k++) {
                R sg = sg0[k * 32];
                if (nlev > 1) sg += sg0[MW_MATW + k * 32];
                if (nlev > 2) sg += sg0[2 * MW_MATW + k * 32];
                const bool in = row && k < m;
                if (in && k <= lane && !(sg < (R)1e-9)) big = 1;
                // lower triangle only (zeros above the diagonal: the in-place inverse below relies on them)
                acol[k * 32 + lane] = (in && k <= lane) ? sg + __ldg(lo.cov + lane * m + k) : ((k == lane) ? (R)1 : (R)0);
            }
        }
        big = __any_sync(0xffffffffu, big);
        R* li = li_of(l - 1);
        if (big) {
            __syncwarp();
            // Cholesky, left-looking, in place: column j of the factor from the lane's own row and row j (broadcast)
            for (int j = 0; j < MW_M; j++) {
                R s0 = acol[j * 32 + lane], s1 = (R)0, s2 = (R)0, s3 = (R)0;
                const R* own = acol + lane;
                const R* piv = acol + j;
                int k = 0;
                for (; k + 3 < j; k += 4) {
                    s0 -= own[k * 32] * piv[k * 32];
                    s1 -= own[(k + 1) * 32] * piv[(k + 1) * 32];
                    s2 -= own[(k + 2) * 32] * piv[(k + 2) * 32];
                    s3 -= own[(k + 3) * 32] * piv[(k + 3) * 32];
                }
                for (; k < j; k++) s0 -= own[k * 32] * piv[k * 32];
                const R s = (s0 + s1) + (s2 + s3);
                R djj = __shfl_sync(0xffffffffu, s, j);
                // a non-positive pivot (near-singular bias covariance in this dtype) is clamped and flagged; the
                // reference's np.linalg.inv does not fail there (distributions.py:402)
                if (!(djj > (R)0)) { djj = (R)1e-30; if (lane == 0) *p.error_flag = 1; }
                const R inv = mw_rsqrt(djj);               // one reciprocal square root instead of a square root and a division
                djj = djj * inv;
                __syncwarp();
                if (lane >= j) acol[j * 32 + lane] = (lane == j) ? djj : s * inv;
                if (lane == j) dinv[j] = inv;
                __syncwarp();
            }
            // inverse of the factor, in place, last column first: Li[j][j] = 1 / L[j][j],
            // Li[i][j] = -(sum_{k = j+1..i} Li[i][k] L[k][j]) Li[j][j]   (rows of the already inverted trailing block)
            for (int j = MW_M - 1; j >= 0; j--) {
                const R dj = dinv[j];                      // 1 / L[j][j], kept from the factorisation
                R s0 = (R)0, s1 = (R)0, s2 = (R)0, s3 = (R)0;
                {
                    // every lane runs the loop of the last row (uniform trip count); rows above the diagonal of the
                    // trailing block hold zeros, so the extra terms of the shorter rows vanish
                    const R* own = acol + lane;
                    const R* col = acol + j * 32;
                    int k = j + 1;
                    for (; k + 3 < MW_M; k += 4) {
                        s0 -= own[k * 32] * col[k];
                        s1 -= own[(k + 1) * 32] * col[k + 1];
                        s2 -= own[(k + 2) * 32] * col[k + 2];
                        s3 -= own[(k + 3) * 32] * col[k + 3];
                    }
                    for (; k < MW_M; k++) s0 -= own[k * 32] * col[k];
                    s0 = (s0 + s1) + (s2 + s3);
                    s1 = (R)0;
                }
                __syncwarp();
                if (lane >= j) acol[j * 32 + lane] = (lane == j) ? dj : (s0 + s1) * dj;
                __syncwarp();
            }
            // -> the level's inverse-factor image (element (i, j) at j * 33 + i, zero above the diagonal)
            for (int j = 0; j < MW_M; j++) li[j * MW_LIS + lane] = (lane >= j) ? acol[j * 32 + lane] : (R)0;
            __syncwarp();
        }
        // re-score level l-1's current link with the new bias
        __syncwarp();
        xb[lane] = (lane < m) ? V(l - 1, V_F) + V(l - 1, V_BIAS) - __ldg(lo.data + lane) : (R)0;
        __syncwarp();
        const R nl = (R)-0.5 * quad_form(li);
        const int sid = I(l - 1, I_SID);
        __syncwarp();
        setS(l - 1, S_LIKE, nl);
        for (int aa = l; aa < L; aa++)
            if (I(aa, I_SID) == sid) setS(l - 1, S_SVL + aa, nl);
        __syncwarp();
    }

    // a step of level l >= 1 (chain.py:415-444, proposal.py:1583-1613)
    __device__ __forceinline__ void upper_accept(int l, R pr, R lk, R& s1, R& s2) {
        int acc = 0;
        if (I(l - 1, I_ACCSUB) > 0) {
            const R post_new = pr + lk, post_cur = S(l, S_PRIOR) + S(l, S_LIKE);
            const R post_below = S(l - 1, S_PRIOR) + S(l - 1, S_LIKE);
            const R x = post_new - post_cur + (S(l - 1, S_SVP + l) + S(l - 1, S_SVL + l)) - post_below;
            const R alpha = texp(x);
            const R u = uniform_at(ucur);
            ucur += 1;
            acc = (u < alpha) ? 1 : 0;
        }
        __syncwarp();
        if (acc) {
            V(l, V_TH) = prop[lane];
            V(l, V_F) = Fp[lane];
            if (lane == 0) {
                sc[l * NSC + S_PRIOR] = pr; sc[l * NSC + S_LIKE] = lk;
                isc[l * NIS + I_SID] = isc[(l - 1) * NIS + I_SID];
                isc[l * NIS + I_NACC] += 1;
                isc[l * NIS + I_ACCSUB] += 1;
            }
        }
        __syncwarp();
        record(l, acc);
        if (l == p.L - 1) { const R th = V(l, V_TH); s1 += th; s2 += th * th; }
        align(l, acc);
        if (p.aem) {
            aem_update(l, p.lvl_steps[l] + I(l, I_REC), acc);      // bias.t: steps of level l so far, this one included
            push_bias(l);
        }
    }
};

template <typename R>
__global__ void __launch_bounds__(MW_MAXW * 32, 1) mlda_warp_kernel(const __grid_constant__ Params<R> p, const __grid_constant__ MwParams mp) {
    extern __shared__ __align__(16) unsigned char mw_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int L = p.L, d = p.d;
    const size_t Cs = (size_t)p.Cs;
    // CTA-shared constants: proposal factor and whitening matrix as [j][32] images, Phi images of the coarse levels
    R* cta = reinterpret_cast<R*>(mw_smem);
    R* Ts = cta;
    R* LPs = cta + 32 * 32;
    for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
        const int j = e >> 5, k = e & 31;
        const bool in = j < d && k < d;
        Ts[e] = in ? p.T[(size_t)j * p.ldD + k] : (R)0;
        LPs[e] = in ? p.LP[(size_t)j * p.ldD + k] : (R)0;
    }
    for (int l = 0; l < L; l++)
        if (mp.phi_smem[l] >= 0) {
            const R* src = reinterpret_cast<const R*>(mp.phiw) + mp.phi_off[l];
            R* dst = cta + mp.phi_smem[l];
            for (int e = threadIdx.x; e < mp.phi_elems[l]; e += blockDim.x) dst[e] = src[e];
        }
    __syncthreads();
    bool diagT, diagLP;
    R dT, dLP;
    {
        bool ndT = false, ndL = false;
        for (int j = 0; j < d; j++) {
            if (lane != j && Ts[j * 32 + lane] != (R)0) ndT = true;
            if (lane != j && LPs[j * 32 + lane] != (R)0) ndL = true;
        }
        diagT = !__any_sync(0xffffffffu, ndT);
        diagLP = !__any_sync(0xffffffffu, ndL);
        dT = Ts[lane * 32 + lane];
        dLP = LPs[lane * 32 + lane];
    }
    R* base = cta + mp.cta_elems + (size_t)w * mw_warp_elems<R>();
    for (int g = blockIdx.x * mp.warps + w; g < p.C; g += gridDim.x * mp.warps) {
        Mw<R> c(p, lane, g);
        c.Ts = Ts; c.LPs = LPs; c.diagT = diagT; c.diagLP = diagLP; c.dT = dT; c.dLP = dLP;
        c.mpp = &mp; c.cta = cta;
        c.vec = base;
        c.sc = c.vec + MAXL * NVEC * 32;
        c.isc = reinterpret_cast<int*>(c.sc + MAXL * NSC);
        c.prop = c.sc + 64;
        c.xb = c.prop + 32;
        c.Fp = c.xb + 32;
        c.dinv = c.Fp + 32;
        c.li0 = c.dinv + 32;
        c.acol = c.li0 + MW_MATW;
        c.sigw = reinterpret_cast<R*>(mp.sigw) + (size_t)g * MAXL * MW_MATW;
        c.liw = reinterpret_cast<R*>(mp.liw) + (size_t)g * MAXL * MW_MATW;
        c.t_base = p.t_base;
        c.wcount = p.wcount;
        c.ucur = p.ucur[g];
        c.scal = p.scaling[g];
        c.win_sum = p.adaptive ? p.win_sum[g] : 0;
        c.win_old = 0;
        __syncwarp();
        // ---- chain state: global (chain-fastest arrays shared with the lock-step kernel) -> shared memory ----
        for (int l = 0; l < L; l++) {
            const LevelP<R>& v = p.lv[l];
            const int m = v.m;
            c.V(l, V_TH) = (lane < d) ? v.theta[(size_t)lane * Cs + g] : (R)0;
            c.V(l, V_F) = (v.need_F && lane < m) ? v.F[(size_t)lane * Cs + g] : (R)0;
            const bool ad = v.lik_kind == TDA_LIK_ADAPTIVE;
            c.V(l, V_BIAS) = (ad && lane < m) ? v.lik_bias[(size_t)lane * Cs + g] : (R)0;
            const bool up = p.aem && l >= 1;
            c.V(l, V_MD) = (up && lane < m) ? v.model_diff[(size_t)lane * Cs + g] : (R)0;
            c.V(l, V_BMU) = (up && lane < m) ? v.bias_mu[(size_t)lane * Cs + g] : (R)0;
            for (int a = 0; a < MAXL; a++)
                if (a >= 1) c.V(l, V_SVF + a - 1) = (v.need_F && a > l && a < L && lane < m) ? v.sv_F[a][(size_t)lane * Cs + g] : (R)0;
            if (lane == 0) {
                c.sc[l * NSC + S_PRIOR] = v.prior[g];
                c.sc[l * NSC + S_LIKE] = v.like[g];
                for (int a = 0; a < MAXL; a++) {
                    c.sc[l * NSC + S_SVP + a] = (a > l && a < L) ? v.sv_prior[a][g] : (R)0;
                    c.sc[l * NSC + S_SVL + a] = (a > l && a < L) ? v.sv_like[a][g] : (R)0;
                }
                c.isc[l * NIS + I_SID] = v.sid[g];
                c.isc[l * NIS + I_ACCSUB] = v.acc_sub[g];
                c.isc[l * NIS + I_NACC] = 0;
                c.isc[l * NIS + I_REC] = 0;
            }
        }
        if (p.lv[0].lik_kind == TDA_LIK_ADAPTIVE)
            for (int e = lane; e < MW_MATW; e += 32) c.li0[e] = c.liw[e];
        c.prop[lane] = (R)0;
        __syncwarp();
        R s1 = (R)0, s2 = (R)0;
        if (lane < d) { s1 = p.sum1[(size_t)lane * Cs + g]; s2 = p.sum2[(size_t)lane * Cs + g]; }

        // ---- the recursion of Tile::run ----
        // Tile::run's recursion as a loop over "the level that steps next": every level's link is built by the one
        // eval() below.  After a step of level l < L-1 its subchain counter advances; a full subchain (J[l] steps)
        // hands over to level l+1, anything else returns to the coarsest level.
        int cnt0 = 0, cnt1 = 0, cnt2 = 0, l = 0;
        long long it = 0;
        while (it < p.iterations) {
            if (l == 0) c.base_proposal();
            else { __syncwarp(); c.prop[lane] = c.V(l - 1, V_TH); }
            R pr = (R)0, lk = (R)0;
            c.eval(l, pr, lk);
            if (l == 0) c.base_accept(pr, lk, s1, s2);
            else c.upper_accept(l, pr, lk, s1, s2);
            if (l == L - 1) { it++; l = 0; continue; }
            int full;
            if (l == 0) { full = (++cnt0 == p.J[0]); if (full) cnt0 = 0; }
            else if (l == 1) { full = (++cnt1 == p.J[1]); if (full) cnt1 = 0; }
            else { full = (++cnt2 == p.J[2]); if (full) cnt2 = 0; }
            l = full ? l + 1 : 0;
        }

        // ---- write the chain state back ----
        __syncwarp();
        for (int l = 0; l < L; l++) {
            const LevelP<R>& v = p.lv[l];
            const int m = v.m;
            if (lane < d) v.theta[(size_t)lane * Cs + g] = c.V(l, V_TH);
            if (lane < m && v.need_F) {
                v.F[(size_t)lane * Cs + g] = c.V(l, V_F);
                if (v.lik_kind == TDA_LIK_ADAPTIVE) v.lik_bias[(size_t)lane * Cs + g] = c.V(l, V_BIAS);
                if (p.aem && l >= 1) {
                    v.model_diff[(size_t)lane * Cs + g] = c.V(l, V_MD);
                    v.bias_mu[(size_t)lane * Cs + g] = c.V(l, V_BMU);
                }
                for (int a = l + 1; a < L; a++) v.sv_F[a][(size_t)lane * Cs + g] = c.V(l, V_SVF + a - 1);
            }
            if (lane == 0) {
                v.prior[g] = c.sc[l * NSC + S_PRIOR];
                v.like[g] = c.sc[l * NSC + S_LIKE];
                for (int a = l + 1; a < L; a++) { v.sv_prior[a][g] = c.sc[l * NSC + S_SVP + a]; v.sv_like[a][g] = c.sc[l * NSC + S_SVL + a]; }
                v.sid[g] = c.isc[l * NIS + I_SID];
                v.acc_sub[g] = c.isc[l * NIS + I_ACCSUB];
                v.n_acc[g] += c.isc[l * NIS + I_NACC];
            }
        }
        if (p.lv[0].lik_kind == TDA_LIK_ADAPTIVE)
            for (int e = lane; e < MW_MATW; e += 32) c.liw[e] = c.li0[e];
        if (lane < d) { p.sum1[(size_t)lane * Cs + g] = s1; p.sum2[(size_t)lane * Cs + g] = s2; }
        if (lane == 0) {
            p.ucur[g] = c.ucur;
            if (p.adaptive) { p.scaling[g] = c.scal; p.win_sum[g] = c.win_sum - c.win_old; }
        }
        __syncwarp();
    }
}

// ---- layout conversion: chain-fastest matrices of the lock-step kernel <-> the warp-major images -----------------
// canon: [m * m][Cs] (element (i, j) of chain g at (i * m + j) * Cs + g).  image: [chain][MAXL][MW_MATW], element at
// e(i, j) = j * 32 + i (sigma = 1) or j * 33 + i (inverse factor).  One CTA moves 32 chains x 32 image elements at
// a time through shared memory so that both sides are accessed in 128-byte lines.
template <typename R>
__global__ void __launch_bounds__(256) mw_convert_kernel(R* canon, R* image, int m, int Cs, int C, int level, int sigma, int to_image) {
    __shared__ R tile[32][33];
    const int g0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 rows of 32 threads
    const int rs = sigma ? 32 : MW_LIS;
    for (int e0 = 0; e0 < MW_MATW; e0 += 32) {
        if (to_image) {
            for (int r = ty; r < 32; r += 8) {
                const int e = e0 + r, j = e / rs, i = e - j * rs;
                R val = (R)0;
                const bool in = i < m && j < m && i < 32;
                if (in && (sigma || i >= j) && g0 + tx < Cs) val = canon[((size_t)i * m + j) * Cs + g0 + tx];
                if (!sigma && !in && i == j && i < MW_M) val = (R)1;     // identity outside m x m
                tile[r][tx] = val;
            }
            __syncthreads();
            for (int r = ty; r < 32; r += 8)
                if (g0 + r < C && e0 + tx < MW_MATW) image[((size_t)(g0 + r) * MAXL + level) * MW_MATW + e0 + tx] = tile[tx][r];
            __syncthreads();
        } else {
            for (int r = ty; r < 32; r += 8)
                tile[tx][r] = (g0 + r < C && e0 + tx < MW_MATW) ? image[((size_t)(g0 + r) * MAXL + level) * MW_MATW + e0 + tx] : (R)0;
            __syncthreads();
            for (int r = ty; r < 32; r += 8) {
                const int e = e0 + r, j = e / rs, i = e - j * rs;
                if (i < m && j < m && i < 32 && (sigma || i >= j) && g0 + tx < C) canon[((size_t)i * m + j) * Cs + g0 + tx] = tile[r][tx];
            }
            __syncthreads();
        }
    }
}

// Phi^T [d][ldA] of a level -> image [k][t][32]: element ((k * stride + t) * 32 + lane) = Phi[k][lane * stride + t]
// (zero for the idle lanes beyond the last segment and for the rows that pad d to a multiple of four)
template <typename R>
__global__ void mw_phi_kernel(const R* A, int ldA, int d, int stride, int m, R* image, int elems) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= elems) return;
    const int lane = e & 31, kt = e >> 5, t = kt % stride, k = kt / stride;
    image[e] = (k < d && lane <= m) ? A[(size_t)k * ldA + lane * stride + t] : (R)0;
}

thread_local std::string g_mwerr;

template <typename R>
size_t mw_smem_bytes(int warps, int cta_elems) { return ((size_t)cta_elems + (size_t)warps * mw_warp_elems<R>()) * sizeof(R) + 16; }

}  // namespace

const char* mlda_warp_last_error() { return g_mwerr.c_str(); }

bool mlda_warp_eligible(const tda_config& c) {
    if (c.n_levels < 1 || c.n_levels > MAXL || c.d > 32 || c.mtm_k || c.randomize_subchain) return false;
    if (c.prop_kind != TDA_PROP_RWMH && c.prop_kind != TDA_PROP_PCN) return false;
    if (c.adaptive && c.period < 1) return false;
    if (c.aem != 0 && c.aem != 1) return false;
    for (int l = 0; l < c.n_levels; l++) {
        const tda_level_config& lc = c.level[l];
        if (lc.m > MW_M || lc.m < 1) return false;
        if (lc.model_kind == TDA_MODEL_POISSON1D) {
            if (lc.n_grid % (lc.m + 1) != 0) return false;
        } else if (lc.model_kind != TDA_MODEL_LINEAR) {
            return false;
        }
        if (lc.lik_kind == TDA_LIK_DENSE) return false;
        if (lc.lik_kind == TDA_LIK_ADAPTIVE && (l == c.n_levels - 1 || !c.aem)) return false;
        if (c.aem && l < c.n_levels - 1 && lc.lik_kind != TDA_LIK_ADAPTIVE) return false;
        if (l > 0 && c.aem && lc.m != c.level[0].m) return false;
    }
    return true;
}

template <typename R>
int mlda_warp_run(Params<R>& P, void* sigw, void* liw, void* phiw, int sm_count, cudaStream_t st) {
    const int L = P.L;
    MwParams mp;
    mp.sigw = sigw; mp.liw = liw; mp.phiw = phiw;
    // Phi images; the coarsest level(s) go to shared memory while they fit into 1024 elements
    int cta_elems = 2 * 32 * 32, off = 0;
    const int d4 = (P.d + 3) / 4 * 4;
    const int budget = cta_elems + 1024;     // the coarsest level of a cfg4-sized problem; more would cost resident warps
    for (int l = 0; l < MAXL; l++) { mp.phi_off[l] = 0; mp.phi_elems[l] = 0; mp.phi_smem[l] = -1; }
    for (int l = 0; l < L; l++) {
        const LevelP<R>& v = P.lv[l];
        if (v.model_kind != TDA_MODEL_POISSON1D) continue;        // linear levels read their operator directly
        const int elems = d4 * v.stride * 32;
        mp.phi_off[l] = off; mp.phi_elems[l] = elems;
        mw_phi_kernel<R><<<(elems + 255) / 256, 256, 0, st>>>(v.A, v.ldA, P.d, v.stride, v.m, reinterpret_cast<R*>(phiw) + off, elems);
        off += elems;
        if (cta_elems + elems <= budget) { mp.phi_smem[l] = cta_elems; cta_elems += elems; }
    }
    mp.cta_elems = cta_elems;
    int warps = MW_MAXW;
    while (warps > 1 && mw_smem_bytes<R>(warps, cta_elems) > 224 * 1024) warps--;
    mp.warps = warps;
    const size_t smem = mw_smem_bytes<R>(warps, cta_elems);
    cudaError_t e = cudaFuncSetAttribute(mlda_warp_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_mwerr = std::string("mlda warp kernel: ") + cudaGetErrorString(e); return -2; }
    const unsigned cgrid = (unsigned)((P.C + 31) / 32);
    // error-model matrices of the lock-step layout -> warp-major images
    if (P.aem) {
        for (int l = 0; l < L; l++) {
            const LevelP<R>& v = P.lv[l];
            if (l >= 1) mw_convert_kernel<R><<<cgrid, 256, 0, st>>>(v.bias_sigma, reinterpret_cast<R*>(sigw), v.m, P.Cs, P.C, l, 1, 1);
            if (v.lik_kind == TDA_LIK_ADAPTIVE) mw_convert_kernel<R><<<cgrid, 256, 0, st>>>(v.lik_prec, reinterpret_cast<R*>(liw), v.m, P.Cs, P.C, l, 0, 1);
        }
    }
    int grid = (P.C + warps - 1) / warps;
    if (grid > sm_count) grid = sm_count;
    mlda_warp_kernel<R><<<grid, warps * 32, smem, st>>>(P, mp);
    if (P.aem) {
        for (int l = 0; l < L; l++) {
            const LevelP<R>& v = P.lv[l];
            if (l >= 1) mw_convert_kernel<R><<<cgrid, 256, 0, st>>>(v.bias_sigma, reinterpret_cast<R*>(sigw), v.m, P.Cs, P.C, l, 1, 0);
            if (v.lik_kind == TDA_LIK_ADAPTIVE) mw_convert_kernel<R><<<cgrid, 256, 0, st>>>(v.lik_prec, reinterpret_cast<R*>(liw), v.m, P.Cs, P.C, l, 0, 0);
        }
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { g_mwerr = std::string("mlda warp kernel: ") + cudaGetErrorString(e); return -2; }
    return 0;
}

size_t mlda_warp_image_elems(int n_chains) { return (size_t)n_chains * MAXL * MW_MATW; }
size_t mlda_warp_phi_elems(const tda_config& c) {
    size_t n = 0;
    for (int l = 0; l < c.n_levels; l++)
        if (c.level[l].model_kind == TDA_MODEL_POISSON1D)
            n += (size_t)((c.d + 3) / 4 * 4) * (size_t)(c.level[l].n_grid / (c.level[l].m + 1)) * 32;
    return n ? n : 32;
}
int mlda_warp_launches(int L, int aem) {
    if (!aem) return 1 + L;
    return 1 + L + 2 * ((L - 1) + (L - 1));
}

template int mlda_warp_run<float>(Params<float>&, void*, void*, void*, int, cudaStream_t);
template int mlda_warp_run<double>(Params<double>&, void*, void*, void*, int, cudaStream_t);

}  // namespace tda
