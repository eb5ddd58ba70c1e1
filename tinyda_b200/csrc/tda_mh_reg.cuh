// Register-resident single-level Metropolis-Hastings kernel for small problems (d <= 8): one
// thread owns one chain for the whole launch, the chain state (parameters, densities, MALA
// gradient, step size, stream cursors, running moments) lives in registers, the shared constants
// (whitening matrix, proposal factor, forward operator, data) in shared memory where every lane of
// a warp reads the same address (broadcast), and the only per-step global traffic is the Link
// write-out: structure-of-arrays records, lane = chain, so every store instruction of a warp
// fills one 128-byte line.  The path is bound by that write-out and by the Philox / Box-Muller
// integer + special-function work (cfg3: 21 B and ~0.75 Philox blocks per transition).
//
// It advances exactly the state the generic lock-step kernel (tda_kernels.cuh) advances -- same
// buffers, same Philox / injected streams, same draw order -- so the two are interchangeable
// between launches and are checked against the same golden fixtures.
//
// Reference map (file:line into the reference's tinyDA/ package):
//   the loop body          Chain.sample chain.py:101-125
//   create_link            posterior.py:78-110 (log-prior, model, log-likelihood)
//   proposals              proposal.py:247-251 (RWMH), :349-355 (pCN), :948-959 (MALA drift)
//   acceptance             proposal.py:253-258, :357-362, :961-988
//   adaptation             proposal.py:228-245
#pragma once
#include "tda_common.cuh"

namespace tda {

constexpr int RG_NT = 128;     // threads (= chains) per CTA; equals the generic kernel's tile
constexpr int RG_MAXM = 256;   // outputs of a linear model held in shared memory

template <typename R, int D>
struct RegConsts {
    R mean[D], LP[D * D], Pp[D * D], T[D * D];
    R A[D * RG_MAXM], b[RG_MAXM], data[RG_MAXM], ivar[RG_MAXM];
};

// log-prior + forward model + log-likelihood (+ MALA gradient) of the state x
template <typename R, int D, bool MALA, int MODEL>
__device__ __forceinline__ void reg_create_link(const Params<R>& p, const RegConsts<R, D>& S, const R (&x)[D], R& prior,
                                                R& like, R& F0, R (&grad)[D]) {
    const int d = p.d;
    const LevelP<R>& v = p.lv[0];
    constexpr bool mala = MALA;
    R xm[D];
#pragma unroll
    for (int k = 0; k < D; k++) xm[k] = (k < d) ? x[k] - S.mean[k] : (R)0;
    R ssq = (R)0;
#pragma unroll
    for (int j = 0; j < D; j++) {
        if (j < d) {
            R w = (R)0;
#pragma unroll
            for (int k = 0; k < D; k++)
                if (k < d) w = fma(xm[k], S.LP[k * D + j], w);
            ssq = fma(w, w, ssq);
        }
    }
    prior = (R)-0.5 * (p.prior_logconst + ssq);
    R gm[D];
#pragma unroll
    for (int k = 0; k < D; k++) gm[k] = (R)0;
    if (MODEL == TDA_MODEL_ROSENBROCK) {
        const R a = v.sc0 - x[0], b = x[D > 1 ? 1 : 0] - x[0] * x[0];
        const R F = a * a + v.sc1 * b * b;
        F0 = F;
        const R res = F - S.data[0];
        like = (v.lik_kind == TDA_LIK_ISO) ? (R)-0.5 * (res * res) / v.lik_var : (R)-0.5 * (res * res * S.ivar[0]);
        if (mala) {
            const R sens = (S.data[0] - F) * S.ivar[0];
            const R dFdx = (R)-2 * a - (R)4 * v.sc1 * x[0] * b;
            const R dFdy = (R)2 * v.sc1 * b;
            gm[0] = sens * dFdx;
            if (D > 1) gm[D > 1 ? 1 : 0] = sens * dFdy;
        }
    } else {   // LINEAR
        const int m = v.m;
        R s = (R)0;
        for (int j = 0; j < m; j++) {
            R acc = (R)0;
#pragma unroll
            for (int k = 0; k < D; k++)
                if (k < d) acc = fma(x[k], S.A[k * RG_MAXM + j], acc);
            const R F = acc + S.b[j];
            const R res = F - S.data[j];
            if (v.lik_kind == TDA_LIK_ISO) s = fma(res, res, s);
            else s += res * res * S.ivar[j];
            if (mala) {
                const R sens = (S.data[j] - F) * S.ivar[j];
#pragma unroll
                for (int k = 0; k < D; k++)
                    if (k < d) gm[k] = fma(sens, S.A[k * RG_MAXM + j], gm[k]);
            }
        }
        like = (v.lik_kind == TDA_LIK_ISO) ? (R)-0.5 * s / v.lik_var : (R)-0.5 * s;
        F0 = (R)0;
    }
    if (mala) {   // grad log posterior = P (mu - x) + J_F^T sensitivity   (utils.py:272-287)
#pragma unroll
        for (int n = 0; n < D; n++) {
            if (n < d) {
                R w = (R)0;
#pragma unroll
                for (int k = 0; k < D; k++)
                    if (k < d) w = fma(xm[k], S.Pp[k * D + n], w);
                grad[n] = gm[n] - w;
            }
        }
    }
}

template <typename R, int D, bool MALA, int MODEL>
__global__ void __launch_bounds__(RG_NT) mh_reg_kernel(const __grid_constant__ Params<R> p) {
    __shared__ RegConsts<R, D> S;
    const int d = p.d, tid = threadIdx.x;
    const LevelP<R>& v = p.lv[0];
    constexpr bool mala = MALA;
    // ---- constants -> shared memory ----------------------------------------------------------
    for (int e = tid; e < D * D; e += RG_NT) {
        const int k = e / D, j = e - k * D;
        const bool in = k < d && j < d;
        S.LP[e] = in ? p.LP[k * p.ldD + j] : (R)0;
        S.T[e] = in ? p.T[k * p.ldD + j] : (R)0;
        S.Pp[e] = (in && mala) ? p.Pprec[k * p.ldD + j] : (R)0;
    }
    if (tid < D) S.mean[tid] = tid < d ? p.prior_mean[tid] : (R)0;
    const int m = v.m;
    for (int j = tid; j < m; j += RG_NT) {
        S.b[j] = v.b[j];
        S.data[j] = v.data[j];
        // DIAG: 1/var[j]; ISO: 1/var (used by the MALA sensitivity; the ISO log-likelihood divides
        // the sum once, like distributions.py:326)
        S.ivar[j] = (R)1 / (v.lik_kind == TDA_LIK_ISO ? v.lik_var : v.var[j]);
        if (MODEL == TDA_MODEL_LINEAR)
            for (int k = 0; k < d; k++) S.A[k * RG_MAXM + j] = v.A[(size_t)k * v.ldA + j];
    }
    __syncthreads();

    const int g = blockIdx.x * RG_NT + tid;      // chain (structure-of-arrays column)
    if (g >= p.Cs) return;
    const size_t Cs = (size_t)p.Cs;
    // ---- chain state -> registers --------------------------------------------------------------
    R th[D], grad[D], s1[D], s2[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
        th[k] = (k < d) ? v.theta[k * Cs + g] : (R)0;
        grad[k] = (k < d && mala) ? p.grad[k * Cs + g] : (R)0;
        s1[k] = (k < d) ? p.sum1[k * Cs + g] : (R)0;
        s2[k] = (k < d) ? p.sum2[k * Cs + g] : (R)0;
    }
    R prior = v.prior[g], like = v.like[g];
    R F0 = v.need_F ? v.F[g] : (R)0;             // Rosenbrock output (m = 1)
    R scal = p.scaling[g];
    const long long uc0 = p.ucur[g];
    long long n_acc = v.n_acc[g];
    // accepts since the last period boundary: at a boundary the reference's window
    // accepted[-period:] (proposal.py:234) is exactly the period that just ended, so no ring read
    // is needed inside the loop; the ring and its sum are kept for the generic kernel's sake
    int win_cnt = 0;
    int sid = v.sid[g];
    ChainStreams<R> rs(p, g);
    rs.seek_normal(p.t_base * d);
    rs.seek_uniform(uc0);
    // history cursors: record r of field f lives at f + r * stride; advanced by one record per step
    R* hp_theta = v.h_theta + ((size_t)p.rec[0] * d) * Cs + g;
    R* hp_prior = v.h_prior + (size_t)p.rec[0] * Cs + g;
    R* hp_like = v.h_like + (size_t)p.rec[0] * Cs + g;
    R* hp_F = v.h_F + (size_t)p.rec[0] * v.m * Cs + g;
    uint8_t* hp_acc = v.h_acc + (size_t)p.rec[0] * Cs + g;
    uint8_t* wp = p.win + (size_t)(p.t_base % p.period) * Cs + g;
    const size_t theta_stride = (size_t)d * Cs;
    // position in the adaptation period, kept incrementally (no 64-bit division per step)
    int wpos = (int)(p.t_base % p.period);
    long long kk = p.t_base / p.period;          // adaptations done so far
    R hs2 = (R)0.5 * scal * scal, qf = (R)-0.5 / (scal * scal);
    if (p.adaptive)
        for (int q = 0; q < wpos; q++) win_cnt += (int)p.win[(size_t)q * Cs + g];
    const bool store_F = (v.store & TDA_STORE_OUTPUT) && v.need_F && MODEL == TDA_MODEL_ROSENBROCK;
    const bool store_lin_F = (v.store & TDA_STORE_OUTPUT) && MODEL == TDA_MODEL_LINEAR;
    const bool pcn = p.prop_kind == TDA_PROP_PCN;
    const int store = v.store;
    const long long rec0 = p.rec[0], hist_cap = v.hist_cap;

    for (long long it = 0; it < p.iterations; it++) {
        const long long t = p.t_base + it;       // base-level steps done before this one
        // ---- proposal --------------------------------------------------------------------------
        R z[D], tp[D];
#pragma unroll
        for (int k = 0; k < D; k++) z[k] = (k < d) ? rs.normal() : (R)0;
        if (mala) {
#pragma unroll
            for (int k = 0; k < D; k++) tp[k] = th[k] + hs2 * grad[k] + scal * z[k];
        } else {
            const R ca = pcn ? tsqrt((R)1 - scal * scal) : (R)1;
#pragma unroll
            for (int j = 0; j < D; j++) {
                R xi = (R)0;
#pragma unroll
                for (int k = 0; k < D; k++)
                    if (k < d) xi = fma(z[k], S.T[k * D + j], xi);
                tp[j] = ca * th[j] + scal * xi;
            }
        }
        // ---- link of the proposal -----------------------------------------------------------------
        R pr, lk, Fp, gp[D];
#pragma unroll
        for (int k = 0; k < D; k++) gp[k] = (R)0;
        reg_create_link<R, D, MALA, MODEL>(p, S, tp, pr, lk, Fp, gp);
        // ---- acceptance --------------------------------------------------------------------------
        R x = pcn ? lk - like : (pr + lk) - (prior + like);
        if (mala) {
            R qxy = (R)0, qyx = (R)0;
#pragma unroll
            for (int k = 0; k < D; k++) {
                if (k < d) {
                    const R a = th[k] - tp[k] - hs2 * gp[k];
                    const R b = tp[k] - th[k] - hs2 * grad[k];
                    qxy = fma(a, a, qxy);
                    qyx = fma(b, b, qyx);
                }
            }
            x = x + qf * qxy - qf * qyx;
        }
        const R alpha = tisnan(pr + lk) ? (R)0 : texp(x);
        const R u = rs.uniform();
        const int acc = (u < alpha) ? 1 : 0;
        if (acc) {
#pragma unroll
            for (int k = 0; k < D; k++) { th[k] = tp[k]; grad[k] = gp[k]; }
            prior = pr; like = lk; F0 = Fp;
            sid = (int)(t + 1);
            n_acc += 1;
        }
        if (p.adaptive) {        // the `accepted` window of proposal.adapt (proposal.py:234)
            *wp = (uint8_t)acc;
            wp += Cs;
            win_cnt += acc;
        }
        // ---- Link write-out (lane = chain: one 128-byte line per store instruction) ---------------
        const long long r = rec0 + it;
        if (r < hist_cap) {
            if (store & TDA_STORE_THETA) {
#pragma unroll
                for (int k = 0; k < D; k++)
                    if (k < d) hp_theta[k * Cs] = th[k];
                hp_theta += theta_stride;
            }
            if (store_F) { *hp_F = F0; hp_F += Cs; }
            if (store_lin_F) {
                // Link.model_output of a linear model: recomputed from the current state with the
                // expression create_link used (same bits), m coalesced stores
                for (int j = 0; j < v.m; j++) {
                    R acc_f = (R)0;
#pragma unroll
                    for (int k = 0; k < D; k++)
                        if (k < d) acc_f = fma(th[k], S.A[k * RG_MAXM + j], acc_f);
                    hp_F[(size_t)j * Cs] = acc_f + S.b[j];
                }
                hp_F += (size_t)v.m * Cs;
            }
            if (store & TDA_STORE_STATS) { *hp_prior = prior; *hp_like = like; hp_prior += Cs; hp_like += Cs; }
            if (store & TDA_STORE_ACCEPT) { *hp_acc = (uint8_t)acc; hp_acc += Cs; }
        }
#pragma unroll
        for (int k = 0; k < D; k++) { s1[k] += th[k]; s2[k] += th[k] * th[k]; }
        // ---- adaptive global scaling -------------------------------------------------------------
        if (++wpos == p.period) {                // (t + 1) % period == 0
            wpos = 0;
            wp = p.win + g;
            if (p.adaptive) {
                const R rate = (R)win_cnt / (R)p.period;
                scal = texp(tlog(scal) + tpow(p.gamma, (R)(-(double)kk)) * (rate - p.alpha_star));
                hs2 = (R)0.5 * scal * scal;
                qf = (R)-0.5 / (scal * scal);
            }
            kk += 1;
            win_cnt = 0;
        }
    }
    // ---- registers -> chain state ------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < D; k++) {
        if (k < d) {
            v.theta[k * Cs + g] = th[k];
            if (mala) p.grad[k * Cs + g] = grad[k];
            p.sum1[k * Cs + g] = s1[k];
            p.sum2[k * Cs + g] = s2[k];
        }
    }
    v.prior[g] = prior; v.like[g] = like;
    if (v.need_F && MODEL == TDA_MODEL_ROSENBROCK) v.F[g] = F0;
    p.scaling[g] = scal;
    p.ucur[g] = (long long)rs.ui;
    v.n_acc[g] = n_acc;
    v.sid[g] = sid;
    if (p.adaptive) {
        // sum of the ring = this period's entries + the previous period's entries not yet overwritten
        int tail = 0;
        if (p.t_base + p.iterations >= p.period)
            for (int q = wpos; q < p.period; q++) tail += (int)p.win[(size_t)q * Cs + g];
        p.win_sum[g] = win_cnt + tail;
    }
}

// which configurations the register kernel advances
inline bool mh_reg_eligible(const tda_config& c, bool need_F_level0) {
    if (c.n_levels != 1 || c.d > 8 || c.mtm_k) return false;
    if (c.prop_kind != TDA_PROP_RWMH && c.prop_kind != TDA_PROP_PCN && c.prop_kind != TDA_PROP_MALA) return false;
    const tda_level_config& lc = c.level[0];
    if (lc.lik_kind != TDA_LIK_ISO && lc.lik_kind != TDA_LIK_DIAG) return false;
    if (lc.model_kind == TDA_MODEL_ROSENBROCK) return true;
    if (lc.model_kind != TDA_MODEL_LINEAR || lc.m > RG_MAXM) return false;
    (void)need_F_level0;
    return true;
}

template <typename R>
inline cudaError_t mh_reg_launch(const Params<R>& P, cudaStream_t st) {
    const int grid = P.Cs / RG_NT;
    const bool mala = P.prop_kind == TDA_PROP_MALA;
    constexpr int LIN = TDA_MODEL_LINEAR, ROS = TDA_MODEL_ROSENBROCK;
    if (P.lv[0].model_kind == TDA_MODEL_ROSENBROCK) {
        if (mala) mh_reg_kernel<R, 2, true, ROS><<<grid, RG_NT, 0, st>>>(P);
        else mh_reg_kernel<R, 2, false, ROS><<<grid, RG_NT, 0, st>>>(P);
    } else if (P.d <= 2) {
        if (mala) mh_reg_kernel<R, 2, true, LIN><<<grid, RG_NT, 0, st>>>(P);
        else mh_reg_kernel<R, 2, false, LIN><<<grid, RG_NT, 0, st>>>(P);
    } else if (P.d <= 4) {
        if (mala) mh_reg_kernel<R, 4, true, LIN><<<grid, RG_NT, 0, st>>>(P);
        else mh_reg_kernel<R, 4, false, LIN><<<grid, RG_NT, 0, st>>>(P);
    } else {
        if (mala) mh_reg_kernel<R, 8, true, LIN><<<grid, RG_NT, 0, st>>>(P);
        else mh_reg_kernel<R, 8, false, LIN><<<grid, RG_NT, 0, st>>>(P);
    }
    return cudaGetLastError();
}

}  // namespace tda
