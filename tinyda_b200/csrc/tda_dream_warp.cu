// DREAM(Z) / DREAM step kernel, one WARP per chain (BASELINE cfg5 class): single level, linear forward
// operator, isotropic / diagonal Gaussian likelihood, d <= 32 parameters, non-adaptive crossover.
// Reference semantics: proposal.py:608-852 (DREAMZ), :1627-1656 + ray.py:366-384 (DREAM, shared archive),
// chain.py:78-129 (the MH loop) -- draw for draw what Tile::propose_dream / Tile::base_step / Tile::adapt of
// the lock-step kernel (tda_kernels.cuh) do, which the golden fixtures pin to the reference.
//
// Why a second kernel: the lock-step kernel walks a 128-chain tile through a dozen block-wide phases per step
// (~80 us per step whatever the chain count: at cfg5's 1024 chains per GPU eight CTAs on 148 SMs).  Here lane k
// of a warp owns parameter k of one chain: the ~2 delta + 2 d + 2 uniforms and d normals of a proposal are
// drawn lane-parallel (counter-based streams give random access), the two gathered archive rows are one
// coalesced 128-byte load each, the forward model is d warp shuffles x m/32 FMAs per lane against the operator
// in shared memory, and a step of a chain is a few hundred instructions of one warp.  The launch is persistent
// over all steps; the shared-archive variant closes every step with the grid barrier / NVLink peer-memory
// handshake of the lock-step kernel (Tile::step_barrier).
#include <cstdlib>
#include <string>

#include "tda_dream_warp.h"

namespace tda {

namespace {

constexpr int DW_THREADS = 256;      // 8 warps per CTA
constexpr int DW_CPW = 4;            // chains per warp at most (state in registers)

template <typename R>
struct DwShared {
    // dynamic shared memory: A [d][m_pad] | LP [d][32] | b [m_pad] | data [m_pad] | var [m_pad] | mean [32]   (zero padded)
    R* A; R* LP; R* b; R* data; R* var; R* mean;
    int m_pad;
};

template <typename R>
__device__ __forceinline__ R dw_uniform(const Params<R>& p, int g, long long gchain, long long idx) {
    if (p.rng_mode == TDA_RNG_INJECTED) return (g < p.C && idx < p.ulen) ? p.us[(size_t)g * p.ulen + idx] : (R)0.5;
    return philox_uniform<R>(p.seed, gchain, idx);
}
template <typename R>
__device__ __forceinline__ R dw_normal(const Params<R>& p, int g, long long gchain, long long idx) {
    if (p.rng_mode == TDA_RNG_INJECTED) return (g < p.C && idx < p.zlen) ? p.zs[(size_t)g * p.zlen + idx] : (R)0;
    return philox_normal<R>(p.seed, gchain, idx, p.z_round);
}

template <typename R>
__device__ __forceinline__ R dw_warp_sum(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// eight consecutive values from (16-byte aligned) shared memory
template <typename R> struct DwVec8;
template <> struct DwVec8<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&o)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    }
};
template <> struct DwVec8<double> {
    static __device__ __forceinline__ void ld(const double* p, double (&o)[8]) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double2 a = *reinterpret_cast<const double2*>(p + 2 * i);
            o[2 * i] = a.x; o[2 * i + 1] = a.y;
        }
    }
};

// output index handled by (lane, j) in the 256-output chunk starting at n0: eight CONSECUTIVE outputs per lane, so that
// a row of the operator is two 16-byte shared-memory loads per lane and chunk
__device__ __forceinline__ int dw_out(int n0, int lane, int j) { return n0 + lane * 8 + j; }

// log-prior and log-likelihood of the parameter vector held one component per lane (lanes >= d hold 0).
// Shared-memory operands are zero-padded (A, LP, b, data; var with ones) to m_pad = multiple of 256 outputs and 32
// parameter columns: no bounds tests in the contractions.
template <typename R>
__device__ __forceinline__ void dw_eval(const Params<R>& p, const DwShared<R>& s, int lane, int g, R th, R& prior, R& like) {
    const LevelP<R>& v = p.lv[0];
    const int d = p.d, m = v.m, mp = s.m_pad;
    // prior: -0.5 * (logconst + |(x - mu) LP|^2)   (posterior.py:92, scipy's whitening matrix)
    {
        const R xc = (lane < d) ? th - s.mean[lane] : (R)0;
        R y0 = (R)0, y1 = (R)0;
        const R* lp = s.LP + lane;
        int k = 0;
        for (; k + 1 < d; k += 2) {
            y0 = fma(__shfl_sync(0xffffffffu, xc, k), lp[k * 32], y0);
            y1 = fma(__shfl_sync(0xffffffffu, xc, k + 1), lp[(k + 1) * 32], y1);
        }
        if (k < d) y0 = fma(__shfl_sync(0xffffffffu, xc, k), lp[k * 32], y0);
        const R y = y0 + y1;
        prior = (R)-0.5 * (p.prior_logconst + dw_warp_sum<R>(y * y));
    }
    // model F = theta @ A + b (posterior.py:95), residual against the data, Gaussian log-likelihood
    R ssq = (R)0;
    for (int n0 = 0; n0 < mp; n0 += 256) {
        R acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = (R)0;
        const R* col = s.A + n0 + lane * 8;
        for (int k = 0; k < d; k++) {
            const R tk = __shfl_sync(0xffffffffu, th, k);
            R a[8];
            DwVec8<R>::ld(col + (size_t)k * mp, a);
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = fma(tk, a[j], acc[j]);
        }
        R bb[8], dd[8], vv[8];
        DwVec8<R>::ld(s.b + n0 + lane * 8, bb);
        DwVec8<R>::ld(s.data + n0 + lane * 8, dd);
        if (v.lik_kind != TDA_LIK_ISO) DwVec8<R>::ld(s.var + n0 + lane * 8, vv);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const R F = acc[j] + bb[j];
            if (v.need_F) {     // Link.model_output of the proposal (the same lane reads it back)
                const int n = dw_out(n0, lane, j);
                if (n < m) v.Fp[(size_t)n * p.Cs + g] = F;
            }
            const R res = F - dd[j];
            if (v.lik_kind == TDA_LIK_ISO) ssq = fma(res, res, ssq);
            else ssq += res * res / vv[j];
        }
    }
    ssq = dw_warp_sum<R>(ssq);
    like = (v.lik_kind == TDA_LIK_ISO) ? (R)-0.5 * ssq / v.lik_var : (R)-0.5 * ssq;
}

// end of a lock-step step of the shared-archive ensemble (same protocol as Tile::step_barrier)
template <typename R>
__device__ __forceinline__ void dw_step_barrier(const Params<R>& p, unsigned step) {
    __syncthreads();
    if (p.arrive_mode) {
        // several GPUs, every rank on this kernel: ONE system-scope fence per CTA, then an arrival on every GPU's
        // counter of this rank (remote reductions over NVLink, own GPU included), then wait until every rank's CTAs
        // have all arrived here.  No local barrier in front, no second fence behind it.
        if (threadIdx.x == 0) {
            __threadfence_system();
            for (int r = 0; r < p.n_peers; r++)
                asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p.peer_flags[r] + 8 + p.my_rank), "r"(1u) : "memory");
            for (int r = 0; r < p.n_peers; r++) {
                const unsigned target = (p.arr_base + step + 1u) * (unsigned)p.peer_grid[r];
                const unsigned* f = p.peer_flags[p.my_rank] + 8 + r;
                unsigned v;
                do {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                    if ((int)(v - target) < 0) __nanosleep(20);
                } while ((int)(v - target) < 0);
            }
        }
        __syncthreads();
        return;
    }
    if (threadIdx.x == 0) {
        if (p.n_peers > 1) __threadfence_system();
        else __threadfence();
        atomicAdd(p.grid_bar, 1u);
        const unsigned target = (step + 1u) * gridDim.x;
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.grid_bar) : "memory");
            if (v < target) __nanosleep(20);
        } while (v < target);
        if (p.n_peers > 1) {
            const unsigned flag = p.flag_base + step + 1u;
            if (blockIdx.x == 0) {
                // one system-scope fence, then the eight flag stores back to back: a release store per peer would wait
                // for the previous peer's store to be acknowledged over NVLink (measured: 19 us per step on 8 GPUs)
                __threadfence_system();
                for (int r = 0; r < p.n_peers; r++)
                    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.peer_flags[r] + p.my_rank), "r"(flag) : "memory");
            }
            for (int r = 0; r < p.n_peers; r++) {
                const unsigned* f = p.peer_flags[p.my_rank] + r;
                do {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                    if ((int)(v - flag) < 0) __nanosleep(40);
                } while ((int)(v - flag) < 0);
            }
        }
    }
    __syncthreads();
}

// OCC = resident CTAs per SM the register budget is cut for: 2 (128 registers) is the fast single-chain-per-warp
// variant; 3 (80 registers, a few spills in float) is selected when the chains would otherwise queue up three and
// four deep on a warp (cfg5's 8192 chains on ONE GPU).
template <typename R, int OCC>
__global__ void __launch_bounds__(DW_THREADS, OCC) dream_warp_kernel(const __grid_constant__ Params<R> p) {
    extern __shared__ __align__(16) unsigned char dw_smem[];
    const LevelP<R>& v = p.lv[0];
    const int d = p.d, m = v.m;
    DwShared<R> s;
    const int mp = (m + 255) / 256 * 256;
    s.m_pad = mp;
    s.A = reinterpret_cast<R*>(dw_smem);
    s.LP = s.A + (size_t)d * mp;
    s.b = s.LP + (size_t)d * 32;
    s.data = s.b + mp;
    s.var = s.data + mp;
    s.mean = s.var + mp;
    for (int i = threadIdx.x; i < d * mp; i += DW_THREADS) {
        const int k = i / mp, n = i - k * mp;
        s.A[i] = (n < m) ? v.A[(size_t)k * v.ldA + n] : (R)0;
    }
    for (int i = threadIdx.x; i < d * 32; i += DW_THREADS) {
        const int k = i >> 5, c = i & 31;
        s.LP[i] = (c < d) ? p.LP[(size_t)k * p.ldD + c] : (R)0;
    }
    for (int i = threadIdx.x; i < mp; i += DW_THREADS) {
        s.b[i] = (i < m) ? v.b[i] : (R)0;
        s.data[i] = (i < m) ? v.data[i] : (R)0;
        s.var[i] = (i < m && v.lik_kind == TDA_LIK_DIAG) ? v.var[i] : (R)1;
    }
    for (int i = threadIdx.x; i < 32; i += DW_THREADS) s.mean[i] = (i < d) ? p.prior_mean[i] : (R)0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wglobal = blockIdx.x * (DW_THREADS / 32) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (DW_THREADS / 32);
    const bool shared_arch = p.prop_kind == TDA_PROP_DREAM;
    const int delta = p.dream_delta, nCR = p.dream_nCR;
    const size_t Cs = (size_t)p.Cs;

    // chains of this warp: wglobal, wglobal + nwarps, ... (at most DW_CPW, checked on the host); state in registers
    R th[DW_CPW], s1[DW_CPW], s2[DW_CPW], prior[DW_CPW], like[DW_CPW], scal[DW_CPW];
    long long ucur[DW_CPW];
    int nacc[DW_CPW], lastacc[DW_CPW];
    int cpw = 0;
#pragma unroll
    for (int q = 0; q < DW_CPW; q++) {
        const int g = wglobal + q * nwarps;
        th[q] = s1[q] = s2[q] = prior[q] = like[q] = scal[q] = (R)0;
        ucur[q] = 0; nacc[q] = 0; lastacc[q] = 0;
        if (g < p.C) {
            cpw = q + 1;
            if (lane < d) {
                th[q] = v.theta[(size_t)lane * Cs + g];
                s1[q] = p.sum1[(size_t)lane * Cs + g];
                s2[q] = p.sum2[(size_t)lane * Cs + g];
            }
            prior[q] = v.prior[g]; like[q] = v.like[g]; scal[q] = p.scaling[g];
            ucur[q] = p.ucur[g];
        }
    }

    long long slots = p.dream_slots;
    const long long Ksync = p.dream_sync;       // shared archive: rows become visible at multiples of Ksync steps
    unsigned nbar = 0;                           // barriers passed in this launch
    for (long long it = 0; it < p.iterations; it++) {
        const long long t = p.t_base + it;                       // base-level steps done before this one
#pragma unroll
        for (int q = 0; q < DW_CPW; q++) {
            if (q >= cpw) break;
            const int g = wglobal + q * nwarps;
            const long long gchain = p.chain_offset + g;
            // ---- proposal (proposal.py:811-852) ----
            // rows per chain every chain may draw from: all rows (own archive), or the rows through the last
            // synchronisation point (shared archive; Ksync = 1: through the previous step)
            const long long nslots = shared_arch ? slots - ((slots - p.dream_M0) % Ksync) : slots;
            const long long M = shared_arch ? nslots * p.Cg : nslots;
            long long uc = ucur[q];
            // lanes 0 .. 2 delta - 1: the pair draws; lane 2 delta: the crossover draw
            const R u_head = (lane <= 2 * delta) ? dw_uniform<R>(p, g, gchain, uc + lane) : (R)0;
            uc += 2 * delta + 1;
            const R ucr = __shfl_sync(0xffffffffu, u_head, 2 * delta);
            int mCR = 0;
            {
                double cs = 0.0;
                for (int i = 0; i < nCR; i++) {
                    cs += 1.0 / nCR;
                    if (cs <= (double)ucr) mCR++;
                }
                if (mCR > nCR - 1) mCR = nCR - 1;
            }
            const R CR = (R)(mCR + 1) / (R)nCR;
            const R um = (lane < d) ? dw_uniform<R>(p, g, gchain, uc + lane) : (R)2;
            uc += d;
            unsigned mask = __ballot_sync(0xffffffffu, lane < d && um < CR);
            int card = __popc(mask);
            if (card == 0) {
                const R ux = dw_uniform<R>(p, g, gchain, uc);
                uc += 1;
                int k = (int)tfloor(ux * (R)d);
                if (k > d - 1) k = d - 1;
                mask = 1u << k; card = 1;
            }
            const R gam = scal[q] * (R)2.38 / tsqrt((R)(2 * delta * card));
            R zsum = (R)0, psum = (R)0;                             // sum Z[r1], sum Z[r2] (proposal.py:818-826)
            for (int i = 0; i < delta; i++) {
                const R u1 = __shfl_sync(0xffffffffu, u_head, 2 * i), u2 = __shfl_sync(0xffffffffu, u_head, 2 * i + 1);
                long long a = (long long)tfloor(u1 * (R)M); if (a > M - 1) a = M - 1;
                long long b = (long long)tfloor(u2 * (R)(M - 1)); if (b > M - 2) b = M - 2;
                if (b >= a) b++;
                long long ga, sa, gb, sb;
                if (shared_arch) { ga = a / nslots; sa = a - ga * nslots; gb = b / nslots; sb = b - gb * nslots; }
                else { ga = g; sa = a; gb = g; sb = b; }
                if (lane < d) {     // L2 reads: the rows of other CTAs / GPUs were written during this launch
                    zsum += __ldcg(p.archive + ((size_t)sa * p.Cg + ga) * d + lane);
                    psum += __ldcg(p.archive + ((size_t)sb * p.Cg + gb) * d + lane);
                }
            }
            R prop = th[q];
            {
                const R ue = (lane < d) ? dw_uniform<R>(p, g, gchain, uc + lane) : (R)0;
                const R zn = (lane < d) ? dw_normal<R>(p, g, gchain, t * d + lane) : (R)0;
                const R e = -p.dream_b + (p.dream_b + p.dream_b) * ue;
                const R eps = p.dream_b_star * zn;
                const R dz = zsum - psum;
                if ((mask >> lane) & 1u) prop = th[q] + (((R)1 + e) * gam * dz + eps);
                uc += d;
            }
            // ---- Link of the proposal (posterior.py:78-110) and the accept test (chain.py:105-117) ----
            R pr, lk;
            dw_eval<R>(p, s, lane, g, (lane < d) ? prop : (R)0, pr, lk);
            const R x = (pr + lk) - (prior[q] + like[q]);
            const R alpha = tisnan(pr + lk) ? (R)0 : texp(x);
            const R u = dw_uniform<R>(p, g, gchain, uc);
            uc += 1;
            ucur[q] = uc;
            const bool acc = u < alpha;
            if (acc) {
                th[q] = prop; prior[q] = pr; like[q] = lk; nacc[q]++; lastacc[q] = (int)(t + 1);
                if (v.need_F)
                    for (int n0 = 0; n0 < m; n0 += 256)
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const int n = dw_out(n0, lane, j);
                            if (n < m) v.F[(size_t)n * Cs + g] = v.Fp[(size_t)n * Cs + g];
                        }
            }
            // ---- record (chain.chain / chain.accepted) and running moments ----
            {
                const long long r = p.rec[0] + it;
                if (r < v.hist_cap) {
                    if ((v.store & TDA_STORE_THETA) && lane < d) v.h_theta[((size_t)r * d + lane) * Cs + g] = th[q];
                    if ((v.store & TDA_STORE_OUTPUT) && v.need_F)
                        for (int n0 = 0; n0 < m; n0 += 256)
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int n = dw_out(n0, lane, j);
                                if (n < m) v.h_F[((size_t)r * m + n) * Cs + g] = v.F[(size_t)n * Cs + g];
                            }
                    if (lane == 0) {
                        if (v.store & TDA_STORE_STATS) { v.h_prior[(size_t)r * Cs + g] = prior[q]; v.h_like[(size_t)r * Cs + g] = like[q]; }
                        if (v.store & TDA_STORE_ACCEPT) v.h_acc[(size_t)r * Cs + g] = (uint8_t)acc;
                    }
                }
                s1[q] += th[q];
                s2[q] += th[q] * th[q];
            }
            // ---- adapt(): the current state joins the archive (proposal.py:794 / :1652) ----
            if (slots < p.dream_cap && lane < d) {
                const size_t o = ((size_t)slots * p.Cg + p.arch_off + g) * d + lane;
                p.archive[o] = th[q];
                for (int rk = 0; rk < p.n_peers; rk++)
                    if (rk != p.my_rank) p.peer_archive[rk][o] = th[q];
            }
        }
        slots += 1;
        if (p.grid_sync && ((slots - p.dream_M0) % Ksync) == 0) dw_step_barrier<R>(p, nbar++);
    }

    // ---- write the chain state back (layout shared with the lock-step kernel) ----
#pragma unroll
    for (int q = 0; q < DW_CPW; q++) {
        if (q >= cpw) break;
        const int g = wglobal + q * nwarps;
        if (lane < d) {
            v.theta[(size_t)lane * Cs + g] = th[q];
            p.sum1[(size_t)lane * Cs + g] = s1[q];
            p.sum2[(size_t)lane * Cs + g] = s2[q];
        }
        if (lane == 0) {
            v.prior[g] = prior[q]; v.like[g] = like[q];
            p.ucur[g] = ucur[q];
            v.n_acc[g] += nacc[q];
            v.acc_sub[g] += nacc[q];
            if (nacc[q]) v.sid[g] = lastacc[q];
        }
    }
}

template <typename R>
size_t dw_smem_bytes(const Params<R>& P) {
    const LevelP<R>& v = P.lv[0];
    const size_t mp = ((size_t)v.m + 255) / 256 * 256;
    return ((size_t)P.d * mp + (size_t)P.d * 32 + 3 * mp + 32) * sizeof(R) + 16;
}

thread_local std::string g_dwerr;

}  // namespace

const char* dream_warp_last_error() { return g_dwerr.c_str(); }

namespace {
template <typename R, int OCC>
int dw_max_blocks(size_t smem, int sm_count) {
    cudaFuncSetAttribute(dream_warp_kernel<R, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dream_warp_kernel<R, OCC>, DW_THREADS, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return 0;
    }
    return per_sm * sm_count;
}
}  // namespace

template <typename R>
int dream_warp_grid(const Params<R>& P, int sm_count, int* occ) {
    const size_t smem = dw_smem_bytes<R>(P);
    if (smem > 100 * 1024) return 0;
    const int wpb = DW_THREADS / 32;
    const int want = (P.C + wpb - 1) / wpb;                      // one chain per warp if they all fit ...
    int variant = 2;
    int max_blocks = dw_max_blocks<R, 2>(smem, sm_count);
    if (max_blocks < 1) return 0;
    if (sizeof(R) == 4 && want > 2 * max_blocks && !getenv("TDA_DREAM_WARP_OCC2")) {
        // more than two chains per warp: the variant with three resident CTAs per SM keeps more chains in flight
        const int m3 = dw_max_blocks<R, 3>(smem, sm_count);
        if (m3 > max_blocks) { max_blocks = m3; variant = 3; }
    }
    int blocks = want < max_blocks ? want : max_blocks;          // ... else several chains per warp, all CTAs resident
    if ((long long)blocks * wpb * DW_CPW < P.C) return 0;        // too many chains for the register-resident state
    if (occ) *occ = variant;
    return blocks;
}

template <typename R>
int dream_warp_launch(Params<R>& P, int grid, int occ, cudaStream_t st) {
    const size_t smem = dw_smem_bytes<R>(P);
    cudaError_t e;
    if (occ == 3) {
        e = cudaFuncSetAttribute(dream_warp_kernel<R, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) dream_warp_kernel<R, 3><<<grid, DW_THREADS, smem, st>>>(P);
    } else {
        e = cudaFuncSetAttribute(dream_warp_kernel<R, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) dream_warp_kernel<R, 2><<<grid, DW_THREADS, smem, st>>>(P);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { g_dwerr = std::string("dream warp kernel: ") + cudaGetErrorString(e); return -2; }
    return 0;
}

template int dream_warp_grid<float>(const Params<float>&, int, int*);
template int dream_warp_grid<double>(const Params<double>&, int, int*);
template int dream_warp_launch<float>(Params<float>&, int, int, cudaStream_t);
template int dream_warp_launch<double>(Params<double>&, int, int, cudaStream_t);

}  // namespace tda
