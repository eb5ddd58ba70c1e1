// Shared device-side definitions: kernel parameter blocks, Philox4x32-10 streams, math
// wrappers selected by the engine dtype.  Part of the tinyda_b200 CUDA engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/tinyda_b200.h"

namespace tda {

constexpr int NT = 256;    // threads per CTA
constexpr int TC = 128;    // chains per CTA tile (lanes run over chains -> coalesced SoA access)
constexpr int RM = 2;      // chains per thread in the tile contraction
constexpr int RN = 16;     // output columns per thread in the tile contraction
constexpr int CW = 4;      // column-warps
constexpr int NB = CW * RN;  // columns per staged chunk of the shared operand (64)
constexpr int MAXL = TDA_MAX_LEVELS;
constexpr int MAXD = TDA_MAX_D;
constexpr int MAX_DELTA = 8;
constexpr int MAX_NCR = 8;
constexpr int MAX_MTM = 16;

__host__ __device__ __forceinline__ bool is_dream(int kind) { return kind == TDA_PROP_DREAMZ || kind == TDA_PROP_DREAM; }
// proposals whose acceptance is the likelihood ratio (proposal.py:357-362, inherited by :515)
__host__ __device__ __forceinline__ bool is_pcn_like(int kind) { return kind == TDA_PROP_PCN || kind == TDA_PROP_OWPCN; }

constexpr int MODE_INIT = 0;
constexpr int MODE_RUN = 1;

constexpr uint32_t STREAM_Z = 0x5a5a0001u;
constexpr uint32_t STREAM_U = 0x5a5a0002u;

// ---- per-level device pointers (SoA, chain index fastest, stride Cs) ---------------------
template <typename R>
struct LevelP {
    int model_kind, m, n_grid, lik_kind, need_F, store, stride, ldA;
    R lik_var, sc0, sc1;
    long long hist_cap;
    const R* A;        // LINEAR: G^T [d][ldA]; POISSON: Phi^T [d][ldA]
    const R* A2;       // LINEAR: G [m][ldA2] (MALA gradient), ldA2 = NB-padded d
    const R* b;        // offset [m]
    const R* data;     // [m]
    const R* var;      // DIAG [m]
    const R* prec;     // DENSE: shared precision [m][m]
    const R* cov;      // ADAPTIVE: base covariance [m][m]
    // current state of the level
    R* theta;          // [d][Cs]
    R* prior;          // [Cs]
    R* like;           // [Cs]
    R* F;              // [m][Cs]   (need_F)
    R* Fp;             // [m][Cs]   proposal's model output (need_F)
    int* sid;          // [Cs] identity of the state (replaces Python object identity)
    // saved[a]: latest version of this level's link whose parameters are level a's state
    R* sv_prior[MAXL];
    R* sv_like[MAXL];
    R* sv_F[MAXL];
    // adaptive error model
    R* lik_bias;       // [m][Cs]     bias currently set on this level's likelihood
    R* lik_prec;       // [m][m][Cs]  per-chain Li = inv(chol(cov + bias covariance)), lower triangular
    R* bias_mu;        // [m][Cs]     moments of F_l - F_{l-1}  (levels >= 1)
    R* bias_sigma;     // [m][m][Cs]
    R* model_diff;     // [m][Cs]
    // history (local records)
    R* h_theta;        // [cap][d][Cs]
    R* h_prior;        // [cap][Cs]
    R* h_like;         // [cap][Cs]
    R* h_F;            // [cap][m][Cs]
    uint8_t* h_acc;    // [cap][Cs]
    long long* n_acc;  // [Cs] accepted local steps
    int* acc_sub;      // [Cs] local accepts in the current subchain
};

template <typename R>
struct Params {
    int mode, L, d, aem, rng_mode, prop_kind, adaptive, period, am_t0, am_device_refactor;
    int z_round;               // Philox normals on the fp16 grid ("z16" stream, float engine)
    int randomize;             // DA: randomize_subchain_length (chain.py:310-321)
    int aem_G;                 // chains per group of the cooperative bias-covariance factorisation
    int m_adapt;               // largest output count of a level with an adaptive likelihood (0 if none)
    int J[MAXL];
    int C, Cs, n_tiles;
    long long chain_offset, Cg, arch_off;   // arch_off: first archive column owned by this engine
    unsigned long long seed;
    long long iterations;
    // resumable counters (uniform over chains because chains advance in lock-step)
    long long t_base;          // base-level steps done  (= proposal.t)
    long long wcount;          // entries appended to the level-0 `accepted` list
    long long rec[MAXL];       // local records written per level
    long long lvl_steps[MAXL]; // steps done per level (bias.t - 1 for levels >= 1)
    // proposal
    R gamma, alpha_star, am_sd, am_eps, dream_b, dream_b_star;
    R scaling0;                // proposal.scaling at construction (restored by init)
    int dream_M0, dream_delta, dream_nCR;
    int dream_sync;            // shared archive: rows become visible at multiples of this many steps (>= 1)
    long long dream_cap, dream_slots;
    R prior_logconst;
    const R* prior_mean;   // [d]
    const R* LP;           // [d][ldD]
    const R* Pprec;        // [d][ldD]
    const R* T;            // [d][ldD]  shared proposal factor
    const R* Sop;          // [d][ldD]  OWPCN state operator (transposed); adaptive OWPCN: V
    const R* Sop2;         // [d][ldD]  adaptive OWPCN: V^T
    const R* ow_lambda;    // [d]       adaptive OWPCN: eigenvalues of B
    int ldD;
    R* scaling;            // [Cs]
    uint8_t* win;          // [period][Cs] ring of the last `period` accept flags
    int* win_sum;          // [Cs]
    long long* ucur;       // [Cs] uniform cursor
    const R* zs;           // injected normals [C][zlen]
    const R* us;           // injected uniforms [C][ulen]
    long long zlen, ulen;
    R* am_mu;              // [d][Cs]
    R* am_sigma;           // [d][d][Cs]
    R* am_T;               // [d][d][Cs]  per-chain factor
    R* qcur;               // [Cs] IndependenceSampler: log q (up to its constant) of the current state
    R* grad;               // [d][Cs] MALA: gradient at the current state
    R* gradp;              // [d][Cs] MALA: gradient at the proposal
    R* archive;            // DREAM: [cap][Cg][d]
    // DREAM(Z) with adaptive=True: crossover distribution (proposal.py:797-809), per chain
    R* dream_pCR;          // [nCR][Cs]
    R* dream_DeltaCR;      // [nCR][Cs]
    R* dream_LCR;          // [nCR][Cs]
    int* dream_mCR;        // [Cs] crossover index of the last proposal
    R* arch_s1;            // [d][Cs] column sums of the chain's local archive (for np.var(Z, axis=0))
    R* arch_s2;            // [d][Cs] column sums of squares
    R* sum1;               // [d][Cs] running sum of finest-level states
    R* sum2;               // [d][Cs] running sum of squares
    // MultipleTry (ray.py:213-354): k candidates per base-level step
    int mtm_k, mtm_include_current;
    long long* zcur;       // [Cs] normal cursor (consumption depends on the chain's control flow)
    R* mt_theta;           // [k][d][Cs] candidates
    R* mt_prior;           // [k][Cs]
    R* mt_like;            // [k][Cs]
    R* mt_w;               // [k][Cs] log-weights of the candidates, then of the reference points
    R* mt_F;               // [k][m0][Cs] (need_F)
    R* mt_y;               // [d][Cs] the chosen candidate (source of the reference points)
    int* mt_sel;           // [Cs] its index; -1: all weights were -inf / NaN posterior (alpha = 0)
    R* mt_lse;             // [Cs] logsumexp of the candidate weights
    // randomize_subchain_length: the link of the running coarse subchain that will be promoted
    int* promo_j;          // [Cs] 1-based coarse step after which it is taken
    R* pm_theta;           // [d][Cs]
    R* pm_prior;           // [Cs]
    R* pm_like;            // [Cs]
    R* pm_F;               // [m0][Cs] (need_F)
    int* pm_sid;           // [Cs]
    R* scratch;            // POISSON: [3][n_max][Cs]  (k-field, c', d')
    int n_max;
    int* error_flag;
    // DREAM with the shared archive as ONE persistent launch: every step ends with a grid-wide barrier (all
    // CTAs co-resident, one tile each) and, across GPUs, with the exchange of the step's new archive rows
    // through peer memory (NVLink stores into every replica) plus a flag handshake
    int grid_sync;                 // 1: barrier after every base-level step
    unsigned int* grid_bar;        // device counter (zeroed before the launch)
    int n_peers, my_rank;          // > 1: archive replicas of the other ranks are mapped (cudaIpc)
    R* peer_archive[8];            // [rank] -> that rank's archive replica (own entry = archive)
    unsigned int* peer_flags[8];   // [rank] -> that rank's flag array [8] (one slot per writer rank)
    unsigned int flag_base;        // flags count steps across launches: value after step t of this launch = flag_base + t + 1
    // warp-per-chain DREAM kernel on several GPUs: every CTA of every rank adds 1 to arrival counter [8 + its rank] of
    // every GPU's flag array after each step (one fence, no local barrier first); a step is complete on a GPU when
    // counter [8 + r] has reached (arr_base + t + 1) * peer_grid[r] for every rank r
    int arrive_mode;
    unsigned int arr_base;
    int peer_grid[8];
    LevelP<R> lv[MAXL];
};

// ---- math wrappers -------------------------------------------------------------------------
__device__ __forceinline__ float tsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double tsqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float texp(float x) { return expf(x); }
__device__ __forceinline__ double texp(double x) { return exp(x); }
__device__ __forceinline__ float tlog(float x) { return logf(x); }
__device__ __forceinline__ double tlog(double x) { return log(x); }
__device__ __forceinline__ float tpow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double tpow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float tfloor(float x) { return floorf(x); }
__device__ __forceinline__ double tfloor(double x) { return floor(x); }
__device__ __forceinline__ void tsincospi(float x, float* s, float* c) { sincospif(x, s, c); }
__device__ __forceinline__ void tsincospi(double x, double* s, double* c) { sincospi(x, s, c); }
__device__ __forceinline__ bool tisnan(float x) { return isnan(x); }
__device__ __forceinline__ bool tisnan(double x) { return isnan(x); }

// ---- packed fp32 pairs (FFMA2 / FMUL2 on sm_100): two lanes of fp32 math per instruction -------
__device__ __forceinline__ unsigned long long f2pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long f2mul(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: random access by draw index --------
// 32 x 32 -> 64-bit product split into halves: one IMAD.WIDE on the device
__host__ __device__ __forceinline__ void mulwide32(uint32_t a, uint32_t m, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    uint64_t p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(m));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
#else
    const uint64_t p = (uint64_t)a * (uint64_t)m;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

// Rounds: 10 for the uniform (accept-test) streams; TDA_PHILOX_Z_ROUNDS for the normal streams.  Philox4x32-7
// is the smallest round count Salmon et al. report as Crush-resistant (BigCrush clean); 10 is their default
// with a safety margin.  The normal generator is the pacing item of the tensor-core kernels (30 % of all
// instructions, half of them Philox rounds), see DESIGN.md section 4.2.
#ifndef TDA_PHILOX_Z_ROUNDS
#define TDA_PHILOX_Z_ROUNDS 10
#endif
template <int ROUNDS>
__host__ __device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < ROUNDS; i++) {
        uint32_t hi0, lo0, hi1, lo1;
        mulwide32(ctr.x, M0, hi0, lo0);
        mulwide32(ctr.z, M1, hi1, lo1);
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) { return philox4x32<10>(ctr, key); }

// The same function with the round keys key + i * (W0, W1) precomputed (they depend on the seed only): a kernel that
// holds them in its parameter block reads them as constant-bank operands of the XORs and saves the two key additions
// of every round -- 20 of ~70 instructions per block on the generator warps of the tensor-core kernels.
struct PhiloxRoundKeys {
    uint32_t k[2 * 10];
};
inline PhiloxRoundKeys philox_round_keys(unsigned long long seed) {
    PhiloxRoundKeys r;
    uint32_t kx = (uint32_t)seed, ky = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; i++) { r.k[2 * i] = kx; r.k[2 * i + 1] = ky; kx += 0x9E3779B9u; ky += 0xBB67AE85u; }
    return r;
}
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32_rk(uint4 ctr, const PhiloxRoundKeys& rk) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int i = 0; i < ROUNDS; i++) {
        uint32_t hi0, lo0, hi1, lo1;
        mulwide32(ctr.x, M0, hi0, lo0);
        mulwide32(ctr.z, M1, hi1, lo1);
        ctr = make_uint4(hi1 ^ ctr.y ^ rk.k[2 * i], lo1, hi0 ^ ctr.w ^ rk.k[2 * i + 1], lo0);
    }
    return ctr;
}
__device__ __forceinline__ uint4 philox_block_z_rk(const PhiloxRoundKeys& rk, long long chain, unsigned long long block) {
    uint4 ctr = make_uint4((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)chain,
                           STREAM_Z ^ (uint32_t)((unsigned long long)chain >> 32));
    return philox4x32_rk<TDA_PHILOX_Z_ROUNDS>(ctr, rk);
}

__device__ __forceinline__ uint4 philox_block(unsigned long long seed, long long chain, uint32_t stream,
                                              unsigned long long block) {
    uint4 ctr = make_uint4((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)chain,
                           stream ^ (uint32_t)((unsigned long long)chain >> 32));
    uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    // `stream` is a compile-time constant at every call site: the branch folds away
    return stream == STREAM_Z ? philox4x32<TDA_PHILOX_Z_ROUNDS>(ctr, key) : philox4x32<10>(ctr, key);
}

template <typename R> __device__ __forceinline__ R u01(uint32_t x);
template <> __device__ __forceinline__ double u01<double>(uint32_t x) {
    return ((double)x + 0.5) * 2.3283064365386963e-10;   // (x + 1/2) / 2^32  in (0,1)
}
template <> __device__ __forceinline__ float u01<float>(uint32_t x) {
    return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f;   // ((x>>8) + 1/2) / 2^24
}

// four standard normals from one Philox block (Box-Muller on two pairs)
template <typename R>
__device__ __forceinline__ void normals4(uint4 b, R out[4]) {
    R r0 = tsqrt((R)-2 * tlog(u01<R>(b.x)));
    R r1 = tsqrt((R)-2 * tlog(u01<R>(b.z)));
    R s0, c0, s1, c1;
    tsincospi((R)2 * u01<R>(b.y), &s0, &c0);
    tsincospi((R)2 * u01<R>(b.w), &s1, &c1);
    out[0] = r0 * c0;
    out[1] = r0 * s0;
    out[2] = r1 * c1;
    out[3] = r1 * s1;
}

// float engine: Box-Muller on the special-function unit (lg2 / sqrt / sin / cos .approx, absolute
// error ~1e-6 on a N(0,1) variate, far below float32 MCMC noise).  Both uniforms are built from the
// low 23 bits of a Philox word with one logic op (no int->float conversion): f = 1.mantissa in
// [1, 2); radius uniform u = 2 - f in (0, 1], angle = (f - 1.5) * 2 pi in [-pi, pi).
// This IS the definition of the float engine's normal stream: every kernel and tda_fill_streams go
// through bm_pair().
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr float BM_C_PLAIN = -1.3862943611198906f;                 // -2 ln 2        -> r
constexpr float BM_C_X4096 = -1.3862943611198906f * 16777216.0f;   // -2 ln 2 * 2^24 -> 4096 r
constexpr float Z16_SCALE = 4096.0f, Z16_UNSCALE = 1.0f / 4096.0f;

// two normals (times sqrt(c / (-2 ln 2))) from two Philox words
__device__ __forceinline__ void bm_pair(uint32_t w_radius, uint32_t w_angle, float c, float& n0, float& n1) {
    const float f = __uint_as_float((w_radius & 0x007FFFFFu) | 0x3F800000u);
    const float g = __uint_as_float((w_angle & 0x007FFFFFu) | 0x3F800000u);
    const float r = mufu_sqrt(c * mufu_lg2(2.0f - f));
    const float a = fmaf(g, 6.283185307179586f, -9.42477796076938f);
    n0 = r * mufu_cos(a);
    n1 = r * mufu_sin(a);
}

template <>
__device__ __forceinline__ void normals4<float>(uint4 b, float out[4]) {
    bm_pair(b.x, b.y, BM_C_PLAIN, out[0], out[1]);
    bm_pair(b.z, b.w, BM_C_PLAIN, out[2], out[3]);
}

// "z16" normal stream of the float engine (used by the fp16-split tensor-core kernel and by every
// other kernel / tda_fill_streams when Params::z_round is set).
//  * bits: normals 16g .. 16g+15 of a chain come from Philox blocks 3g, 3g+1, 3g+2 of its Z stream;
//    their 384 bits are cut into sixteen 24-bit fields (bm_pair uses the low 23 bits of a field),
//    pair p of the group takes field 2p (radius) and 2p+1 (angle) -> normals 2p (cos), 2p+1 (sin).
//    12 blocks per 64 normals instead of 16: the Philox multiplies dominate the generator's cost.
//  * value: generated at scale 4096, rounded to the nearest fp16 and scaled back, i.e. on an
//    11-bit-mantissa grid, so that it is ONE exact fp16 tensor-core operand (rounding is unbiased).
__device__ __forceinline__ float z16_round(float scaled) {
    unsigned short h;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(scaled));
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
    return f * Z16_UNSCALE;
}
__device__ __forceinline__ void z16_fields4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t f[4]) {
    f[0] = w0;
    f[1] = __funnelshift_r(w0, w1, 24);
    f[2] = __funnelshift_r(w1, w2, 16);
    f[3] = w2 >> 8;
}
// sixteen normals at scale 4096 (not yet rounded) from the three blocks of a group
__device__ __forceinline__ void z16_group_scaled(uint4 b0, uint4 b1, uint4 b2, float s[16]) {
    uint32_t f[16];
    z16_fields4(b0.x, b0.y, b0.z, f);
    z16_fields4(b0.w, b1.x, b1.y, f + 4);
    z16_fields4(b1.z, b1.w, b2.x, f + 8);
    z16_fields4(b2.y, b2.z, b2.w, f + 12);
#pragma unroll
    for (int pr = 0; pr < 8; pr++) bm_pair(f[2 * pr], f[2 * pr + 1], BM_C_X4096, s[2 * pr], s[2 * pr + 1]);
}
__device__ __forceinline__ float philox_normal_z16(unsigned long long seed, long long chain, long long idx) {
    const unsigned long long g = (unsigned long long)idx >> 4;
    float s[16];
    z16_group_scaled(philox_block(seed, chain, STREAM_Z, 3 * g), philox_block(seed, chain, STREAM_Z, 3 * g + 1),
                     philox_block(seed, chain, STREAM_Z, 3 * g + 2), s);
    float v = s[0];
#pragma unroll
    for (int i = 1; i < 16; i++) v = ((int)(idx & 15) == i) ? s[i] : v;
    return z16_round(v);
}

template <typename R>
__device__ __forceinline__ R philox_normal(unsigned long long seed, long long chain, long long idx, int z_round = 0) {
    if (z_round) return (R)philox_normal_z16(seed, chain, idx);
    R v[4];
    normals4<R>(philox_block(seed, chain, STREAM_Z, (unsigned long long)idx >> 2), v);
    return v[idx & 3];
}

template <typename R>
__device__ __forceinline__ R philox_uniform(unsigned long long seed, long long chain, long long idx) {
    uint4 b = philox_block(seed, chain, STREAM_U, (unsigned long long)idx >> 2);
    uint32_t x = (idx & 3) == 0 ? b.x : (idx & 3) == 1 ? b.y : (idx & 3) == 2 ? b.z : b.w;
    return u01<R>(x);
}

// ---- sequential per-chain stream cursors (shared by the register-resident kernel and the DREAM
// proposal of the lock-step kernel) -------------------------------------------------------------
template <typename R>
struct PairGen;
template <>
struct PairGen<float> {
    static __device__ __forceinline__ void pair(const uint4& b, int half, float& n0, float& n1) {
        if (half == 0) bm_pair(b.x, b.y, BM_C_PLAIN, n0, n1);
        else bm_pair(b.z, b.w, BM_C_PLAIN, n0, n1);
    }
};
template <>
struct PairGen<double> {
    static __device__ __forceinline__ void pair(const uint4& b, int half, double& n0, double& n1) {
        const uint32_t wr = half == 0 ? b.x : b.z, wa = half == 0 ? b.y : b.w;
        double r = tsqrt(-2.0 * tlog(u01<double>(wr)));
        double s, c;
        tsincospi(2.0 * u01<double>(wa), &s, &c);
        n0 = r * c;
        n1 = r * s;
    }
};

// per-thread view of the chain's two streams.  The kernel consumes both strictly sequentially, so
// the generator keeps a cursor per stream and decides from its low bits when a new Philox block
// (4 normals or 4 uniforms) or Box-Muller pair is due; seek_*() primes the cached block when the
// launch starts in the middle of one.
template <typename R>
struct ChainStreams {
    const Params<R>& p;
    long long chain;       // global chain id (Philox counter word)
    int local;             // chain index on this engine (injected streams)
    uint4 zb, ub;
    unsigned long long zi, ui;     // next normal / uniform index
    R n0, n1;
    __device__ ChainStreams(const Params<R>& p_, int local_) : p(p_), chain(p_.chain_offset + local_), local(local_) {}
    __device__ __forceinline__ void seek_normal(long long idx) {
        zi = (unsigned long long)idx;
        if (p.rng_mode == TDA_RNG_INJECTED) return;
        if (zi & 3) zb = philox_block(p.seed, chain, STREAM_Z, zi >> 2);
        if (zi & 1) PairGen<R>::pair(zb, (int)((zi >> 1) & 1), n0, n1);
    }
    __device__ __forceinline__ void seek_uniform(long long idx) {
        ui = (unsigned long long)idx;
        if (p.rng_mode == TDA_RNG_INJECTED) return;
        if (ui & 3) ub = philox_block(p.seed, chain, STREAM_U, ui >> 2);
    }
    __device__ __forceinline__ R normal() {
        const unsigned long long idx = zi++;
        if (p.rng_mode == TDA_RNG_INJECTED)
            return (local < p.C && (long long)idx < p.zlen) ? p.zs[(size_t)local * p.zlen + idx] : (R)0;
        const unsigned o = (unsigned)idx & 3u;
        if (o == 0) zb = philox_block(p.seed, chain, STREAM_Z, idx >> 2);
        if ((o & 1u) == 0) PairGen<R>::pair(zb, (int)(o >> 1), n0, n1);
        return (o & 1u) ? n1 : n0;
    }
    __device__ __forceinline__ R uniform() {
        const unsigned long long idx = ui++;
        if (p.rng_mode == TDA_RNG_INJECTED)
            return (local < p.C && (long long)idx < p.ulen) ? p.us[(size_t)local * p.ulen + idx] : (R)0.5;
        const unsigned o = (unsigned)idx & 3u;
        if (o == 0) ub = philox_block(p.seed, chain, STREAM_U, idx >> 2);
        const uint32_t x = o == 0 ? ub.x : o == 1 ? ub.y : o == 2 ? ub.z : ub.w;
        return u01<R>(x);
    }
};

}  // namespace tda
