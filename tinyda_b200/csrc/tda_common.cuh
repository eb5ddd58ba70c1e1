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
    R* lik_prec;       // [m][m][Cs]  per-chain inverse of (cov + bias covariance)
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
    int dream_M0, dream_delta, dream_nCR;
    long long dream_cap, dream_slots;
    R prior_logconst;
    const R* prior_mean;   // [d]
    const R* LP;           // [d][ldD]
    const R* Pprec;        // [d][ldD]
    const R* T;            // [d][ldD]  shared proposal factor
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
    R* grad;               // [d][Cs] MALA: gradient at the current state
    R* gradp;              // [d][Cs] MALA: gradient at the proposal
    R* archive;            // DREAM: [cap][Cg][d]
    R* sum1;               // [d][Cs] running sum of finest-level states
    R* sum2;               // [d][Cs] running sum of squares
    R* scratch;            // POISSON: [3][n_max][Cs]  (k-field, c', d')
    int n_max;
    int* error_flag;
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

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: random access by draw index --------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t hi0 = mulhi32(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = mulhi32(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ uint4 philox_block(unsigned long long seed, long long chain, uint32_t stream,
                                              unsigned long long block) {
    uint4 ctr = make_uint4((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)chain,
                           stream ^ (uint32_t)((unsigned long long)chain >> 32));
    uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    return philox4x32_10(ctr, key);
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
// error ~1e-6 on a N(0,1) variate, far below float32 MCMC noise) -- 7 instructions per normal
// instead of ~45 for the correctly-rounded logf / sincospif versions.  This IS the definition of
// the float engine's normal stream: every kernel and tda_fill_streams go through this function.
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <>
__device__ __forceinline__ void normals4<float>(uint4 b, float out[4]) {
    const float TWO_PI_24 = 6.283185307179586f * 5.9604644775390625e-08f;   // 2 pi / 2^24
    const float A0 = -3.141592653589793f + 0.5f * TWO_PI_24;                // angle in (-pi, pi)
    float r0 = mufu_sqrt(-1.3862943611198906f * mufu_lg2(u01<float>(b.x)));  // sqrt(-2 ln u)
    float r1 = mufu_sqrt(-1.3862943611198906f * mufu_lg2(u01<float>(b.z)));
    float a0 = fmaf((float)(b.y >> 8), TWO_PI_24, A0);
    float a1 = fmaf((float)(b.w >> 8), TWO_PI_24, A0);
    out[0] = r0 * mufu_cos(a0);
    out[1] = r0 * mufu_sin(a0);
    out[2] = r1 * mufu_cos(a1);
    out[3] = r1 * mufu_sin(a1);
}

template <typename R>
__device__ __forceinline__ R philox_normal(unsigned long long seed, long long chain, long long idx) {
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

}  // namespace tda
