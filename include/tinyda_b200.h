/*
 * tinyda_b200 -- C ABI of the B200-native batched MCMC engine.
 *
 * The reference (tinyDA, pure Python) has no FFI of its own: its seams are Python protocols.
 * This header is the boundary a maintainer would bind (ctypes / cffi / pybind) to replace the
 * reference's per-chain Python loops with one device engine.  Each entry point names the
 * reference interface it replaces (file:line into the reference's tinyDA/ package):
 *
 *   tda_engine_create     sampler.py:21-290   argument normalisation + chain construction
 *                         (chain.py:37-76 Chain.__init__, :185-321 DAChain.__init__,
 *                          :570-678 MLDAChain.__init__, proposal.py:1339-1440 MLDA.__init__)
 *   tda_upload            the constants those constructors capture (prior, likelihood data /
 *                         covariance, model operator, proposal covariance factor, initial
 *                         parameters sampler.py:196-209, DREAM initial archive proposal.py:788)
 *   tda_engine_init       the initial Links (posterior.py:78-110 create_link on every level),
 *                         AEM set-up (chain.py:272-305, :644-678, proposal.py:1442-1467)
 *   tda_engine_run        Chain.sample chain.py:78-129 / DAChain.sample :325-444 /
 *                         MLDAChain.sample :680-769 with MLDA.make_*_proposal
 *                         proposal.py:1502-1613, for ALL chains in lock-step (replaces
 *                         ray.py:12-210, one actor per chain).  Re-entrant: calling it again
 *                         continues the chains, like calling .sample() again does.
 *   tda_fetch             the Link lists chain.chain / chain_fine / compress(chain_coarse,
 *                         is_coarse) / compress(chain, is_local) (sampler.py:305-309,
 *                         :406-439, :510-547) as structure-of-arrays
 *   tda_get / tda_set     proposal state the reference keeps on the proposal object
 *                         (scaling, AM moments, DREAM archive; proposal.py:228-245, :502-512,
 *                         :790-809) and accept counters (chain.accepted)
 *   tda_fill_streams      np.random.* (proposal.py:249, :353, :834-848, :958; chain.py:112,
 *                         :385, :429, :729): exports the engine's counter-based streams so a
 *                         CPU oracle can be fed the identical draws
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * negative code on failure, with a message available from tda_last_error(); no C++ exception
 * crosses the boundary; no host callbacks from inside tda_engine_run.  All host-side constant
 * uploads are float64 and are converted to the engine dtype on the device side of the call.
 * The engine owns its device memory; callers own every host pointer they pass.
 */
#ifndef TINYDA_B200_H
#define TINYDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDA_ABI_VERSION 3
#define TDA_MAX_LEVELS 4
#define TDA_MAX_D 64

/* dtype */
#define TDA_F32 0
#define TDA_F64 1
/* rng_mode */
#define TDA_RNG_PHILOX 0    /* per-chain Philox4x32-10 counter streams generated in-kernel  */
#define TDA_RNG_INJECTED 1  /* normals / uniforms read from uploaded per-chain streams      */
/* proposal kinds (reference class in parentheses) */
#define TDA_PROP_RWMH 0     /* GaussianRandomWalk  proposal.py:132  */
#define TDA_PROP_PCN 1      /* CrankNicolson       proposal.py:261  */
#define TDA_PROP_AM 2       /* AdaptiveMetropolis  proposal.py:372  */
#define TDA_PROP_MALA 3     /* MALA                proposal.py:861  */
#define TDA_PROP_DREAMZ 4   /* DREAMZ              proposal.py:608  */
#define TDA_PROP_DREAM 5    /* DREAM (shared)      proposal.py:1627 */
#define TDA_PROP_OWPCN 6    /* OperatorWeightedCrankNicolson proposal.py:515; with adaptive != 0 the operators
                             * follow each chain's step size through the eigen-decomposition of B     */
/* likelihood kinds */
#define TDA_LIK_ISO 0       /* IsotropicGaussianLogLike distributions.py:318 */
#define TDA_LIK_DIAG 1      /* DiagonalGaussianLogLike  distributions.py:304 */
#define TDA_LIK_DENSE 2     /* DefaultGaussianLogLike   distributions.py:246 */
#define TDA_LIK_ADAPTIVE 3  /* AdaptiveGaussianLogLike  distributions.py:332 */
/* forward model kinds (device-resident replacements of the user's Python callable) */
#define TDA_MODEL_LINEAR 0
#define TDA_MODEL_ROSENBROCK 1
#define TDA_MODEL_POISSON1D 2
/* history storage flags per level */
#define TDA_STORE_THETA 1
#define TDA_STORE_STATS 2   /* log-prior, log-likelihood */
#define TDA_STORE_OUTPUT 4  /* model output F(theta)     */
#define TDA_STORE_ACCEPT 8

typedef struct tda_level_config {
    int32_t model_kind;
    int32_t m;              /* number of model outputs / observations          */
    int32_t n_grid;         /* Poisson1D: number of cells                      */
    int32_t lik_kind;
    double lik_var;         /* isotropic variance                              */
    double model_scalars[4];
    int32_t store;          /* TDA_STORE_* flags                               */
    int32_t n_qoi;          /* quantities of interest the model returns next to its output
                             * (posterior.py:95-105: a tuple -> (output, qoi)); 0 = none  */
    int64_t hist_capacity;  /* records per chain the level's history can hold  */
} tda_level_config;

typedef struct tda_config {
    int32_t abi_version;
    int32_t dtype;
    int32_t n_levels;
    int32_t d;
    int32_t subchain[TDA_MAX_LEVELS];   /* J[l] for l < n_levels-1                     */
    int32_t aem;                        /* 0 none, 1 state-independent (chain.py:485-499, proposal.py:1442-1467),
                                         * 2 state-dependent (two levels; chain.py:446-473, :501-522) */
    int32_t rng_mode;
    int32_t randomize_subchain;         /* DAChain randomize_subchain_length, chain.py:310-321, :369, :525-527 */
    int32_t mtm_k;                      /* > 0: MultipleTry with k tries around the proposal kernel (ray.py:213-354;
                                         * RWMH / AM / pCN kernels, 2 <= k <= 16) */
    int32_t mtm_include_current;        /* 0 = the reference's k-1 reference points; 1 = the current state is
                                         * the k-th reference point (Liu et al. 2000; detailed balance)      */
    int32_t dream_sync_every;           /* DREAM (shared archive): the chains' view of the archive is refreshed every so many
                                         * steps (0 or 1: every step, the lock-step rule of the golden fixtures; larger values
                                         * need kernel 6).  Occupies the field that was `reserved0` (must-be-zero) in ABI 3. */
    uint64_t seed;
    int64_t n_chains;                   /* chains on THIS device                       */
    int64_t chain_offset;               /* global index of local chain 0 (Philox key)  */
    int64_t n_chains_global;            /* DREAM shared archive width                  */
    /* proposal */
    int32_t prop_kind;
    int32_t adaptive;
    int32_t period;
    int32_t am_t0;
    double scaling;
    double gamma;
    double alpha_star;
    double am_sd;
    double am_eps;
    int32_t am_device_refactor;         /* 1: Cholesky refactor on device; 0: host uploads factors */
    int32_t dream_M0;
    int32_t dream_delta;
    int32_t dream_nCR;
    double dream_b;
    double dream_b_star;
    int64_t dream_capacity;             /* archive slots per chain (M0 + max steps)    */
    /* injected streams */
    int64_t stream_z_len;
    int64_t stream_u_len;
    /* prior: logpdf = -0.5*(logconst + |(x-mean) @ LP|^2) */
    double prior_logconst;
    tda_level_config level[TDA_MAX_LEVELS];
} tda_config;

/* tda_upload 'what' */
#define TDA_UP_PRIOR_MEAN 1     /* [d]                                              */
#define TDA_UP_PRIOR_LP 2       /* [d][d] whitening matrix                          */
#define TDA_UP_PRIOR_PREC 3     /* [d][d] inverse covariance (MALA drift)           */
#define TDA_UP_PROP_T 4         /* [d][d] factor T, xi = z @ T                      */
#define TDA_UP_MODEL_A 5        /* level: LINEAR G^T [d][m]; POISSON Phi^T [d][n]   */
#define TDA_UP_MODEL_B 6        /* level: LINEAR offset [m]                         */
#define TDA_UP_LIK_DATA 7       /* level: [m]                                       */
#define TDA_UP_LIK_VAR 8        /* level: DIAG variances [m]                        */
#define TDA_UP_LIK_PREC 9       /* level: DENSE inverse covariance [m][m]; ADAPTIVE: Li = inv(chol(cov)),
                                 * lower triangular [m][m] (Li^T Li = inv(cov))     */
#define TDA_UP_LIK_COV 10       /* level: ADAPTIVE covariance [m][m]                */
#define TDA_UP_INIT_THETA 11    /* [n_chains][d]                                    */
#define TDA_UP_STREAM_Z 12      /* [n_chains][stream_z_len] standard normals        */
#define TDA_UP_STREAM_U 13      /* [n_chains][stream_u_len] U(0,1)                  */
#define TDA_UP_DREAM_ARCHIVE0 14 /* [n_chains_global][M0][d]                        */
#define TDA_UP_AM_FACTORS 15    /* [n_chains][d][d] per-chain T                     */
#define TDA_UP_PROP_S 16        /* OWPCN: [d][d] state operator, transposed: theta' = theta @ S + z @ T,
                                 * S = sqrtm(I - scaling*B)^T, T = svd_factor(prior cov) @ sqrtm(scaling*B)^T.
                                 * Adaptive OWPCN (B = V diag(lambda) V^T): S = V, T = svd_factor(prior cov) @ V */
#define TDA_PROP_INDEP 7    /* IndependenceSampler proposal.py:64 with a multivariate normal q: upload its SVD factor
                             * as TDA_UP_PROP_T, its whitening matrix as TDA_UP_PROP_S, its mean as TDA_UP_PROP_LAMBDA */
#define TDA_UP_PROP_S2 17       /* adaptive OWPCN: [d][d] V^T                        */
#define TDA_UP_PROP_LAMBDA 18   /* adaptive OWPCN: [d] eigenvalues of B              */
#define TDA_UP_QOI_W 19         /* level: [n_qoi][m] Q, the quantity of interest is the linear functional
                                 * qoi = Q @ F(theta) + q0 of the model output       */
#define TDA_UP_QOI_B 20         /* level: [n_qoi] q0                                 */

/* tda_fetch 'field' */
#define TDA_F_THETA 1           /* [nrec][d][n_chains]   engine dtype               */
#define TDA_F_PRIOR 2           /* [nrec][n_chains]      engine dtype               */
#define TDA_F_LIKE 3            /* [nrec][n_chains]      engine dtype               */
#define TDA_F_OUTPUT 4          /* [nrec][m][n_chains]   engine dtype               */
#define TDA_F_ACCEPT 5          /* [nrec][n_chains]      uint8                      */
#define TDA_F_QOI 6             /* [nrec][n_qoi][n_chains] engine dtype: Link.qoi (link.py:38-48), rebuilt from the
                                 * recorded parameters (linear models) or model outputs when it is fetched */
#define TDA_CF_OFFSETS 7        /* tda_compact_fetch only: int64 [n_chains + 1] row offsets per chain */

/* tda_get / tda_set 'what' (float64 on the host side unless noted) */
#define TDA_G_SCALING 1         /* [n_chains]                                       */
#define TDA_G_ACCEPT_COUNTS 2   /* int64 [n_levels][n_chains] accepted local steps  */
#define TDA_G_CURSORS 3         /* int64 [2][n_chains] consumed normals, uniforms   */
#define TDA_G_AM_SIGMA 4        /* [n_chains][d][d]                                 */
#define TDA_G_AM_MU 5           /* [n_chains][d]                                    */
#define TDA_G_THETA 6           /* level: current state [n_chains][d]               */
#define TDA_G_NRECORDS 7        /* int64 [n_levels] records written so far          */
#define TDA_G_MOMENTS 8         /* finest level running sums: [2][d][n_chains] (sum x, sum x^2) */
#define TDA_G_TC16_TIMELINE 10  /* get only, diagnostic: int64 [4][256] clock64 stamps of CTA 0 of kernel 3 (first call arms the probe) */
#define TDA_G_KERNEL 11         /* get only: int64 [1], the kernel tda_engine_run would launch now (ids of tda_select_kernel) */
#define TDA_G_ERROR_FLAGS 12    /* get only: int64 [1], sticky numerical-trouble flags since tda_engine_init: bit 0 = a
                                 * non-positive Cholesky pivot was clamped (error-model covariance / AM covariance);
                                 * the reference's SVD / np.linalg.inv do not fail there, so the chains go on */
#define TDA_G_ZROUND 9          /* set only: one float64 flag; non-zero = Philox normals on the fp16 grid
                                 * ("z16" stream) also for the generic / 3xTF32 kernels (float32 engine) */

typedef struct tda_engine tda_engine;

int tda_abi_version(void);
const char *tda_last_error(void);

int tda_engine_create(const tda_config *cfg, int device, tda_engine **out);
int tda_engine_destroy(tda_engine *e);

int tda_upload(tda_engine *e, int what, int level, const double *host, size_t count);
int tda_engine_init(tda_engine *e, void *cuda_stream);
/* Fails (-1) without launching when a level that stores history would run past its hist_capacity. */
int tda_engine_run(tda_engine *e, int64_t iterations, void *cuda_stream);
/* Same transitions with nothing recorded (burn-in): history position and buffers are untouched. */
int tda_engine_burn(tda_engine *e, int64_t iterations, void *cuda_stream);
int tda_engine_sync(tda_engine *e, void *cuda_stream);

/* Copies history records [rec0, rec0+nrec) of one level into host memory (pinned or
 * pageable), in the device layout documented at TDA_F_*; async on cuda_stream when the host
 * buffer is pinned.  Returns the number of bytes written through *bytes.
 * Records written by the tensor-core Delayed-Acceptance kernel carry the parameters, the
 * log-likelihood and the accept flag; the coarse Links' log-prior (TDA_F_PRIOR) and the model
 * outputs (TDA_F_OUTPUT, Link.model_output link.py:1-48) are rebuilt from the recorded
 * parameters on cuda_stream the first time either field is fetched. */
int tda_fetch(tda_engine *e, int level, int field, int64_t rec0, int64_t nrec,
              void *host_dst, size_t dst_bytes, size_t *bytes, void *cuda_stream);

int tda_get(tda_engine *e, int what, int level, void *host_dst, size_t dst_bytes);
int tda_set(tda_engine *e, int what, int level, const void *host_src, size_t src_bytes);

/* Raw device pointer + byte size of an engine buffer, for zero-copy interop (e.g. wrapping
 * the DREAM archive in a torch tensor for the NCCL all-gather).  buffer ids: */
#define TDA_BUF_DREAM_ARCHIVE 1 /* [capacity][n_chains_global][d] engine dtype      */
#define TDA_BUF_HIST_THETA 2    /* level                                            */
int tda_device_buffer(tda_engine *e, int buffer, int level, void **dev_ptr, size_t *bytes);
/* DREAM shared archive: number of filled slots (M0 + steps) */
int tda_dream_slots(tda_engine *e, int64_t *slots);

/* Fills host arrays z[n_chains][nz], u[n_chains][nu] (float64) with the values the engine's
 * Philox streams deliver in the engine dtype -- what TDA_RNG_PHILOX mode consumes. */
int tda_fill_streams(tda_engine *e, double *z, int64_t nz, double *u, int64_t nu);

/* Compacted history of the FINEST level: a rejected step re-appends the same Link object in the
 * reference (chain.py:116, :434; proposal.py:1601), so a record whose accept flag is 0 equals the record
 * before it.  tda_compact_begin enqueues, on cuda_stream, the device-side compaction of records
 * [rec0, rec0+nrec) into engine-owned buffers of `slot` (0 or 1): the accept flag of every record
 * (first_is_full != 0: record rec0 counts as accepted whatever its flag -- the initial Link) and, chain by
 * chain, the fields (TDA_STORE_THETA | TDA_STORE_STATS | TDA_STORE_OUTPUT, plus TDA_STORE_QOI) of the
 * accepted records only.  tda_compact_rows blocks until the row count of the slot is known.
 * tda_compact_fetch enqueues the device->host copy of one field on the engine's copy stream (pinned host
 * memory makes it asynchronous, so that it overlaps the next tda_engine_run): TDA_F_ACCEPT
 * [nrec][n_chains] uint8, TDA_CF_OFFSETS int64 [n_chains+1] (rows of chain c: offsets[c]..offsets[c+1]),
 * TDA_F_THETA [n_rows][d], TDA_F_PRIOR / TDA_F_LIKE [n_rows], TDA_F_OUTPUT [n_rows][m], TDA_F_QOI
 * [n_rows][n_qoi], engine dtype.  tda_compact_sync waits for the copies.  The next tda_engine_run may
 * overwrite the dense records as soon as tda_compact_begin has been enqueued on the same stream. */
#define TDA_STORE_QOI 16
int tda_compact_begin(tda_engine *e, int level, int64_t rec0, int64_t nrec, int first_is_full, int fields,
                      int slot, void *cuda_stream);
int tda_compact_rows(tda_engine *e, int slot, int64_t *n_rows);
int tda_compact_fetch(tda_engine *e, int slot, int field, void *host_dst, size_t dst_bytes, size_t *bytes);
int tda_compact_sync(tda_engine *e);

/* DREAM with the shared archive (proposal.py:1627-1656, ray.py:366-384) over several GPUs of one node, one
 * process per GPU: tda_peer_export writes an opaque block (cudaIpc handles of this engine's archive replica and
 * step flags; *needed = its size) that the ranks exchange (e.g. torch.distributed all-gather);
 * tda_peer_import maps the other ranks' replicas.  From then on tda_engine_run is ONE persistent launch per
 * call on every rank: each step's new rows are stored straight into every replica over NVLink and a flag
 * handshake in peer memory closes the step (lock-step visibility, as on one GPU) -- every rank must call
 * tda_engine_run with the same iteration count.  Without the import, callers all-gather each step's rows
 * themselves (tda_device_buffer + NCCL). */
int tda_peer_export(tda_engine *e, void *out, size_t bytes, size_t *needed);
int tda_peer_import(tda_engine *e, int n_ranks, int my_rank, const void *handles, size_t bytes);

/* Rank-normalised split R-hat / bulk ESS (Vehtari et al. 2021: what ArviZ computes on the reference's
 * to_inference_data output, diagnostics.py:6-69) of a level's recorded parameters [rec0, rec0+nrec), computed
 * ON THE DEVICE -- the history does not cross PCIe.  For each of the d parameters the chains are split in
 * halves (n_half = nrec / 2), every value is ranked (average ranks for ties) and turned into a normal score, and
 * the plain sums the multi-chain estimators need are returned (so ranks of a multi-GPU job can all-reduce them):
 *   sums[k][0 .. n_lag)      sum over the 2 n_chains split chains of the biased autocovariance at lag t
 *   sums[k][n_lag + 0..3]    sum of split-chain means, sum of their squares, number of split chains, n_half
 *   folded[k][0..3]          the same (lag 0 only) for the normal scores of |x - median|
 * n_lag <= 0 or > n_half means n_half.  sums: float64 [d][n_lag + 4], folded: float64 [d][4]. */
int tda_ess_sums(tda_engine *e, int level, int64_t rec0, int64_t nrec, int n_lag, double *sums, double *folded,
                 void *cuda_stream);

/* Device blocks of destroyed engines are kept in a process-wide pool and reused by the next engine (cudaMalloc /
 * cudaFree of multi-GB history buffers cost more than a short run); this returns them to the driver. */
int tda_pool_trim(void);

/* Page-locked host memory for asynchronous copies (cudaHostAlloc / cudaFreeHost). */
int tda_host_alloc(size_t bytes, void **ptr);
int tda_host_free(void *ptr);

/* Rewinds the history write position of every level to record 0 (the buffers are reused;
 * call after the records have been fetched).  The chains themselves are unaffected. */
int tda_history_reset(tda_engine *e);

/* Kernel selection for tda_engine_run: 0 = automatic, 1 = generic lock-step kernel,
 * 4 = register-resident single-level kernel (one thread per chain; d <= 8, RWMH / pCN / MALA,
 * isotropic or diagonal likelihood, Rosenbrock or a linear model with m <= 256; fails otherwise),
 * 2 = tcgen05 tensor-core Delayed-Acceptance kernel with 3xTF32 operands, 3 = tcgen05
 * Delayed-Acceptance kernel with two-term fp16-split operands and normals produced by dedicated
 * warps, 5 = tcgen05 Delayed-Acceptance kernel with whitened chain state and output recursion
 * (tda_da_tcr.cu; per-chain pCN step sizes), 6 = warp-per-chain DREAM(Z) / DREAM kernel (tda_dream_warp.cu:
 * single level, linear model, d <= 32, non-adaptive crossover), 7 = warp-per-chain MH / DA / MLDA kernel for the
 * 1-D Poisson model with the state-independent error model (tda_mlda_warp.cu) (2, 3, 5, 6 and 7 fail if the
 * configuration is not supported).  Kernels 3 and 5 consume the "z16"
 * Philox normal stream (normals rounded to the fp16 grid at scale 4096); tda_fill_streams
 * exports whatever stream the selected kernel consumes. */
int tda_select_kernel(tda_engine *e, int which);

/* Diagnostic: D[128][N] = A[128][64] @ B[64][N] (row-major float32 host arrays) computed by one
 * CTA on the tcgen05 tensor cores with the conventions of the tensor-core DA kernel
 * (A operand in TMEM when a_in_tmem != 0, else in shared memory; B K-major in shared memory;
 * 3xTF32 split when split != 0, a single TF32 pass on B otherwise).  N: multiple of 8, <= 256. */
int tda_tc_gemm_selftest(const float *A, const float *B, int N, float *D, int a_in_tmem, int split);

/* Same diagnostic for the fp16-split conventions of kernel 3 (kind::f16 MMAs, two-term split of
 * both operands at power-of-two scales, A packed in TMEM or canonical in shared memory).
 * N: multiple of 16, <= 256. */
int tda_tc16_gemm_selftest(const float *A, const float *B, int N, float *D, int a_in_tmem);

/* Checkpoint / resume across processes.  The reference's chain objects are resumable only in
 * memory (calling .sample() again, chain.py:78-129); here the complete sampler state -- constants,
 * the current Links of every level, proposal state (step sizes, adaptation windows, AM moments,
 * DREAM archive), error-model moments and factors, stream cursors -- can be written to a host blob
 * and loaded into an engine created with the same configuration; the run continues bit for bit.
 * The stored history is not part of the blob (fetch it first); the history position restarts at 0. */
int tda_state_size(tda_engine *e, size_t *bytes);
int tda_state_save(tda_engine *e, void *host_dst, size_t dst_bytes);
int tda_state_load(tda_engine *e, const void *host_src, size_t src_bytes);

/* Kernel launches issued by this library since load (for bench.py's gpu_launches). */
int64_t tda_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TINYDA_B200_H */
