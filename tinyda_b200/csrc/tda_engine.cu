// Host side of the tinyda_b200 engine: owns device memory, lays constants out for the kernels,
// launches the lock-step chain kernel, exposes the C ABI declared in include/tinyda_b200.h.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "tda_kernels.cuh"
#include "tda_da_tc.cuh"
#include "tda_da_tc16.cuh"
#include "tda_mh_reg.cuh"
#include "tda_post.h"
#include "tda_da_tcr.h"
#include "tda_dream_warp.h"
#include "tda_mlda_warp.h"
#include <map>
#include <mutex>

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return fail(-2, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    } while (0)

inline int round_up(int x, int q) { return (x + q - 1) / q * q; }

// Device-memory pool.  cudaMalloc / cudaFree of the multi-GB history and compaction buffers cost tens to
// hundreds of milliseconds (and cudaFree synchronises the device), which would dominate a tda.sample() call
// of a few hundred iterations: blocks released by a destroyed engine are kept, keyed by device and size
// class (<= 12.5 % slack), and handed to the next engine.  tda_pool_trim() returns them to the driver; a
// failed cudaMalloc trims and retries.
struct DevPool {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> blocks;
    size_t held = 0;
    size_t cap() const {
        static const size_t c = [] { const char* e = getenv("TDA_DEVICE_POOL_MAX_GB"); return (size_t)(e ? atof(e) : 64.0) << 30; }();
        return c;
    }
    static size_t size_class(size_t bytes) {
        if (bytes < 512) return 512;
        int lg = 63 - __builtin_clzll((unsigned long long)bytes);
        const size_t step = (size_t)1 << (lg > 12 ? lg - 3 : 9);
        return (bytes + step - 1) / step * step;
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto& kv : blocks) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
        cudaSetDevice(cur);
        blocks.clear();
        held = 0;
    }
    cudaError_t alloc(void** out, size_t cls, int device) {
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = blocks.find({device, cls});
            if (it != blocks.end()) {
                *out = it->second;
                blocks.erase(it);
                held -= cls;
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(out, cls);
        if (e != cudaSuccess) {
            cudaGetLastError();
            trim();
            e = cudaMalloc(out, cls);
        }
        return e;
    }
    void release(void* ptr, size_t cls, int device) {
        std::lock_guard<std::mutex> lk(mu);
        if (held + cls <= cap()) {
            blocks.insert({{device, cls}, ptr});
            held += cls;
        } else {
            cudaFree(ptr);
        }
    }
};
DevPool g_pool;

// small page-locked scalars, copy streams and events are recycled too (cudaHostAlloc / cudaFreeHost /
// cudaStreamCreate cost a millisecond or more each and synchronise)
struct SmallPool {
    std::mutex mu;
    std::vector<long long*> pinned;
    std::vector<std::pair<int, cudaStream_t>> streams;
    std::vector<std::pair<int, cudaEvent_t>> events;
    long long* get_pinned() {
        std::lock_guard<std::mutex> lk(mu);
        if (!pinned.empty()) { long long* p = pinned.back(); pinned.pop_back(); return p; }
        long long* p = nullptr;
        if (cudaHostAlloc((void**)&p, 64, cudaHostAllocPortable) != cudaSuccess) return nullptr;
        return p;
    }
    void put_pinned(long long* p) { std::lock_guard<std::mutex> lk(mu); pinned.push_back(p); }
    cudaStream_t get_stream(int dev) {
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < streams.size(); i++)
            if (streams[i].first == dev) { cudaStream_t s = streams[i].second; streams.erase(streams.begin() + i); return s; }
        cudaStream_t s = nullptr;
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        return s;
    }
    void put_stream(int dev, cudaStream_t s) { std::lock_guard<std::mutex> lk(mu); streams.push_back({dev, s}); }
    cudaEvent_t get_event(int dev) {
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < events.size(); i++)
            if (events[i].first == dev) { cudaEvent_t e = events[i].second; events.erase(events.begin() + i); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        return e;
    }
    void put_event(int dev, cudaEvent_t e) { std::lock_guard<std::mutex> lk(mu); events.push_back({dev, e}); }
};
SmallPool g_small;

}  // namespace

struct tda_engine {
    virtual ~tda_engine() {}
    virtual int upload(int what, int level, const double* host, size_t count) = 0;
    virtual int init(cudaStream_t st) = 0;
    virtual int run(long long iterations, cudaStream_t st, bool record = true) = 0;
    virtual int compact_begin(int level, long long rec0, long long nrec, int first_is_full, int fields, int slot, cudaStream_t st) = 0;
    virtual int compact_rows(int slot, long long* n_rows) = 0;
    virtual int compact_fetch(int slot, int field, void* dst, size_t dst_bytes, size_t* bytes) = 0;
    virtual int compact_sync() = 0;
    virtual int peer_export(void* out, size_t bytes, size_t* needed) = 0;
    virtual int peer_import(int n_ranks, int my_rank, const void* handles, size_t bytes) = 0;
    virtual int ess_sums(int level, long long rec0, long long nrec, int n_lag, double* sums, double* folded, cudaStream_t st) = 0;
    virtual int fetch(int level, int field, long long rec0, long long nrec, void* dst, size_t dst_bytes,
                      size_t* bytes, cudaStream_t st) = 0;
    virtual int get(int what, int level, void* dst, size_t bytes) = 0;
    virtual int set(int what, int level, const void* src, size_t bytes) = 0;
    virtual int device_buffer(int buffer, int level, void** ptr, size_t* bytes) = 0;
    virtual int fill_streams(double* z, long long nz, double* u, long long nu) = 0;
    virtual int history_reset() = 0;
    virtual int select_kernel(int which) = 0;
    virtual int state_io(void* host, size_t bytes, size_t* needed, int load) = 0;
    tda_config cfg;
    int device = 0;
    long long dream_slots = 0;
};

namespace {

template <typename R>
__global__ void fill_kernel(R* dst, R value, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = value;
}

template <typename R>
__global__ void broadcast_kernel(R* dst, const R* src, int n, int ld_src, int cols, int Cs) {
    // dst[(i*cols + j)][c] = src[i*ld_src + j]
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n * cols * Cs;
    if (e < total) {
        size_t ij = e / Cs;
        int i = (int)(ij / cols), j = (int)(ij - (size_t)i * cols);
        dst[e] = src[(size_t)i * ld_src + j];
    }
}

// initial states: host layout [C][d] float64 -> device SoA [d][Cs] of the engine dtype, for every
// level at once (shared-memory tile transpose: coalesced on both sides)
template <typename R>
__global__ void init_theta_kernel(const double* __restrict__ src, int C, int d, int Cs, R* t0, R* t1, R* t2, R* t3) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, k = k0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && k < d) ? src[(size_t)c * d + k] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, c = c0 + threadIdx.x;
        if (k < d && c < Cs) {
            const R v = (R)tile[threadIdx.x][i];
            const size_t o = (size_t)k * Cs + c;
            t0[o] = v;
            if (t1) t1[o] = v;
            if (t2) t2[o] = v;
            if (t3) t3[o] = v;
        }
    }
}

// Derived Link fields of history records written by the tensor-core DA kernel, which stores the
// parameters, the log-likelihood and the accept flag of a record only:
//   mode 0  Link.model_output  F[r][n][c] = b[n] + sum_k theta[r][k][c] * A[k][n]   (posterior.py:95-105, LinearModel)
//   mode 1  Link.prior         prior[r][c] = -0.5 * (logconst + |(theta[r][:, c] - mu) @ LP|^2)   (posterior.py:92)
// One thread = one (record, chain): the parameter row sits in registers, the operator comes through
// the read-only cache at warp-uniform addresses, stores are coalesced (lane = chain).  Runs when the
// fields are first fetched, i.e. next to a PCIe copy of the same records, never inside the sampling loop.
template <typename R, int D>
__global__ void __launch_bounds__(256) hist_fill_kernel(const R* theta, const R* __restrict__ W, int ldW, int ncol,
                                                        const R* __restrict__ off, const R* __restrict__ mu, R* out,
                                                        int d, int Cs, int mode, R logconst) {
    const int cblocks = Cs / 256;
    const size_t r = blockIdx.x / cblocks;
    const int c = (int)(blockIdx.x - r * cblocks) * 256 + threadIdx.x;
    const R* th_src = theta + r * (size_t)d * Cs + c;
    R th[D];                                   // d <= D; the missing rows are zero and skipped below
#pragma unroll
    for (int k = 0; k < D; k++) th[k] = (k < d) ? th_src[(size_t)k * Cs] - (mu ? mu[k] : (R)0) : (R)0;
    R ssq = 0;
    R* dst = out + (mode == 0 ? r * (size_t)ncol * Cs : r * (size_t)Cs) + c;
    for (int n = 0; n < ncol; n += 4) {
        R a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (n + 3 < ncol) {
#pragma unroll
            for (int k = 0; k < D; k++) {
                if (k >= d) break;
                const R* w = W + (size_t)k * ldW + n;
                a0 = fma(th[k], __ldg(w), a0); a1 = fma(th[k], __ldg(w + 1), a1);
                a2 = fma(th[k], __ldg(w + 2), a2); a3 = fma(th[k], __ldg(w + 3), a3);
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; k++) {
                if (k >= d) break;
                const R* w = W + (size_t)k * ldW + n;
                a0 = fma(th[k], __ldg(w), a0);
                if (n + 1 < ncol) a1 = fma(th[k], __ldg(w + 1), a1);
                if (n + 2 < ncol) a2 = fma(th[k], __ldg(w + 2), a2);
            }
        }
        if (mode == 0) {
            dst[(size_t)n * Cs] = a0 + (off ? off[n] : (R)0);
            if (n + 1 < ncol) dst[(size_t)(n + 1) * Cs] = a1 + (off ? off[n + 1] : (R)0);
            if (n + 2 < ncol) dst[(size_t)(n + 2) * Cs] = a2 + (off ? off[n + 2] : (R)0);
            if (n + 3 < ncol) dst[(size_t)(n + 3) * Cs] = a3 + (off ? off[n + 3] : (R)0);
        } else {
            ssq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
    }
    if (mode == 1) *dst = (R)-0.5 * (logconst + ssq);
}

// the tcr kernel lives in its own translation unit and is float only
template <typename R>
struct TcrAdapter {
    std::string err;
    bool eligible(const tda_config&, const tda::Params<R>&) const { return false; }
    int ready(const tda::Params<R>&, const tda_config&, cudaStream_t) { return 1; }
    int run(tda::Params<R>&, const tda_config&, long long, int, cudaStream_t) { return -5; }
    void invalidate() {}
    void destroy() {}
};
template <>
struct TcrAdapter<float> {
    tda::DaTcrState s;
    std::string err;
    bool eligible(const tda_config& c, const tda::Params<float>& P) const {
        static const bool off = getenv("TDA_DISABLE_TCR") != nullptr;
        return !off && s.eligible(c, P);
    }
    int ready(const tda::Params<float>& P, const tda_config& c, cudaStream_t st) { int r = s.ready(P, c, st); err = s.err; return r; }
    int run(tda::Params<float>& P, const tda_config& c, long long it, int sms, cudaStream_t st) { int r = s.run(P, c, it, sms, st); err = s.err; return r; }
    void invalidate() { s.invalidate(); }
    void destroy() { s.destroy(); }
};

// Link.qoi of models that are not linear in theta: qoi[r][j][c] = q0[j] + sum_n Q[j][n] * F[r][n][c]
// (thread = one (record, chain); lane = chain -> coalesced)
template <typename R>
__global__ void __launch_bounds__(256) qoi_from_F_kernel(const R* __restrict__ F, const R* __restrict__ Q, const R* __restrict__ q0,
                                                         R* __restrict__ out, int m, int nq, int Cs) {
    const int cblocks = Cs / 256;
    const size_t r = blockIdx.x / cblocks;
    const int c = (int)(blockIdx.x - r * cblocks) * 256 + threadIdx.x;
    const R* f = F + r * (size_t)m * Cs + c;
    for (int j = 0; j < nq; j++) {
        R a = q0 ? q0[j] : (R)0;
        for (int n = 0; n < m; n++) a = fma(__ldg(Q + (size_t)j * m + n), f[(size_t)n * Cs], a);
        out[(r * nq + j) * (size_t)Cs + c] = a;
    }
}

template <typename R>
struct EngineT : tda_engine {
    tda::Params<R> P;
    std::vector<void*> allocs;
    std::vector<size_t> alloc_bytes;
    std::vector<size_t> alloc_cls;      // size class the block came from (device pool)
    std::vector<char> alloc_is_state;   // 0: history buffer (not part of a checkpoint)
    bool alloc_history = false;
    int Cs = 0, n_tiles = 0, kt = 0, sm_count = 148;
    size_t smem_bytes = 0;
    bool initialised = false;
    int kernel_choice = 0;     // 0 auto, 1 generic, 2 tensor-core DA (3xTF32), 3 tensor-core DA (fp16 split),
                               // 4 register-resident single-level MH, 5 tensor-core DA (whitened state),
                               // 6 warp-per-chain DREAM(Z) / DREAM
    void* mw_sig = nullptr;    // warp-major images of the error-model matrices (kernel 7), allocated on first use
    void* mw_li = nullptr;
    void* mw_phi = nullptr;
    size_t mw_cls = 0, mw_phi_cls = 0;
    int dreamw_occ = 2;        // its variant: resident CTAs per SM
    int dreamw_grid = -1;      // CTAs of the warp-per-chain DREAM kernel (-1: not asked yet, 0: the job does not fit)
    int z_round_user = 0;      // generic / TF32 kernels: use the z16 normal stream (tda_set TDA_G_ZROUND)
    std::vector<int> ldA;
    double* stage_theta = nullptr;   // device staging for the initial states ([C][d] float64)
    tda::DaTcState<R> tc;      // tcgen05 fast path (float only; inert for double)
    tda::DaTc16State<R> tc16;  // fp16-split tcgen05 fast path (float only)
    bool tc16_unfit = false;   // prepare() found operands (or a chain state) that do not fit the fp16 range
    bool tc16_state_checked = false;   // the current states were checked against the operand range since they last changed hands
    TcrAdapter<R> tcr;         // whitened-state / output-recursion tcgen05 kernel (float only), tda_da_tcr.cu
    bool tcr_unfit = false;
    // records written by the fp16-split kernel whose derived fields (coarse Link.prior, Link.model_output)
    // have not been filled yet, per level; and: the levels' current model outputs lag behind theta
    long long lazy_lo[tda::MAXL] = {0, 0, 0, 0}, lazy_hi[tda::MAXL] = {0, 0, 0, 0};
    // coarse records written by the tcr kernel hold WHITENED parameters until they are first needed
    long long lazy_w_lo = 0, lazy_w_hi = 0;
    bool state_F_stale = false;
    bool burning = false;      // inside tda_engine_burn: nothing is recorded
    unsigned int* dream_flags = nullptr;   // [8] step flags written by the peers (shared-archive DREAM over several GPUs)
    unsigned int flag_next = 0;            // flag value of the next persistent launch's first step, minus one
    unsigned int arr_next = 0;             // lock-step steps the warp-per-chain kernel has run with arrival counters
    std::vector<void*> peer_mapped;        // cudaIpcOpenMemHandle results
    // quantities of interest: qoi = Q @ F(theta) + q0 per level.  Linear models: composed with the operator
    // on upload (qoi_W = G^T Q^T [d][ldq], qoi_b = Q b + q0) so that Link.qoi comes straight from theta
    int nq[tda::MAXL] = {0, 0, 0, 0}, ldq[tda::MAXL] = {0, 0, 0, 0};
    R* qoi_Q[tda::MAXL] = {nullptr, nullptr, nullptr, nullptr};    // [nq][m]
    R* qoi_q0[tda::MAXL] = {nullptr, nullptr, nullptr, nullptr};   // [nq]
    R* qoi_W[tda::MAXL] = {nullptr, nullptr, nullptr, nullptr};    // linear: [d][ldq]
    R* qoi_b[tda::MAXL] = {nullptr, nullptr, nullptr, nullptr};    // linear: [nq]
    // compaction of the finest level's records to the accepted ones (tda_post.h), double-buffered
    struct CompactSlot {
        long long nrec = 0, n_rows = -1;
        int fields = 0;
        long long* offsets = nullptr;     // [C + 1]
        long long* scratch = nullptr;
        uint8_t* flags = nullptr;         // [nrec][Cs]
        size_t flags_cap = 0, offsets_cap = 0, scratch_cap = 0;
        void* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // theta, prior, like, output, qoi rows
        size_t cap[5] = {0, 0, 0, 0, 0};
        long long* total_pinned = nullptr;
        cudaEvent_t ready = nullptr, copied = nullptr;
        bool pending = false, has_copies = false;
    } cslot[2];
    cudaStream_t copy_stream = nullptr;

    ~EngineT() override {
        cudaSetDevice(device);
        cudaDeviceSynchronize();           // nothing of this engine is in flight when its blocks go back to the pool
        for (void* m : peer_mapped) cudaIpcCloseMemHandle(m);
        for (size_t i = 0; i < allocs.size(); i++) g_pool.release(allocs[i], alloc_cls[i], device);
        for (auto& s : cslot) {
            if (s.offsets) g_pool.release(s.offsets, s.offsets_cap, device);
            if (s.scratch) g_pool.release(s.scratch, s.scratch_cap, device);
            if (s.flags) g_pool.release(s.flags, s.flags_cap, device);
            for (int f = 0; f < 5; f++) if (s.buf[f]) g_pool.release(s.buf[f], s.cap[f], device);
            if (s.total_pinned) g_small.put_pinned(s.total_pinned);
            if (s.ready) g_small.put_event(device, s.ready);
            if (s.copied) g_small.put_event(device, s.copied);
        }
        if (copy_stream) g_small.put_stream(device, copy_stream);
        if (mw_sig) g_pool.release(mw_sig, mw_cls, device);
        if (mw_li) g_pool.release(mw_li, mw_cls, device);
        if (mw_phi) g_pool.release(mw_phi, mw_phi_cls, device);
        tc.destroy();
        tc16.destroy();
        tcr.destroy();
    }

    template <typename T>
    int dalloc(T** out, size_t n) {
        void* p = nullptr;
        size_t bytes = (n ? n : 1) * sizeof(T);
        const size_t cls = DevPool::size_class(bytes);
        cudaError_t e = g_pool.alloc(&p, cls, device);
        if (e != cudaSuccess) return fail(-3, std::string("cudaMalloc ") + std::to_string(bytes) + " bytes: " + cudaGetErrorString(e));
        // history buffers are written before they are read (the record counters say how far): no clearing
        if (!alloc_history) e = cudaMemsetAsync(p, 0, bytes, 0);
        if (e != cudaSuccess) return fail(-3, std::string("cudaMemset: ") + cudaGetErrorString(e));
        allocs.push_back(p);
        alloc_bytes.push_back(bytes);
        alloc_cls.push_back(cls);
        alloc_is_state.push_back(alloc_history ? 0 : 1);
        *out = reinterpret_cast<T*>(p);
        return 0;
    }
#define DALLOC(ptr, n)                         \
    do {                                       \
        int _r = dalloc(&(ptr), (size_t)(n));  \
        if (_r) return _r;                     \
    } while (0)

    int create() {
        const tda_config& c = cfg;
        CUDA_TRY(cudaSetDevice(device));
        // one attribute, not cudaGetDeviceProperties: that call also reads clocks and bus state through the driver and
        // was measured at 3-70 ms, erratically, inside every tda.sample() call
        CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
        memset(&P, 0, sizeof(P));
        const int L = c.n_levels, d = c.d;
        P.L = L; P.d = d; P.aem = c.aem; P.rng_mode = c.rng_mode; P.prop_kind = c.prop_kind;
        P.adaptive = c.adaptive; P.period = c.period > 0 ? c.period : 1; P.am_t0 = c.am_t0;
        P.am_device_refactor = c.am_device_refactor;
        P.randomize = c.randomize_subchain;
        P.mtm_k = c.mtm_k;
        P.mtm_include_current = c.mtm_include_current;
        for (int l = 0; l < TDA_MAX_LEVELS; l++) P.J[l] = c.subchain[l];
        P.C = (int)c.n_chains;
        // padded to whole pairs of 128-chain tiles (the tensor-core kernel processes tile pairs)
        Cs = (P.C + 2 * tda::TC - 1) / (2 * tda::TC) * (2 * tda::TC);
        n_tiles = Cs / tda::TC;
        P.Cs = Cs; P.n_tiles = n_tiles;
        P.chain_offset = c.chain_offset;
        P.Cg = (c.prop_kind == TDA_PROP_DREAM && c.n_chains_global > 0) ? c.n_chains_global : c.n_chains;
        P.arch_off = (c.prop_kind == TDA_PROP_DREAM) ? c.chain_offset : 0;
        P.seed = c.seed;
        P.gamma = (R)c.gamma; P.alpha_star = (R)c.alpha_star; P.am_sd = (R)c.am_sd; P.am_eps = (R)c.am_eps;
        P.dream_b = (R)c.dream_b; P.dream_b_star = (R)c.dream_b_star;
        P.dream_M0 = c.dream_M0; P.dream_delta = c.dream_delta; P.dream_nCR = c.dream_nCR;
        P.dream_sync = (c.prop_kind == TDA_PROP_DREAM && c.dream_sync_every > 1) ? c.dream_sync_every : 1;
        P.dream_cap = c.dream_capacity;
        P.prior_logconst = (R)c.prior_logconst;
        P.scaling0 = (R)c.scaling;
        P.ldD = round_up(d, tda::NB);
        P.zlen = c.stream_z_len; P.ulen = c.stream_u_len;

        R* tmp;
        DALLOC(tmp, d); P.prior_mean = tmp;
        DALLOC(tmp, (size_t)d * P.ldD); P.LP = tmp;
        DALLOC(tmp, (size_t)d * P.ldD); P.Pprec = tmp;
        DALLOC(tmp, (size_t)d * P.ldD); P.T = tmp;
        if (c.prop_kind == TDA_PROP_INDEP) {
            DALLOC(tmp, (size_t)d * P.ldD); P.Sop = tmp;
            DALLOC(tmp, d); P.ow_lambda = tmp;
            DALLOC(P.qcur, Cs);
        }
        if (c.prop_kind == TDA_PROP_OWPCN) {
            DALLOC(tmp, (size_t)d * P.ldD); P.Sop = tmp;
            DALLOC(tmp, (size_t)d * P.ldD); P.Sop2 = tmp;
            DALLOC(tmp, d); P.ow_lambda = tmp;
        }
        DALLOC(P.scaling, Cs);
        DALLOC(P.ucur, Cs);
        DALLOC(P.sum1, (size_t)d * Cs);
        DALLOC(P.sum2, (size_t)d * Cs);
        DALLOC(P.error_flag, 1);
        if (c.adaptive) { DALLOC(P.win, (size_t)P.period * Cs); DALLOC(P.win_sum, Cs); }
        if (c.rng_mode == TDA_RNG_INJECTED) {
            DALLOC(tmp, (size_t)P.C * P.zlen); P.zs = tmp;
            DALLOC(tmp, (size_t)P.C * P.ulen); P.us = tmp;
        }
        if (c.prop_kind == TDA_PROP_AM) {
            DALLOC(P.am_mu, (size_t)d * Cs);
            DALLOC(P.am_sigma, (size_t)d * d * Cs);
            DALLOC(P.am_T, (size_t)d * d * Cs);
        }
        if (c.prop_kind == TDA_PROP_MALA) { DALLOC(P.grad, (size_t)d * Cs); DALLOC(P.gradp, (size_t)d * Cs); }
        DALLOC(P.grid_bar, 1);
        if (tda::is_dream(c.prop_kind)) {
            DALLOC(P.archive, (size_t)c.dream_capacity * P.Cg * d);
            DALLOC(dream_flags, 16);        // [0..8): step flags per writer rank, [8..16): arrival counters per writer rank
            dream_slots = c.dream_M0;
            if (c.adaptive) {
                DALLOC(P.dream_pCR, (size_t)tda::MAX_NCR * Cs);
                DALLOC(P.dream_DeltaCR, (size_t)tda::MAX_NCR * Cs);
                DALLOC(P.dream_LCR, (size_t)tda::MAX_NCR * Cs);
                DALLOC(P.dream_mCR, Cs);
                DALLOC(P.arch_s1, (size_t)d * Cs);
                DALLOC(P.arch_s2, (size_t)d * Cs);
            }
        }
        kt = d;
        int n_max = 0;
        ldA.assign(L, 0);
        for (int l = 0; l < L; l++) {
            const tda_level_config& lc = c.level[l];
            tda::LevelP<R>& v = P.lv[l];
            const int m = lc.m;
            v.model_kind = lc.model_kind; v.m = m; v.n_grid = lc.n_grid; v.lik_kind = lc.lik_kind;
            v.store = lc.store; v.hist_cap = lc.hist_capacity;
            if (lc.n_qoi < 0 || lc.n_qoi > 64) return fail(-1, "n_qoi out of range (0..64)");
            nq[l] = lc.n_qoi;
            if (nq[l]) {
                ldq[l] = round_up(nq[l], 4);
                DALLOC(qoi_Q[l], (size_t)nq[l] * m); DALLOC(qoi_q0[l], nq[l]);
                if (lc.model_kind == TDA_MODEL_LINEAR) { DALLOC(qoi_W[l], (size_t)d * ldq[l]); DALLOC(qoi_b[l], nq[l]); }
            }
            v.lik_var = (R)lc.lik_var; v.sc0 = (R)lc.model_scalars[0]; v.sc1 = (R)lc.model_scalars[1];
            v.stride = (int)lc.model_scalars[0];
            v.need_F = (lc.lik_kind >= TDA_LIK_DENSE) || (lc.model_kind != TDA_MODEL_LINEAR) || c.aem ||
                       (lc.store & TDA_STORE_OUTPUT) || (c.prop_kind == TDA_PROP_MALA);
            const int ncols = (lc.model_kind == TDA_MODEL_POISSON1D) ? lc.n_grid : m;
            v.ldA = round_up(ncols, tda::NB);
            ldA[l] = v.ldA;
            if (lc.model_kind == TDA_MODEL_POISSON1D && lc.n_grid > n_max) n_max = lc.n_grid;
            if (lc.model_kind != TDA_MODEL_ROSENBROCK) { DALLOC(tmp, (size_t)d * v.ldA); v.A = tmp; }
            if (lc.model_kind == TDA_MODEL_LINEAR && c.prop_kind == TDA_PROP_MALA) {
                DALLOC(tmp, (size_t)m * P.ldD); v.A2 = tmp;
                if (m > kt) kt = m;
            }
            DALLOC(tmp, m); v.b = tmp;
            DALLOC(tmp, m); v.data = tmp;
            DALLOC(tmp, m); v.var = tmp;
            if (lc.lik_kind >= TDA_LIK_DENSE) { DALLOC(tmp, (size_t)m * m); v.prec = tmp; }
            if (lc.lik_kind == TDA_LIK_ADAPTIVE) { DALLOC(tmp, (size_t)m * m); v.cov = tmp; }
            DALLOC(v.theta, (size_t)d * Cs);
            DALLOC(v.prior, Cs);
            DALLOC(v.like, Cs);
            DALLOC(v.sid, Cs);
            DALLOC(v.n_acc, Cs);
            DALLOC(v.acc_sub, Cs);
            if (v.need_F) { DALLOC(v.F, (size_t)m * Cs); DALLOC(v.Fp, (size_t)m * Cs); }
            for (int a = l + 1; a < L; a++) {
                DALLOC(v.sv_prior[a], Cs);
                DALLOC(v.sv_like[a], Cs);
                if (v.need_F) DALLOC(v.sv_F[a], (size_t)m * Cs);
            }
            if (lc.lik_kind == TDA_LIK_ADAPTIVE) {
                DALLOC(v.lik_bias, (size_t)m * Cs);
                DALLOC(v.lik_prec, (size_t)m * m * Cs);
            }
            if (c.aem && l >= 1) {
                DALLOC(v.bias_mu, (size_t)m * Cs);
                DALLOC(v.model_diff, (size_t)m * Cs);
                DALLOC(v.bias_sigma, (size_t)m * m * Cs);
            }
            const size_t cap = (size_t)lc.hist_capacity;
            alloc_history = true;
            if (lc.store & TDA_STORE_THETA) DALLOC(v.h_theta, cap * d * Cs);
            if (lc.store & TDA_STORE_STATS) { DALLOC(v.h_prior, cap * Cs); DALLOC(v.h_like, cap * Cs); }
            if (lc.store & TDA_STORE_OUTPUT) DALLOC(v.h_F, cap * m * Cs);
            if (lc.store & TDA_STORE_ACCEPT) DALLOC(v.h_acc, cap * Cs);
            alloc_history = false;
        }
        if (c.mtm_k) {
            const size_t K = (size_t)c.mtm_k;
            DALLOC(P.zcur, Cs);
            DALLOC(P.mt_theta, K * d * Cs);
            DALLOC(P.mt_prior, K * Cs);
            DALLOC(P.mt_like, K * Cs);
            DALLOC(P.mt_w, (K + 1) * Cs);
            if (P.lv[0].need_F) DALLOC(P.mt_F, K * P.lv[0].m * Cs);
            DALLOC(P.mt_y, (size_t)d * Cs);
            DALLOC(P.mt_sel, Cs);
            DALLOC(P.mt_lse, Cs);
        }
        if (c.randomize_subchain) {
            DALLOC(P.promo_j, Cs);
            DALLOC(P.pm_theta, (size_t)d * Cs);
            DALLOC(P.pm_prior, Cs);
            DALLOC(P.pm_like, Cs);
            DALLOC(P.pm_sid, Cs);
            if (P.lv[0].need_F) DALLOC(P.pm_F, (size_t)P.lv[0].m * Cs);
        }
        if (kt > tda::MAXD) return fail(-1, "contraction dimension exceeds TDA_MAX_D");
        if (n_max > 0) { P.n_max = n_max; DALLOC(P.scratch, (size_t)3 * n_max * Cs); }
        smem_bytes = ((size_t)2 * kt * tda::TC + (size_t)2 * kt * tda::NB + tda::CW * tda::TC + 4 * tda::TC) * sizeof(R) +
                     2 * tda::TC * sizeof(int);
        for (int l = 0; l < L; l++)
            if (c.level[l].lik_kind == TDA_LIK_ADAPTIVE && c.level[l].m > P.m_adapt) P.m_adapt = c.level[l].m;
        smem_bytes += (size_t)P.m_adapt * tda::TC * sizeof(R);
        if (c.aem) {
            // workspace of the cooperative factorisation: (m*m + m) values per chain of a group; the
            // largest power-of-two group that stays within 64 KB
            const size_t m = (size_t)c.level[0].m;
            int G = 16;
            while (G > 1 && (m * m + m) * G * sizeof(R) > 64 * 1024) G /= 2;
            if ((m * m + m) * G * sizeof(R) > 96 * 1024) return fail(-1, "adaptive error model supports up to 128 outputs (float) / 96 (double)");
            P.aem_G = G;
            smem_bytes += (m * m + m) * G * sizeof(R);
        }
        CUDA_TRY(cudaFuncSetAttribute(tda::chain_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        // initial scaling
        {
            size_t n = Cs;
            fill_kernel<R><<<(unsigned)((n + 255) / 256), 256>>>(P.scaling, (R)c.scaling, n);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaDeviceSynchronize());
        return 0;
    }

    // ---- uploads ---------------------------------------------------------------------------
    int put(const R* dst, const std::vector<R>& h) {
        CUDA_TRY(cudaMemcpy(const_cast<R*>(dst), h.data(), h.size() * sizeof(R), cudaMemcpyHostToDevice));
        // a copy from pageable memory may return before its DMA has landed: other streams (the operand fetch of the
        // tensor-core kernels' prepare step) must see the data
        CUDA_TRY(cudaStreamSynchronize(0));
        return 0;
    }

    int put_matrix(const R* dst, const double* src, int rows, int cols, int ld) {
        std::vector<R> h((size_t)rows * ld, (R)0);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) h[(size_t)i * ld + j] = (R)src[(size_t)i * cols + j];
        return put(dst, h);
    }

    int upload(int what, int level, const double* host, size_t count) override {
        CUDA_TRY(cudaSetDevice(device));
        const int d = P.d, L = P.L;
        if (what >= TDA_UP_MODEL_A && what <= TDA_UP_LIK_COV && (level < 0 || level >= L))
            return fail(-1, "upload: bad level");
        auto need = [&](size_t n) -> int {
            if (count != n) return fail(-1, "upload: expected " + std::to_string(n) + " values, got " + std::to_string(count));
            return 0;
        };
        int r;
        switch (what) {
        case TDA_UP_PRIOR_MEAN:
            if ((r = need(d))) return r;
            return put_matrix(P.prior_mean, host, 1, d, d);
        case TDA_UP_PRIOR_LP:
            if ((r = need((size_t)d * d))) return r;
            return put_matrix(P.LP, host, d, d, P.ldD);
        case TDA_UP_PRIOR_PREC:
            if ((r = need((size_t)d * d))) return r;
            return put_matrix(P.Pprec, host, d, d, P.ldD);
        case TDA_UP_PROP_T: {
            if ((r = need((size_t)d * d))) return r;
            if ((r = put_matrix(P.T, host, d, d, P.ldD))) return r;
            if (P.prop_kind == TDA_PROP_AM) {
                size_t total = (size_t)d * d * Cs;
                broadcast_kernel<R><<<(unsigned)((total + 255) / 256), 256>>>(P.am_T, P.T, d, P.ldD, d, Cs);
                g_launches++;
                CUDA_TRY(cudaGetLastError());
                CUDA_TRY(cudaDeviceSynchronize());
            }
            return 0;
        }
        case TDA_UP_PROP_S:
            if (!P.Sop) return fail(-1, "upload: proposal has no state operator");
            if ((r = need((size_t)d * d))) return r;
            return put_matrix(P.Sop, host, d, d, P.ldD);
        case TDA_UP_PROP_S2:
            if (!P.Sop2) return fail(-1, "upload: proposal has no state operator");
            if ((r = need((size_t)d * d))) return r;
            return put_matrix(P.Sop2, host, d, d, P.ldD);
        case TDA_UP_PROP_LAMBDA:
            if (!P.ow_lambda) return fail(-1, "upload: proposal has no operator spectrum");
            if ((r = need((size_t)d))) return r;
            return put_matrix(P.ow_lambda, host, 1, d, d);
        case TDA_UP_MODEL_A: {
            const tda::LevelP<R>& v = P.lv[level];
            const int ncols = (v.model_kind == TDA_MODEL_POISSON1D) ? v.n_grid : v.m;
            if ((r = need((size_t)d * ncols))) return r;
            if ((r = put_matrix(v.A, host, d, ncols, v.ldA))) return r;
            if (v.A2) {   // G [m][ldD] = transpose of the uploaded G^T
                std::vector<R> h((size_t)v.m * P.ldD, (R)0);
                for (int k = 0; k < d; k++)
                    for (int j = 0; j < v.m; j++) h[(size_t)j * P.ldD + k] = (R)host[(size_t)k * v.m + j];
                return put(v.A2, h);
            }
            return 0;
        }
        case TDA_UP_MODEL_B:
            if ((r = need(P.lv[level].m))) return r;
            return put_matrix(P.lv[level].b, host, 1, P.lv[level].m, P.lv[level].m);
        case TDA_UP_LIK_DATA:
            if ((r = need(P.lv[level].m))) return r;
            return put_matrix(P.lv[level].data, host, 1, P.lv[level].m, P.lv[level].m);
        case TDA_UP_LIK_VAR:
            if ((r = need(P.lv[level].m))) return r;
            return put_matrix(P.lv[level].var, host, 1, P.lv[level].m, P.lv[level].m);
        case TDA_UP_LIK_PREC: {
            const int m = P.lv[level].m;
            if (!P.lv[level].prec) return fail(-1, "upload: level has no dense precision");
            if ((r = need((size_t)m * m))) return r;
            return put_matrix(P.lv[level].prec, host, m, m, m);
        }
        case TDA_UP_LIK_COV: {
            const int m = P.lv[level].m;
            if (!P.lv[level].cov) return fail(-1, "upload: level is not adaptive");
            if ((r = need((size_t)m * m))) return r;
            return put_matrix(P.lv[level].cov, host, m, m, m);
        }
        case TDA_UP_QOI_W: {
            if (level < 0 || level >= L || !nq[level]) return fail(-1, "upload: level has no quantity of interest");
            const tda::LevelP<R>& v = P.lv[level];
            if ((r = need((size_t)nq[level] * v.m))) return r;
            if ((r = put_matrix(qoi_Q[level], host, nq[level], v.m, v.m))) return r;
            if (qoi_W[level]) {            // compose with the linear operator: W[k][j] = sum_n G^T[k][n] Q[j][n]
                std::vector<R> A((size_t)d * v.ldA);
                CUDA_TRY(cudaMemcpy(A.data(), v.A, A.size() * sizeof(R), cudaMemcpyDeviceToHost));
                std::vector<R> W((size_t)d * ldq[level], (R)0);
                for (int k = 0; k < d; k++)
                    for (int j = 0; j < nq[level]; j++) {
                        double a = 0;
                        for (int n = 0; n < v.m; n++) a += (double)A[(size_t)k * v.ldA + n] * host[(size_t)j * v.m + n];
                        W[(size_t)k * ldq[level] + j] = (R)a;
                    }
                if ((r = put(qoi_W[level], W))) return r;
                return compose_qoi_offset(level);
            }
            return 0;
        }
        case TDA_UP_QOI_B: {
            if (level < 0 || level >= L || !nq[level]) return fail(-1, "upload: level has no quantity of interest");
            if ((r = need((size_t)nq[level]))) return r;
            if ((r = put_matrix(qoi_q0[level], host, 1, nq[level], nq[level]))) return r;
            return qoi_W[level] ? compose_qoi_offset(level) : 0;
        }
        case TDA_UP_INIT_THETA: {
            if ((r = need((size_t)P.C * d))) return r;
            // one H2D copy of the caller's [C][d] float64 array (asynchronous when it is pinned),
            // transposed / converted on the device
            if (!stage_theta) DALLOC(stage_theta, (size_t)P.C * d);
            CUDA_TRY(cudaMemcpyAsync(stage_theta, host, (size_t)P.C * d * sizeof(double), cudaMemcpyHostToDevice, 0));
            R* t[4] = {nullptr, nullptr, nullptr, nullptr};
            for (int l = 0; l < L && l < 4; l++) t[l] = P.lv[l].theta;
            dim3 grid((unsigned)((Cs + 31) / 32), (unsigned)((d + 31) / 32)), block(32, 8);
            init_theta_kernel<R><<<grid, block>>>(stage_theta, P.C, d, Cs, t[0], t[1], t[2], t[3]);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaStreamSynchronize(0));
            return 0;
        }
        case TDA_UP_STREAM_Z:
            if (!P.zs) return fail(-1, "upload: engine is not in injected-stream mode");
            if ((r = need((size_t)P.C * P.zlen))) return r;
            return put_matrix(P.zs, host, 1, (int)count, (int)count);
        case TDA_UP_STREAM_U:
            if (!P.us) return fail(-1, "upload: engine is not in injected-stream mode");
            if ((r = need((size_t)P.C * P.ulen))) return r;
            return put_matrix(P.us, host, 1, (int)count, (int)count);
        case TDA_UP_DREAM_ARCHIVE0: {
            if (!P.archive) return fail(-1, "upload: proposal has no archive");
            const int M0 = P.dream_M0;
            if ((r = need((size_t)P.Cg * M0 * d))) return r;
            std::vector<R> h((size_t)M0 * P.Cg * d);
            for (long long g = 0; g < P.Cg; g++)
                for (int s = 0; s < M0; s++)
                    for (int k = 0; k < d; k++)
                        h[((size_t)s * P.Cg + g) * d + k] = (R)host[((size_t)g * M0 + s) * d + k];
            return put(P.archive, h);
        }
        case TDA_UP_AM_FACTORS: {
            if (!P.am_T) return fail(-1, "upload: proposal is not AdaptiveMetropolis");
            if ((r = need((size_t)P.C * d * d))) return r;
            std::vector<R> h((size_t)d * d * Cs, (R)0);
            for (int c = 0; c < P.C; c++)
                for (int e = 0; e < d * d; e++) h[(size_t)e * Cs + c] = (R)host[(size_t)c * d * d + e];
            return put(P.am_T, h);
        }
        default:
            return fail(-1, "upload: unknown item");
        }
    }

    // linear models: qoi_b = Q b + q0 from the device copies of Q, b and q0
    int compose_qoi_offset(int l) {
        const tda::LevelP<R>& v = P.lv[l];
        std::vector<R> Q((size_t)nq[l] * v.m), b(v.m), q0(nq[l]), o(nq[l]);
        CUDA_TRY(cudaMemcpy(Q.data(), qoi_Q[l], Q.size() * sizeof(R), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(b.data(), v.b, b.size() * sizeof(R), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(q0.data(), qoi_q0[l], q0.size() * sizeof(R), cudaMemcpyDeviceToHost));
        for (int j = 0; j < nq[l]; j++) {
            double a = q0[j];
            for (int n = 0; n < v.m; n++) a += (double)Q[(size_t)j * v.m + n] * b[n];
            o[j] = (R)a;
        }
        return put(qoi_b[l], o);
    }

    // Link.qoi of records [rec0, rec0 + nrec) of a level into out [nrec][nq][Cs]
    int qoi_fill(int l, long long rec0, long long nrec, R* out, cudaStream_t st) {
        const tda::LevelP<R>& v = P.lv[l];
        if (!nq[l]) return fail(-1, "qoi: the level's model has no quantity of interest");
        if (qoi_W[l]) {
            if (!v.h_theta) return fail(-1, "qoi: needs the stored parameters of the level");
            if (P.d > 64) return fail(-1, "history fill: d > 64");
            const long long per = Cs / 256;
            for (long long r0 = 0; r0 < nrec;) {
                const long long n = nrec - r0 < (1 << 20) ? nrec - r0 : (1 << 20);
                hist_fill_kernel<R, 64><<<(unsigned)(n * per), 256, 0, st>>>(v.h_theta + (size_t)(rec0 + r0) * P.d * Cs, qoi_W[l], ldq[l], nq[l],
                                                                          qoi_b[l], (const R*)nullptr, out + (size_t)r0 * nq[l] * Cs, P.d, Cs, 0, (R)0);
                g_launches++;
                r0 += n;
            }
        } else {
            if (!v.h_F) return fail(-1, "qoi: needs the stored model outputs of the level (store_model_output=True)");
            const long long per = Cs / 256;
            for (long long r0 = 0; r0 < nrec;) {
                const long long n = nrec - r0 < (1 << 20) ? nrec - r0 : (1 << 20);
                qoi_from_F_kernel<R><<<(unsigned)(n * per), 256, 0, st>>>(v.h_F + (size_t)(rec0 + r0) * v.m * Cs, qoi_Q[l], qoi_q0[l],
                                                                       out + (size_t)r0 * nq[l] * Cs, v.m, nq[l], Cs);
                g_launches++;
                r0 += n;
            }
        }
        CUDA_TRY(cudaGetLastError());
        return 0;
    }

    // ---- launches ----------------------------------------------------------------------------
    int launch(int mode, long long iterations, cudaStream_t st) {
        CUDA_TRY(cudaSetDevice(device));
        P.mode = mode;
        P.iterations = iterations;
        P.dream_slots = dream_slots;
        int occ = (sizeof(R) == 4) ? 2 : 1;
        int grid = n_tiles < sm_count * occ ? n_tiles : sm_count * occ;
        tda::chain_kernel<R><<<grid, tda::NT, smem_bytes, st>>>(P, kt);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        return 0;
    }

    int hist_fill(int l, const R* theta, long long nrec, R* out, int mode, cudaStream_t st) {
        if (P.d > 64) return fail(-1, "history fill: d > 64");
        const tda::LevelP<R>& v = P.lv[l];
        const long long per = Cs / 256;
        for (long long r0 = 0; r0 < nrec;) {
            const long long n = nrec - r0 < (1 << 20) ? nrec - r0 : (1 << 20);
            const R* th = theta + (size_t)r0 * P.d * Cs;
            if (mode == 0)
                hist_fill_kernel<R, 64><<<(unsigned)(n * per), 256, 0, st>>>(th, v.A, v.ldA, v.m, v.b, (const R*)nullptr,
                                                                          out + (size_t)r0 * v.m * Cs, P.d, Cs, 0, (R)0);
            else
                hist_fill_kernel<R, 64><<<(unsigned)(n * per), 256, 0, st>>>(th, P.LP, P.ldD, P.d, (const R*)nullptr, P.prior_mean,
                                                                          out + (size_t)r0 * Cs, P.d, Cs, 1, P.prior_logconst);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            r0 += n;
        }
        return 0;
    }
    // fills Link.prior (levels below the finest) and Link.model_output of the records the fp16-split
    // kernel wrote since the last call
    int fill_lazy_history(cudaStream_t st) {
        CUDA_TRY(cudaSetDevice(device));
        if (lazy_w_hi > lazy_w_lo && P.lv[0].h_theta) {
            // theta = w @ T, in place: a thread reads its (record, chain) column before it writes it
            const long long hi = lazy_w_hi < P.lv[0].hist_cap ? lazy_w_hi : P.lv[0].hist_cap;
            if (hi > lazy_w_lo) {
                if (P.d > 64) return fail(-1, "history fill: d > 64");
                R* th = P.lv[0].h_theta + (size_t)lazy_w_lo * P.d * Cs;
                const long long per = Cs / 256, nrec = hi - lazy_w_lo;
                for (long long r0 = 0; r0 < nrec;) {
                    const long long n = nrec - r0 < (1 << 20) ? nrec - r0 : (1 << 20);
                    hist_fill_kernel<R, 64><<<(unsigned)(n * per), 256, 0, st>>>(th + (size_t)r0 * P.d * Cs, P.T, P.ldD, P.d, (const R*)nullptr,
                                                                              (const R*)nullptr, th + (size_t)r0 * P.d * Cs, P.d, Cs, 0, (R)0);
                    g_launches++;
                    r0 += n;
                }
                CUDA_TRY(cudaGetLastError());
            }
        }
        lazy_w_lo = lazy_w_hi = 0;
        for (int l = 0; l < P.L; l++) {
            const tda::LevelP<R>& v = P.lv[l];
            const long long lo = lazy_lo[l], hi = lazy_hi[l] < v.hist_cap ? lazy_hi[l] : v.hist_cap;
            lazy_lo[l] = lazy_hi[l];
            if (hi <= lo || !v.h_theta) continue;
            const R* th = v.h_theta + (size_t)lo * P.d * Cs;
            int r = 0;
            if (l < P.L - 1 && (v.store & TDA_STORE_STATS)) r = hist_fill(l, th, hi - lo, v.h_prior + (size_t)lo * Cs, 1, st);
            if (!r && (v.store & TDA_STORE_OUTPUT)) r = hist_fill(l, th, hi - lo, v.h_F + (size_t)lo * v.m * Cs, 0, st);
            if (r) return r;
        }
        return 0;
    }
    // the fp16-split kernel keeps no model outputs: rebuild the levels' current F from theta before
    // another kernel (or a checkpoint) reads them
    int refresh_state_outputs(cudaStream_t st) {
        if (!state_F_stale) return 0;
        CUDA_TRY(cudaSetDevice(device));
        state_F_stale = false;
        for (int l = 0; l < P.L; l++) {
            tda::LevelP<R>& v = P.lv[l];
            if (!v.need_F || !v.F || v.model_kind != TDA_MODEL_LINEAR) continue;
            int r = hist_fill(l, v.theta, 1, v.F, 0, st);
            if (r) return r;
            for (int a = 0; a < tda::MAXL; a++)
                if (v.sv_F[a]) CUDA_TRY(cudaMemcpyAsync(v.sv_F[a], v.F, (size_t)v.m * Cs * sizeof(R), cudaMemcpyDeviceToDevice, st));
        }
        return 0;
    }

    int init(cudaStream_t st) override {
        P.t_base = 0; P.wcount = 0;
        if (tda::is_dream(P.prop_kind)) dream_slots = cfg.dream_M0;
        for (int l = 0; l < tda::MAXL; l++) { P.rec[l] = 0; P.lvl_steps[l] = 0; lazy_lo[l] = lazy_hi[l] = 0; }
        lazy_w_lo = lazy_w_hi = 0;
        state_F_stale = false;
        tcr.invalidate();
        tcr_unfit = false;
        tc16_state_checked = false;
        if (tc16.prepared) tc16_unfit = false;      // the operands fit; only a previous run's states did not
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaMemsetAsync(P.error_flag, 0, sizeof(int), st));
        int r = launch(tda::MODE_INIT, 0, st);
        if (r) return r;
        P.rec[P.L - 1] = 1;
        initialised = true;
        return 0;
    }

    // every tile of the lock-step kernel gets its own CTA and all of them fit on the device at once
    bool dream_persistent_ok() const {
        const int occ = (sizeof(R) == 4) ? 2 : 1;
        static const bool off = getenv("TDA_DREAM_PER_STEP_LAUNCH") != nullptr;
        if (!off && resolved_kernel() == 6) return true;      // its grid is co-resident by construction
        return !off && n_tiles <= sm_count * occ;
    }
    // warp-per-chain DREAM(Z) kernel: the configuration fits and every chain gets a register-resident slot
    bool dreamw_eligible() const {
        static const bool off = getenv("TDA_NO_DREAM_WARP") != nullptr;
        if (off || !tda::dream_warp_eligible(cfg)) return false;
        if (dreamw_grid < 0) const_cast<EngineT*>(this)->dreamw_grid = tda::dream_warp_grid<R>(P, sm_count, &const_cast<EngineT*>(this)->dreamw_occ);
        return dreamw_grid > 0;
    }
    // 0: every current state fits the fp16-split kernel's theta image, 1: not, < 0: error
    int check_tc16_state_range(cudaStream_t st) {
        if constexpr (sizeof(R) != 4) {
            return 0;
        } else {
            CUDA_TRY(cudaSetDevice(device));
            unsigned int* bits = reinterpret_cast<unsigned int*>(P.grid_bar);      // a spare device word (zeroed by the check)
            int r = tda::post::max_abs_f32(reinterpret_cast<const float*>(P.lv[P.L - 1].theta), (size_t)P.d * Cs, bits, st);
            if (r) return fail(r, tda::post::last_error());
            g_launches++;
            unsigned int h = 0;
            CUDA_TRY(cudaMemcpyAsync(&h, bits, sizeof(h), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            float m;
            memcpy(&m, &h, sizeof(m));
            return (m < tc16.theta_limit) ? 0 : 1;
        }
    }
    bool tc_eligible() const { return tc.eligible(cfg, P); }
    bool tc16_eligible() const { return !tc16_unfit && tc16.eligible(cfg, P); }
    bool tcr_eligible() const { return !tcr_unfit && tcr.eligible(cfg, P); }
    bool reg_eligible() const { return tda::mh_reg_eligible(cfg, P.lv[0].need_F != 0) && !z_round_user; }
    bool mldaw_eligible() const {
        static const bool off = getenv("TDA_NO_MLDA_WARP") != nullptr;
        if (off || z_round_user || !tda::mlda_warp_eligible(cfg)) return false;
        // A warp per chain pays when the per-chain work is wide (Poisson grids, 31 x 31 error-model matrices: cfg4 2.9 x
        // the lock-step kernel at 32768 chains) or when there are too few chains to fill the lock-step kernel's
        // 128-chain tiles (one round of either kernel: 38 against 119 us per finest iteration on the mlda3_linear
        // fixture).  Small problems at tens of thousands of chains stay on the lock-step kernel, which deals a chain
        // to a thread (2.8e8 against 6.3e7 finest transitions/s on that fixture at 32768 chains).
        bool poisson = true;
        for (int l = 0; l < cfg.n_levels; l++) poisson = poisson && cfg.level[l].model_kind == TDA_MODEL_POISSON1D;
        const bool wide = poisson || (cfg.aem && cfg.level[0].m >= 16);
        return wide || cfg.n_chains <= 4096;
    }
    // which kernel tda_engine_run launches: 1 generic, 2 tensor-core 3xTF32, 3 tensor-core fp16 split,
    // 4 register-resident single-level
    int resolved_kernel() const {
        if (kernel_choice >= 1 && kernel_choice <= 7) return kernel_choice;
        // fixed common step size: the fp16-split kernel with the step folded into its operators; per-chain /
        // adaptively scaled steps: the whitened-state kernel (TDA_PREFER_TCR: always the latter when it can)
        static const bool prefer_tcr = getenv("TDA_PREFER_TCR") != nullptr;
        if (prefer_tcr && tcr_eligible()) return 5;
        if (tc16_eligible()) return 3;
        if (tcr_eligible()) return 5;
        if (tc_eligible()) return 2;
        if (reg_eligible()) return 4;
        if (dreamw_eligible()) return 6;
        if (mldaw_eligible()) return 7;
        return 1;
    }
    // the fp16-split kernel consumes the z16 normal stream; the others do on request
    int z_round_effective() const {
        if (cfg.dtype != TDA_F32 || P.rng_mode != TDA_RNG_PHILOX) return 0;
        const int k = resolved_kernel();
        return (k == 3 || k == 5 || z_round_user) ? 1 : 0;
    }

    int run(long long iterations, cudaStream_t st, bool record = true) override {
        if (!initialised) return fail(-1, "run: call tda_engine_init first");
        if (iterations <= 0) return 0;
        const int L = P.L;
        if (record) {
            // a level that stores history must be able to hold every record of this run (the kernels
            // drop records past the capacity)
            long long st_l = iterations;
            for (int l = L - 1; l >= 0; l--) {
                if (l < L - 1) st_l *= P.J[l];
                if (!burning && cfg.level[l].store && P.rec[l] + st_l > P.lv[l].hist_cap)
                    return fail(-1, "run: level " + std::to_string(l) + " would write records " + std::to_string(P.rec[l]) + ".." +
                                        std::to_string(P.rec[l] + st_l - 1) + " but its history holds " + std::to_string(P.lv[l].hist_cap) +
                                        " (hist_capacity): fetch and tda_history_reset, or use tda_engine_burn for unrecorded iterations");
            }
        } else {
            // burn-in: the same transitions with every store flag off for the launch(es)
            int saved[tda::MAXL];
            for (int l = 0; l < L; l++) { saved[l] = P.lv[l].store; P.lv[l].store = 0; }
            long long rec_saved[tda::MAXL], lo_s[tda::MAXL], hi_s[tda::MAXL];
            for (int l = 0; l < tda::MAXL; l++) { rec_saved[l] = P.rec[l]; lo_s[l] = lazy_lo[l]; hi_s[l] = lazy_hi[l]; }
            burning = true;
            const int r = run(iterations, st, true);
            burning = false;
            for (int l = 0; l < L; l++) P.lv[l].store = saved[l];
            for (int l = 0; l < tda::MAXL; l++) { P.rec[l] = rec_saved[l]; lazy_lo[l] = lo_s[l]; lazy_hi[l] = hi_s[l]; }
            return r;
        }
        if (P.prop_kind == TDA_PROP_AM && !P.am_device_refactor) {
            long long to_boundary = P.period - (P.t_base % P.period);
            if (iterations > to_boundary)
                return fail(-1, "run: AdaptiveMetropolis with host refactoring must stop at period boundaries");
        }
        if (tda::is_dream(P.prop_kind)) {
            long long base_steps = iterations;
            for (int l = 0; l + 1 < L; l++) base_steps *= P.J[l];
            if (dream_slots + base_steps > P.dream_cap)
                return fail(-1, "run: the DREAM(Z) archive (dream_capacity slots) cannot hold one row per base-level step of this run");
        }
        if (kernel_choice == 2 && !tc_eligible()) return fail(-1, "run: tensor-core DA kernel does not support this configuration");
        if (kernel_choice == 3 && !tc16.eligible(cfg, P)) return fail(-1, "run: fp16-split tensor-core DA kernel does not support this configuration");
        if (kernel_choice == 4 && !reg_eligible()) return fail(-1, "run: register-resident kernel does not support this configuration");
        if (kernel_choice == 5 && !tcr.eligible(cfg, P)) return fail(-1, "run: whitened-state tensor-core DA kernel does not support this configuration");
        if (kernel_choice == 7 && !tda::mlda_warp_eligible(cfg)) return fail(-1, "run: warp-per-chain MLDA kernel does not support this configuration");
        if (kernel_choice == 6 && !dreamw_eligible()) return fail(-1, "run: warp-per-chain DREAM kernel does not support this configuration");
        int which = resolved_kernel();
        if (P.dream_sync > 1 && which != 6)
            return fail(-1, "run: DREAM sync_every > 1 needs the warp-per-chain DREAM kernel (d <= 32, linear model, fixed crossover distribution)");
        if (which == 5) {
            // operand images on first use, whitened state when theta was last written by someone else; a
            // problem (or a chain state) outside the fp16 range falls back to the older kernels
            int r = tcr.ready(P, cfg, st);
            if (r < 0) return fail(r, tcr.err);
            if (r > 0) {
                if (kernel_choice == 5) return fail(-1, "run: " + tcr.err);
                tcr_unfit = true;
                which = resolved_kernel();
            }
        }
        if (which == 3 && !tc16.prepared) {
            // operand scaling happens on first use; a problem that does not fit fp16 falls back
            if (!copy_stream) copy_stream = g_small.get_stream(device);
            tc16.fetch_stream = copy_stream;
            int r = tc16.prepare(P, cfg);
            if (r < 0) return fail(r, tc16.err);
            if (r > 0) {
                if (kernel_choice == 3) return fail(-1, "run: " + tc16.err);
                tc16_unfit = true;
                which = resolved_kernel();
            }
        }
        if (which == 3 && !tc16_state_checked) {
            // The theta operand image is fixed-point at a scale chosen from the prior (|mean| + 12 sd): a CURRENT
            // state outside it (initial parameters far from the prior, a state another kernel left) would
            // overflow to inf and freeze the chain.  Checked once per hand-over; such a job runs on another kernel.
            int r = check_tc16_state_range(st);
            if (r < 0) return r;
            if (r > 0) {
                if (kernel_choice == 3) return fail(-1, "run: a chain state lies outside the fp16 operand range of the fp16-split kernel");
                tc16_unfit = true;
                which = resolved_kernel();
                if (which == 5) {
                    int r5 = tcr.ready(P, cfg, st);
                    if (r5 < 0) return fail(r5, tcr.err);
                    if (r5 > 0) { tcr_unfit = true; which = resolved_kernel(); }
                }
            } else {
                tc16_state_checked = true;
            }
        }
        P.z_round = z_round_effective();
        int r;
        // the tensor-core kernels record parameters only; prior / model outputs of the records are filled on first fetch
        const bool lazy_kernel = (which == 2 || which == 3 || which == 5);
        if (!lazy_kernel) {
            r = fill_lazy_history(st);
            if (!r) r = refresh_state_outputs(st);
            if (r) return r;
        }
        if (which != 5) tcr.invalidate();          // theta moves without the whitened copy
        if (which != 3) tc16_state_checked = false;
        if ((which == 3 || which == 5) && iterations > (1 << 20)) {
            // the fp16-split kernel counts its coarse steps per launch in 32 bits
            for (long long done = 0; done < iterations;) {
                const long long n = iterations - done < (1 << 20) ? iterations - done : (1 << 20);
                r = run(n, st, true);
                if (r) return r;
                done += n;
            }
            return 0;
        }
        if (which == 5) {
            r = tcr.run(P, cfg, iterations, sm_count, st);
            if (r) return fail(r < 0 ? r : -1, tcr.err);
            g_launches++;
        } else if (which == 3) {
            r = tc16.run(P, cfg, iterations, sm_count, st);
            if (r) return fail(r, tc16.err);
            g_launches++;
        } else if (which == 2) {
            r = tc.run(P, cfg, iterations, sm_count, st);
            if (r) return fail(r, tc.err);
            g_launches++;
        } else if (which == 4) {
            CUDA_TRY(cudaSetDevice(device));
            P.mode = tda::MODE_RUN;
            P.iterations = iterations;
            CUDA_TRY(tda::mh_reg_launch<R>(P, st));
            g_launches++;
        } else if (which == 7) {
            CUDA_TRY(cudaSetDevice(device));
            if (P.aem && !mw_sig) {
                mw_cls = DevPool::size_class(tda::mlda_warp_image_elems(P.C) * sizeof(R));
                cudaError_t e1 = g_pool.alloc(&mw_sig, mw_cls, device);
                cudaError_t e2 = e1 == cudaSuccess ? g_pool.alloc(&mw_li, mw_cls, device) : e1;
                if (e1 != cudaSuccess || e2 != cudaSuccess) {
                    if (mw_sig) { g_pool.release(mw_sig, mw_cls, device); mw_sig = nullptr; }
                    return fail(-3, std::string("warp-per-chain MLDA kernel: error-model images: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
                }
                CUDA_TRY(cudaMemsetAsync(mw_sig, 0, mw_cls, st));
                CUDA_TRY(cudaMemsetAsync(mw_li, 0, mw_cls, st));
            }
            if (!mw_phi) {
                mw_phi_cls = DevPool::size_class(tda::mlda_warp_phi_elems(cfg) * sizeof(R));
                cudaError_t e3 = g_pool.alloc(&mw_phi, mw_phi_cls, device);
                if (e3 != cudaSuccess) return fail(-3, std::string("warp-per-chain MLDA kernel: Phi images: ") + cudaGetErrorString(e3));
            }
            P.mode = tda::MODE_RUN;
            P.iterations = iterations;
            r = tda::mlda_warp_run<R>(P, mw_sig, mw_li, mw_phi, sm_count, st);
            if (r) return fail(r, tda::mlda_warp_last_error());
            g_launches += tda::mlda_warp_launches(L, P.aem);
        } else if (which == 6) {
            // one warp per chain, one persistent launch; the shared-archive variant ends every step in the grid
            // barrier / peer-memory exchange
            CUDA_TRY(cudaSetDevice(device));
            P.mode = tda::MODE_RUN;
            P.iterations = iterations;
            P.dream_slots = dream_slots;
            P.grid_sync = (P.prop_kind == TDA_PROP_DREAM) ? 1 : 0;
            if (P.grid_sync) {
                CUDA_TRY(cudaMemsetAsync(P.grid_bar, 0, sizeof(unsigned int), st));
                // barriers of this launch: one per multiple of dream_sync among the steps it takes
                const long long K = P.dream_sync, done0 = dream_slots - P.dream_M0;
                const unsigned int nbar = (unsigned int)((done0 + iterations) / K - done0 / K);
                if (P.n_peers > 1 && P.arrive_mode) {
                    P.arr_base = arr_next;
                    arr_next += nbar;
                } else {
                    P.flag_base = flag_next;
                    flag_next += nbar;
                }
            }
            const int arrive_saved = P.arrive_mode;
            if (P.n_peers <= 1) P.arrive_mode = 0;
            r = tda::dream_warp_launch<R>(P, dreamw_grid, dreamw_occ, st);
            P.arrive_mode = arrive_saved;
            P.grid_sync = 0;
            if (r) return fail(r, tda::dream_warp_last_error());
            g_launches++;
        } else if (P.prop_kind == TDA_PROP_DREAM && (iterations > 1 || P.n_peers > 1) && dream_persistent_ok()) {
            // shared archive, lock-step visibility (every chain sees all rows through the previous step): ONE
            // persistent launch whose steps end in a grid-wide barrier (and, over several GPUs, in the exchange
            // of the new rows through peer memory); every tile has its own co-resident CTA
            CUDA_TRY(cudaSetDevice(device));
            CUDA_TRY(cudaMemsetAsync(P.grid_bar, 0, sizeof(unsigned int), st));
            P.grid_sync = 1;
            P.flag_base = flag_next;
            flag_next += (unsigned int)iterations;
            r = launch(tda::MODE_RUN, iterations, st);
            P.grid_sync = 0;
            if (r) return r;
        } else if (P.prop_kind == TDA_PROP_DREAM && iterations > 1) {
            // more tiles than resident CTAs: one launch per step is the grid-wide boundary
            if (P.n_peers > 1) return fail(-1, "run: the peer-memory archive exchange needs every tile of the ensemble resident (fewer chains per GPU)");
            for (long long i = 0; i < iterations; i++) {
                r = run(1, st, true);
                if (r) return r;
            }
            return 0;
        } else {
            r = launch(tda::MODE_RUN, iterations, st);
            if (r) return r;
        }
        // advance the uniform counters exactly like the kernel did
        long long steps[tda::MAXL];
        steps[L - 1] = iterations;
        for (int l = L - 2; l >= 0; l--) steps[l] = steps[l + 1] * P.J[l];
        P.t_base += steps[0];
        long long w = steps[0];
        for (int l = 1; l < L; l++) w += steps[l];
        P.wcount += (L == 1) ? steps[0] : w;
        if (which == 5 && !burning && (cfg.level[0].store & TDA_STORE_THETA)) {
            if (lazy_w_lo == lazy_w_hi) lazy_w_lo = P.rec[0];
            lazy_w_hi = P.rec[0] + steps[0];
        }
        if (which == 2 || which == 3 || which == 5) {
            if (!burning)
                for (int l = 0; l < L; l++) {
                    if (lazy_lo[l] == lazy_hi[l]) lazy_lo[l] = P.rec[l];
                    lazy_hi[l] = P.rec[l] + steps[l];
                }
            state_F_stale = true;
        }
        for (int l = 0; l < L; l++) { P.rec[l] += steps[l]; if (l >= 1) P.lvl_steps[l] += steps[l]; }
        if (tda::is_dream(P.prop_kind)) dream_slots += steps[0];
        return 0;
    }

    int history_reset() override {
        for (int l = 0; l < P.L; l++) { P.rec[l] = 0; lazy_lo[l] = lazy_hi[l] = 0; }
        lazy_w_lo = lazy_w_hi = 0;
        return 0;
    }

    int select_kernel(int which) override {
        kernel_choice = which;
        return 0;
    }

    // Checkpoint: the host-side counters + every device buffer except the history (constants, chain
    // state of every level, proposal state, stream cursors).  load == 0: engine -> host blob,
    // load == 1: host blob -> engine (created with the same configuration).
    struct StateHeader {
        unsigned long long magic;
        tda_config cfg;
        long long t_base, wcount, rec[tda::MAXL], lvl_steps[tda::MAXL], dream_slots;
        int initialised, n_buffers;
        unsigned long long payload;
    };
    int state_io(void* host, size_t bytes, size_t* needed, int load) override {
        CUDA_TRY(cudaSetDevice(device));
        size_t payload = 0;
        int nb = 0;
        for (size_t i = 0; i < allocs.size(); i++)
            if (alloc_is_state[i] && allocs[i] != (void*)stage_theta) { payload += alloc_bytes[i]; nb++; }
        const size_t total = sizeof(StateHeader) + payload;
        if (needed) *needed = total;
        if (!host) return 0;
        if (bytes < total) return fail(-1, "state: buffer too small");
        if (!load) {
            int r = refresh_state_outputs(0);
            if (r) return r;
        } else {
            state_F_stale = false;
            tcr.invalidate();
            tc16_state_checked = false;
            for (int l = 0; l < tda::MAXL; l++) lazy_lo[l] = lazy_hi[l] = 0;
        }
        CUDA_TRY(cudaDeviceSynchronize());
        StateHeader h;
        char* q = reinterpret_cast<char*>(host) + sizeof(StateHeader);
        if (!load) {
            memset(&h, 0, sizeof(h));
            h.magic = 0x7464615f73746174ull;
            h.cfg = cfg;
            h.t_base = P.t_base; h.wcount = P.wcount; h.dream_slots = dream_slots;
            for (int l = 0; l < tda::MAXL; l++) { h.rec[l] = P.rec[l]; h.lvl_steps[l] = P.lvl_steps[l]; }
            h.initialised = initialised ? 1 : 0; h.n_buffers = nb; h.payload = payload;
            memcpy(host, &h, sizeof(h));
        } else {
            memcpy(&h, host, sizeof(h));
            tda_config a = h.cfg, b = cfg;
            // the history capacity may differ between the engine that saved and the one that loads
            for (int l = 0; l < TDA_MAX_LEVELS; l++) { a.level[l].hist_capacity = 0; b.level[l].hist_capacity = 0; a.level[l].store = b.level[l].store; }
            if (h.magic != 0x7464615f73746174ull || h.n_buffers != nb || h.payload != payload || memcmp(&a, &b, sizeof(a)) != 0)
                return fail(-1, "state: blob was saved by an engine with a different configuration");
            P.t_base = h.t_base; P.wcount = h.wcount; dream_slots = h.dream_slots;
            for (int l = 0; l < tda::MAXL; l++) { P.rec[l] = 0; P.lvl_steps[l] = h.lvl_steps[l]; }
            initialised = h.initialised != 0;
        }
        for (size_t i = 0; i < allocs.size(); i++) {
            if (!alloc_is_state[i] || allocs[i] == (void*)stage_theta) continue;
            if (!load) CUDA_TRY(cudaMemcpy(q, allocs[i], alloc_bytes[i], cudaMemcpyDeviceToHost));
            else CUDA_TRY(cudaMemcpy(allocs[i], q, alloc_bytes[i], cudaMemcpyHostToDevice));
            q += alloc_bytes[i];
        }
        return 0;
    }

    // ---- downloads ---------------------------------------------------------------------------
    int fetch(int level, int field, long long rec0, long long nrec, void* dst, size_t dst_bytes, size_t* bytes,
              cudaStream_t st) override {
        CUDA_TRY(cudaSetDevice(device));
        if (level < 0 || level >= P.L) return fail(-1, "fetch: bad level");
        const tda::LevelP<R>& v = P.lv[level];
        if (rec0 < 0 || nrec < 0 || rec0 + nrec > v.hist_cap) return fail(-1, "fetch: record range outside the history buffer");
        if (field == TDA_F_PRIOR || field == TDA_F_OUTPUT || (field == TDA_F_THETA && lazy_w_hi > lazy_w_lo)) {
            int r = fill_lazy_history(st);
            if (r) return r;
        }
        if (field == TDA_F_QOI) {
            // rebuilt into a scratch buffer, copied out, released: all in stream order
            const size_t n = (size_t)nrec * nq[level];
            const size_t need = n * (size_t)P.C * sizeof(R);
            if (bytes) *bytes = need;
            if (!nq[level]) return fail(-1, "fetch: the level's model has no quantity of interest");
            if (dst_bytes < need) return fail(-1, "fetch: destination too small");
            if (n == 0) return 0;
            if (!qoi_W[level]) { int r0 = fill_lazy_history(st); if (r0) return r0; }
            R* tmp = nullptr;
            CUDA_TRY(cudaMallocAsync((void**)&tmp, n * Cs * sizeof(R), st));
            int r = qoi_fill(level, rec0, nrec, tmp, st);
            if (!r) {
                cudaError_t ce = cudaMemcpy2DAsync(dst, (size_t)P.C * sizeof(R), tmp, (size_t)Cs * sizeof(R), (size_t)P.C * sizeof(R), n,
                                                   cudaMemcpyDeviceToHost, st);
                if (ce != cudaSuccess) r = fail(-2, std::string("fetch qoi: ") + cudaGetErrorString(ce));
            }
            cudaFreeAsync(tmp, st);
            return r;
        }
        const void* src = nullptr;
        size_t rows = 0, esz = sizeof(R);
        switch (field) {
        case TDA_F_THETA: src = v.h_theta ? v.h_theta + (size_t)rec0 * P.d * Cs : nullptr; rows = (size_t)nrec * P.d; break;
        case TDA_F_PRIOR: src = v.h_prior ? v.h_prior + (size_t)rec0 * Cs : nullptr; rows = (size_t)nrec; break;
        case TDA_F_LIKE: src = v.h_like ? v.h_like + (size_t)rec0 * Cs : nullptr; rows = (size_t)nrec; break;
        case TDA_F_OUTPUT: src = v.h_F ? v.h_F + (size_t)rec0 * v.m * Cs : nullptr; rows = (size_t)nrec * v.m; break;
        case TDA_F_ACCEPT: src = v.h_acc ? v.h_acc + (size_t)rec0 * Cs : nullptr; rows = (size_t)nrec; esz = 1; break;
        default: return fail(-1, "fetch: unknown field");
        }
        if (!src) return fail(-1, "fetch: field was not stored for this level");
        size_t need = rows * (size_t)P.C * esz;
        if (bytes) *bytes = need;
        if (dst_bytes < need) return fail(-1, "fetch: destination too small");
        if (rows == 0) return 0;
        if (P.C == Cs)
            CUDA_TRY(cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, st));
        else
            CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)P.C * esz, src, (size_t)Cs * esz, (size_t)P.C * esz, rows,
                                       cudaMemcpyDeviceToHost, st));
        return 0;
    }

    // ---- compacted history (tda_post.h) ------------------------------------------------------------
    template <typename T>
    int grow(T** ptr, size_t* cap, size_t need_bytes) {
        if (*cap >= need_bytes && *ptr) return 0;
        if (*ptr) { CUDA_TRY(cudaDeviceSynchronize()); g_pool.release(*ptr, *cap, device); }
        *ptr = nullptr; *cap = 0;
        void* q = nullptr;
        const size_t cls = DevPool::size_class(need_bytes ? need_bytes : 1);
        cudaError_t e = g_pool.alloc(&q, cls, device);
        if (e != cudaSuccess) return fail(-3, std::string("cudaMalloc ") + std::to_string(need_bytes) + " bytes (compacted history): " + cudaGetErrorString(e));
        *ptr = reinterpret_cast<T*>(q); *cap = cls;
        return 0;
    }

    int compact_begin(int level, long long rec0, long long nrec, int first_is_full, int fields, int slot, cudaStream_t st) override {
        CUDA_TRY(cudaSetDevice(device));
        if (slot < 0 || slot > 1) return fail(-1, "compact: slot must be 0 or 1");
        if (level != P.L - 1) return fail(-1, "compact: only the finest level's records repeat on rejection");
        const tda::LevelP<R>& v = P.lv[level];
        if (!v.h_acc) return fail(-1, "compact: the level does not store accept flags");
        if (rec0 < 0 || nrec < 1 || rec0 + nrec > P.rec[level] || rec0 + nrec > v.hist_cap) return fail(-1, "compact: record range has not been written");
        if ((fields & TDA_STORE_THETA) && !v.h_theta) return fail(-1, "compact: parameters were not stored");
        if ((fields & TDA_STORE_STATS) && !(v.h_prior && v.h_like)) return fail(-1, "compact: log-densities were not stored");
        if ((fields & TDA_STORE_OUTPUT) && !v.h_F) return fail(-1, "compact: model outputs were not stored");
        if ((fields & TDA_STORE_QOI) && !nq[level]) return fail(-1, "compact: the level's model has no quantity of interest");
        CompactSlot& s = cslot[slot];
        if (!copy_stream) copy_stream = g_small.get_stream(device);
        if (!s.ready) { s.ready = g_small.get_event(device); s.copied = g_small.get_event(device); }
        if (!s.total_pinned) s.total_pinned = g_small.get_pinned();
        if (!copy_stream || !s.ready || !s.copied || !s.total_pinned) return fail(-2, "compact: could not create the copy stream / events / pinned scalar");
        if (!s.offsets) {
            int r0 = grow(&s.offsets, &s.offsets_cap, ((size_t)P.C + 1) * sizeof(long long));
            if (!r0) r0 = grow(&s.scratch, &s.scratch_cap, ((size_t)(P.C + 255) / 256 + 2) * sizeof(long long));
            if (r0) return r0;
        }
        // copies of the slot's previous contents must have left the device buffers
        if (s.has_copies) CUDA_TRY(cudaStreamWaitEvent(st, s.copied, 0));
        if (fields & (TDA_STORE_OUTPUT | TDA_STORE_STATS | TDA_STORE_QOI)) { int r0 = fill_lazy_history(st); if (r0) return r0; }
        const size_t rows_max = (size_t)nrec * P.C;
        int r = grow(&s.flags, &s.flags_cap, (size_t)nrec * Cs);
        const int widths[5] = {P.d, 1, 1, v.m, nq[level]};
        const bool want[5] = {(fields & TDA_STORE_THETA) != 0, (fields & TDA_STORE_STATS) != 0, (fields & TDA_STORE_STATS) != 0,
                              (fields & TDA_STORE_OUTPUT) != 0, (fields & TDA_STORE_QOI) != 0};
        for (int f = 0; f < 5 && !r; f++)
            if (want[f]) r = grow(&s.buf[f], &s.cap[f], rows_max * widths[f] * sizeof(R));
        if (r) return r;
        const uint8_t* acc = v.h_acc + (size_t)rec0 * Cs;
        namespace tp = tda::post;
        r = tp::compact_offsets(acc, nrec, P.C, Cs, first_is_full, s.offsets, s.scratch, st);
        if (!r) r = tp::compact_flags(acc, nrec, P.C, Cs, first_is_full, s.flags, st);
        g_launches += 4;
        CUDA_TRY(cudaMemcpyAsync(s.total_pinned, s.offsets + P.C, sizeof(long long), cudaMemcpyDeviceToHost, st));
        const void* srcs[5] = {v.h_theta ? v.h_theta + (size_t)rec0 * P.d * Cs : nullptr, v.h_prior ? v.h_prior + (size_t)rec0 * Cs : nullptr,
                               v.h_like ? v.h_like + (size_t)rec0 * Cs : nullptr, v.h_F ? v.h_F + (size_t)rec0 * v.m * Cs : nullptr, nullptr};
        R* qtmp = nullptr;
        if (want[4] && !r) {
            CUDA_TRY(cudaMallocAsync((void**)&qtmp, (size_t)nrec * nq[level] * Cs * sizeof(R), st));
            r = qoi_fill(level, rec0, nrec, qtmp, st);
            srcs[4] = qtmp;
        }
        // the log-densities ride with the parameter rows (one pass over the records) when both are wanted
        const bool fused_stats = want[0] && want[1];
        for (int f = 0; f < 5 && !r; f++)
            if (want[f]) {
                if (fused_stats && (f == 1 || f == 2)) continue;
                if (f == 0 && fused_stats)
                    r = tp::compact_gather(srcs[0], (int)sizeof(R), widths[0], acc, nrec, P.C, Cs, first_is_full, s.offsets, s.buf[0], st,
                                           srcs[1], srcs[2], s.buf[1], s.buf[2]);
                else
                    r = tp::compact_gather(srcs[f], (int)sizeof(R), widths[f], acc, nrec, P.C, Cs, first_is_full, s.offsets, s.buf[f], st);
                g_launches++;
            }
        if (qtmp) cudaFreeAsync(qtmp, st);
        if (r) return fail(r, tp::last_error());
        CUDA_TRY(cudaEventRecord(s.ready, st));
        s.nrec = nrec; s.fields = fields; s.n_rows = -1; s.pending = true; s.has_copies = false;
        return 0;
    }

    int compact_rows(int slot, long long* n_rows) override {
        CUDA_TRY(cudaSetDevice(device));
        if (slot < 0 || slot > 1 || !cslot[slot].pending) return fail(-1, "compact: no compaction was begun on this slot");
        CompactSlot& s = cslot[slot];
        if (s.n_rows < 0) {
            CUDA_TRY(cudaEventSynchronize(s.ready));
            s.n_rows = *s.total_pinned;
        }
        *n_rows = s.n_rows;
        return 0;
    }

    int compact_fetch(int slot, int field, void* dst, size_t dst_bytes, size_t* bytes) override {
        long long n = 0;
        int r = compact_rows(slot, &n);
        if (r) return r;
        CompactSlot& s = cslot[slot];
        const tda::LevelP<R>& v = P.lv[P.L - 1];
        const void* src = nullptr;
        size_t need = 0;
        bool strided = false;
        switch (field) {
        case TDA_F_ACCEPT: src = s.flags; need = (size_t)s.nrec * P.C; strided = (P.C != Cs); break;
        case TDA_CF_OFFSETS: src = s.offsets; need = ((size_t)P.C + 1) * sizeof(long long); break;
        case TDA_F_THETA: src = (s.fields & TDA_STORE_THETA) ? s.buf[0] : nullptr; need = (size_t)n * P.d * sizeof(R); break;
        case TDA_F_PRIOR: src = (s.fields & TDA_STORE_STATS) ? s.buf[1] : nullptr; need = (size_t)n * sizeof(R); break;
        case TDA_F_LIKE: src = (s.fields & TDA_STORE_STATS) ? s.buf[2] : nullptr; need = (size_t)n * sizeof(R); break;
        case TDA_F_OUTPUT: src = (s.fields & TDA_STORE_OUTPUT) ? s.buf[3] : nullptr; need = (size_t)n * v.m * sizeof(R); break;
        case TDA_F_QOI: src = (s.fields & TDA_STORE_QOI) ? s.buf[4] : nullptr; need = (size_t)n * nq[P.L - 1] * sizeof(R); break;
        default: return fail(-1, "compact fetch: unknown field");
        }
        if (!src) return fail(-1, "compact fetch: field was not compacted on this slot");
        if (bytes) *bytes = need;
        if (dst_bytes < need) return fail(-1, "compact fetch: destination too small");
        CUDA_TRY(cudaStreamWaitEvent(copy_stream, s.ready, 0));
        if (need) {
            if (strided) CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)P.C, src, (size_t)Cs, (size_t)P.C, (size_t)s.nrec, cudaMemcpyDeviceToHost, copy_stream));
            else CUDA_TRY(cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, copy_stream));
        }
        CUDA_TRY(cudaEventRecord(s.copied, copy_stream));
        s.has_copies = true;
        return 0;
    }

    // ---- shared archive over several GPUs: the replicas and the step flags are mapped into each other ----
    struct PeerHandles {
        cudaIpcMemHandle_t archive, flags;
        int dreamw_grid;          // CTAs of this rank's warp-per-chain DREAM kernel (0: it runs the lock-step kernel)
        int pad[15];
    };
    int peer_export(void* out, size_t bytes, size_t* needed) override {
        CUDA_TRY(cudaSetDevice(device));
        if (needed) *needed = sizeof(PeerHandles);
        if (!out) return 0;
        if (bytes < sizeof(PeerHandles)) return fail(-1, "peer export: buffer too small");
        if (!P.archive || !dream_flags) return fail(-1, "peer export: the proposal has no shared archive");
        PeerHandles h;
        CUDA_TRY(cudaIpcGetMemHandle(&h.archive, P.archive));
        CUDA_TRY(cudaIpcGetMemHandle(&h.flags, dream_flags));
        memset(h.pad, 0, sizeof(h.pad));
        h.dreamw_grid = (resolved_kernel() == 6) ? dreamw_grid : 0;
        memcpy(out, &h, sizeof(h));
        return 0;
    }
    int peer_import(int n_ranks, int my_rank, const void* handles, size_t bytes) override {
        CUDA_TRY(cudaSetDevice(device));
        if (n_ranks < 1 || n_ranks > 8 || my_rank < 0 || my_rank >= n_ranks) return fail(-1, "peer import: 1..8 ranks");
        if (bytes < (size_t)n_ranks * sizeof(PeerHandles)) return fail(-1, "peer import: one handle block per rank expected");
        if (!P.archive || !dream_flags) return fail(-1, "peer import: the proposal has no shared archive");
        if (!dream_persistent_ok()) return fail(-1, "peer import: the ensemble's tiles must all be resident for the in-kernel exchange");
        const PeerHandles* h = reinterpret_cast<const PeerHandles*>(handles);
        for (int r = 0; r < n_ranks; r++) {
            if (r == my_rank) { P.peer_archive[r] = P.archive; P.peer_flags[r] = dream_flags; continue; }
            void *a = nullptr, *f = nullptr;
            cudaError_t e1 = cudaIpcOpenMemHandle(&a, h[r].archive, cudaIpcMemLazyEnablePeerAccess);
            cudaError_t e2 = e1 == cudaSuccess ? cudaIpcOpenMemHandle(&f, h[r].flags, cudaIpcMemLazyEnablePeerAccess) : e1;
            if (e1 != cudaSuccess || e2 != cudaSuccess) {
                cudaGetLastError();
                if (a) cudaIpcCloseMemHandle(a);
                return fail(-2, std::string("peer import: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            }
            peer_mapped.push_back(a); peer_mapped.push_back(f);
            P.peer_archive[r] = reinterpret_cast<R*>(a);
            P.peer_flags[r] = reinterpret_cast<unsigned int*>(f);
        }
        P.n_peers = n_ranks; P.my_rank = my_rank;
        // arrival counters instead of the barrier + flag handshake when every rank runs the warp-per-chain kernel
        P.arrive_mode = 1;
        for (int r = 0; r < 8; r++) P.peer_grid[r] = 0;
        for (int r = 0; r < n_ranks; r++) {
            P.peer_grid[r] = h[r].dreamw_grid;
            if (h[r].dreamw_grid <= 0) P.arrive_mode = 0;
        }
        if (getenv("TDA_DREAM_FLAG_HANDSHAKE")) P.arrive_mode = 0;
        return 0;
    }

    // rank-normalised split-chain diagnostics of a level's parameter history, on the device (tda_ess.cu)
    int ess_sums(int level, long long rec0, long long nrec, int n_lag, double* sums, double* folded, cudaStream_t st) override {
        CUDA_TRY(cudaSetDevice(device));
        if (level < 0 || level >= P.L) return fail(-1, "ess: bad level");
        const tda::LevelP<R>& v = P.lv[level];
        if (!v.h_theta) return fail(-1, "ess: the level does not store parameters");
        if (rec0 < 0 || nrec < 4 || rec0 + nrec > P.rec[level] || rec0 + nrec > v.hist_cap) return fail(-1, "ess: record range has not been written");
        if (level == 0 && lazy_w_hi > lazy_w_lo) { int r0 = fill_lazy_history(st); if (r0) return r0; }
        const int nh = (int)(nrec / 2);
        if (n_lag < 1 || n_lag > nh) n_lag = nh;
        namespace tp = tda::post;
        tp::EssWorkspace* w = tp::ess_workspace_create(nrec, P.C, Cs, n_lag);
        if (!w) return fail(-3, tp::ess_last_error());
        int r = 0;
        for (int k = 0; k < P.d && !r; k++) {
            r = tp::ess_sums(w, v.h_theta + ((size_t)rec0 * P.d + k) * Cs, (int)sizeof(R), (long long)P.d * Cs,
                             sums + (size_t)k * (n_lag + 4), folded + (size_t)k * 4, st);
            g_launches += 10;
        }
        tp::ess_workspace_destroy(w);
        if (r) return fail(r, tp::ess_last_error());
        return 0;
    }

    int compact_sync() override {
        CUDA_TRY(cudaSetDevice(device));
        if (copy_stream) CUDA_TRY(cudaStreamSynchronize(copy_stream));
        return 0;
    }

    template <typename T>
    int get_soa(const T* src, int rows, void* dst, size_t bytes, bool as_double, bool transpose) {
        // device [rows][Cs] -> host [rows][C] or (transpose) [C][rows]; optionally widened to float64
        std::vector<T> h((size_t)rows * Cs);
        CUDA_TRY(cudaMemcpy(h.data(), src, h.size() * sizeof(T), cudaMemcpyDeviceToHost));
        size_t need = (size_t)rows * P.C * (as_double ? sizeof(double) : sizeof(T));
        if (bytes < need) return fail(-1, "get: destination too small");
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < P.C; c++) {
                size_t o = transpose ? (size_t)c * rows + r : (size_t)r * P.C + c;
                if (as_double) reinterpret_cast<double*>(dst)[o] = (double)h[(size_t)r * Cs + c];
                else reinterpret_cast<T*>(dst)[o] = h[(size_t)r * Cs + c];
            }
        return 0;
    }

    int get(int what, int level, void* dst, size_t bytes) override {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaDeviceSynchronize());
        const int d = P.d;
        switch (what) {
        case TDA_G_SCALING: return get_soa(P.scaling, 1, dst, bytes, true, false);
        case TDA_G_ACCEPT_COUNTS: {
            if (bytes < (size_t)P.L * P.C * sizeof(long long)) return fail(-1, "get: destination too small");
            for (int l = 0; l < P.L; l++) {
                int r = get_soa(P.lv[l].n_acc, 1, reinterpret_cast<long long*>(dst) + (size_t)l * P.C,
                                (size_t)P.C * sizeof(long long), false, false);
                if (r) return r;
            }
            return 0;
        }
        case TDA_G_CURSORS: {
            if (bytes < (size_t)2 * P.C * sizeof(long long)) return fail(-1, "get: destination too small");
            long long* o = reinterpret_cast<long long*>(dst);
            for (int c = 0; c < P.C; c++) o[c] = P.t_base * d;
            if (P.zcur) { int r0 = get_soa(P.zcur, 1, o, (size_t)P.C * sizeof(long long), false, false); if (r0) return r0; }
            return get_soa(P.ucur, 1, o + P.C, (size_t)P.C * sizeof(long long), false, false);
        }
        case TDA_G_AM_SIGMA:
            if (!P.am_sigma) return fail(-1, "get: proposal is not AdaptiveMetropolis");
            return get_soa(P.am_sigma, d * d, dst, bytes, true, true);
        case TDA_G_AM_MU:
            if (!P.am_mu) return fail(-1, "get: proposal is not AdaptiveMetropolis");
            return get_soa(P.am_mu, d, dst, bytes, true, true);
        case TDA_G_THETA:
            if (level < 0 || level >= P.L) return fail(-1, "get: bad level");
            return get_soa(P.lv[level].theta, d, dst, bytes, true, true);
        case TDA_G_NRECORDS: {
            if (bytes < (size_t)P.L * sizeof(long long)) return fail(-1, "get: destination too small");
            for (int l = 0; l < P.L; l++) reinterpret_cast<long long*>(dst)[l] = P.rec[l];
            return 0;
        }
        case TDA_G_ERROR_FLAGS: {
            if (bytes < sizeof(long long)) return fail(-1, "get: destination too small");
            int f = 0;
            CUDA_TRY(cudaMemcpy(&f, P.error_flag, sizeof(int), cudaMemcpyDeviceToHost));
            reinterpret_cast<long long*>(dst)[0] = f;
            return 0;
        }
        case TDA_G_KERNEL: {
            if (bytes < sizeof(long long)) return fail(-1, "get: destination too small");
            reinterpret_cast<long long*>(dst)[0] = resolved_kernel();
            return 0;
        }
        case TDA_G_TC16_TIMELINE: {
            // diagnostic: first call arms the probe, later calls return [4][256] clock64 stamps
            if (bytes < 4 * 256 * sizeof(long long)) return fail(-1, "get: destination too small");
            int r = tc16.timeline(reinterpret_cast<long long*>(dst));
            return r ? fail(r, "tc16 timeline probe unavailable") : 0;
        }
        case TDA_G_MOMENTS: {
            if (bytes < (size_t)2 * d * P.C * sizeof(double)) return fail(-1, "get: destination too small");
            int r = get_soa(P.sum1, d, dst, (size_t)d * P.C * sizeof(double), true, false);
            if (r) return r;
            return get_soa(P.sum2, d, reinterpret_cast<double*>(dst) + (size_t)d * P.C, (size_t)d * P.C * sizeof(double), true, false);
        }
        default: return fail(-1, "get: unknown item");
        }
    }

    int set(int what, int level, const void* src, size_t bytes) override {
        CUDA_TRY(cudaSetDevice(device));
        if (what == TDA_G_SCALING) {
            if (bytes != (size_t)P.C * sizeof(double)) return fail(-1, "set: scaling needs n_chains float64 values");
            std::vector<R> h(Cs, (R)cfg.scaling);
            for (int c = 0; c < P.C; c++) h[c] = (R)reinterpret_cast<const double*>(src)[c];
            tc16.prepared = false;     // the pCN step is folded into its operators
            tc16_unfit = false;
            return put(P.scaling, h);
        }
        if (what == TDA_G_ZROUND) {
            if (bytes != sizeof(double)) return fail(-1, "set: z-round flag is one float64");
            z_round_user = *reinterpret_cast<const double*>(src) != 0.0;
            return 0;
        }
        return fail(-1, "set: unknown item");
    }

    int device_buffer(int buffer, int level, void** ptr, size_t* bytes) override {
        if (buffer == TDA_BUF_DREAM_ARCHIVE) {
            if (!P.archive) return fail(-1, "device_buffer: proposal has no archive");
            *ptr = P.archive;
            *bytes = (size_t)P.dream_cap * P.Cg * P.d * sizeof(R);
            return 0;
        }
        if (buffer == TDA_BUF_HIST_THETA) {
            if (level < 0 || level >= P.L || !P.lv[level].h_theta) return fail(-1, "device_buffer: no such history");
            *ptr = P.lv[level].h_theta;
            *bytes = (size_t)P.lv[level].hist_cap * P.d * Cs * sizeof(R);
            return 0;
        }
        return fail(-1, "device_buffer: unknown buffer");
    }

    int fill_streams(double* z, long long nz, double* u, long long nu) override {
        CUDA_TRY(cudaSetDevice(device));
        double *dz = nullptr, *du = nullptr;
        size_t tz = (size_t)P.C * nz, tu = (size_t)P.C * nu;
        CUDA_TRY(cudaMalloc(&dz, (tz ? tz : 1) * sizeof(double)));
        CUDA_TRY(cudaMalloc(&du, (tu ? tu : 1) * sizeof(double)));
        size_t n = tz > tu ? tz : tu;
        if (n) {
            tda::fill_streams_kernel<R><<<(unsigned)((n + 255) / 256), 256>>>(P.seed, P.chain_offset, P.C, dz, nz, du, nu, z_round_effective());
            g_launches++;
        }
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(z, dz, tz * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(u, du, tu * sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(dz);
        cudaFree(du);
        if (e != cudaSuccess) return fail(-2, std::string("fill_streams: ") + cudaGetErrorString(e));
        return 0;
    }
};

int validate(const tda_config* c) {
    if (!c) return fail(-1, "null config");
    if (c->abi_version != TDA_ABI_VERSION) return fail(-1, "ABI version mismatch");
    if (c->dtype != TDA_F32 && c->dtype != TDA_F64) return fail(-1, "dtype must be TDA_F32 or TDA_F64");
    if (c->n_levels < 1 || c->n_levels > TDA_MAX_LEVELS) return fail(-1, "n_levels out of range");
    if (c->d < 1 || c->d > TDA_MAX_D) return fail(-1, "d out of range (1..64)");
    if (c->n_chains < 1) return fail(-1, "n_chains must be positive");
    for (int l = 0; l + 1 < c->n_levels; l++)
        if (c->subchain[l] < 1) return fail(-1, "subchain lengths must be >= 1");
    if (c->prop_kind < TDA_PROP_RWMH || c->prop_kind > TDA_PROP_INDEP) return fail(-1, "unknown proposal kind");
    if ((c->prop_kind == TDA_PROP_MALA || c->prop_kind == TDA_PROP_DREAM || c->prop_kind == TDA_PROP_INDEP) && c->n_levels != 1)
        return fail(-1, "MALA / DREAM (shared archive) / IndependenceSampler are single-level proposals in this engine");
    if (c->prop_kind == TDA_PROP_INDEP && (c->adaptive || c->mtm_k)) return fail(-1, "IndependenceSampler is neither adaptive nor a MultipleTry kernel");
    if (tda::is_dream(c->prop_kind)) {
        if (c->dream_delta < 1 || c->dream_delta > tda::MAX_DELTA) return fail(-1, "DREAM delta out of range (1..8)");
        if (c->dream_M0 < 2 || c->dream_capacity < c->dream_M0) return fail(-1, "DREAM archive capacity too small");
        if (c->dream_nCR < 1 || c->dream_nCR > tda::MAX_NCR) return fail(-1, "DREAM nCR out of range (1..8)");
    }
    if (c->adaptive && c->period < 1) return fail(-1, "period must be >= 1");
    if (c->mtm_k < 0 || c->mtm_k == 1 || c->mtm_k > tda::MAX_MTM) return fail(-1, "mtm_k must be 0 or 2..16");
    if (c->mtm_k && c->prop_kind != TDA_PROP_RWMH && c->prop_kind != TDA_PROP_AM && c->prop_kind != TDA_PROP_PCN)
        return fail(-1, "MultipleTry needs an RWMH, AM or pCN kernel");
    if (c->mtm_k && (c->randomize_subchain || c->aem == 2)) return fail(-1, "MultipleTry cannot be combined with randomize_subchain / state-dependent AEM");
    if (c->aem < 0 || c->aem > 2) return fail(-1, "aem must be 0, 1 (state-independent) or 2 (state-dependent)");
    if (c->aem == 2 && c->n_levels != 2) return fail(-1, "the state-dependent error model is a two-level method");
    if (c->aem == 2 && c->prop_kind != TDA_PROP_RWMH && c->prop_kind != TDA_PROP_AM && c->prop_kind != TDA_PROP_PCN)
        return fail(-1, "the state-dependent error model needs a symmetric proposal or pCN");
    if (c->randomize_subchain && tda::is_dream(c->prop_kind))
        return fail(-1, "randomize_subchain needs a proposal with a fixed number of uniform draws per step");
    if (c->randomize_subchain && (c->n_levels != 2 || c->subchain[0] < 2))
        return fail(-1, "randomize_subchain needs two levels and a subchain length > 1");
    for (int l = 0; l < c->n_levels; l++) {
        const tda_level_config& lc = c->level[l];
        if (lc.m < 1) return fail(-1, "level has no outputs");
        if (lc.model_kind < TDA_MODEL_LINEAR || lc.model_kind > TDA_MODEL_POISSON1D) return fail(-1, "unknown model kind");
        if (lc.lik_kind < TDA_LIK_ISO || lc.lik_kind > TDA_LIK_ADAPTIVE) return fail(-1, "unknown likelihood kind");
        if (lc.model_kind == TDA_MODEL_ROSENBROCK && (c->d != 2 || lc.m != 1)) return fail(-1, "Rosenbrock needs d=2, m=1");
        if (lc.model_kind == TDA_MODEL_POISSON1D && lc.n_grid < 4) return fail(-1, "Poisson grid too small");
        if (c->aem && l + 1 < c->n_levels && lc.lik_kind != TDA_LIK_ADAPTIVE)
            return fail(-1, "adaptive error model needs adaptive likelihoods on the coarse levels");
        if (c->aem && lc.m != c->level[0].m) return fail(-1, "adaptive error model needs equal output sizes");
        if (c->prop_kind == TDA_PROP_MALA && lc.lik_kind >= TDA_LIK_DENSE) return fail(-1, "MALA supports isotropic / diagonal likelihoods");
        if (c->prop_kind == TDA_PROP_MALA && lc.model_kind == TDA_MODEL_POISSON1D) return fail(-1, "MALA needs a model with a gradient");
        if (c->prop_kind == TDA_PROP_MALA && lc.model_kind == TDA_MODEL_LINEAR && lc.m > TDA_MAX_D)
            return fail(-1, "MALA with a linear model supports m <= 64");
    }
    if (c->rng_mode == TDA_RNG_INJECTED && (c->stream_z_len < 1 || c->stream_u_len < 1)) return fail(-1, "injected streams need lengths");
    return 0;
}

}  // namespace

// ---- C ABI -----------------------------------------------------------------------------------
extern "C" {

int tda_abi_version(void) { return TDA_ABI_VERSION; }
const char* tda_last_error(void) { return g_err.c_str(); }
int64_t tda_launch_count(void) { return (int64_t)g_launches.load(); }

int tda_engine_create(const tda_config* cfg, int device, tda_engine** out) {
    if (!out) return fail(-1, "null output pointer");
    *out = nullptr;
    int r = validate(cfg);
    if (r) return r;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(-4, "no CUDA device available: the tinyda_b200 engine has no CPU path");
    if (device < 0 || device >= ndev) return fail(-1, "bad device index");
    tda_engine* e = nullptr;
    if (cfg->dtype == TDA_F32) {
        auto* t = new EngineT<float>();
        t->cfg = *cfg; t->device = device;
        r = t->create();
        e = t;
    } else {
        auto* t = new EngineT<double>();
        t->cfg = *cfg; t->device = device;
        r = t->create();
        e = t;
    }
    if (r) { delete e; return r; }
    *out = e;
    return 0;
}

int tda_engine_destroy(tda_engine* e) {
    delete e;
    return 0;
}

int tda_upload(tda_engine* e, int what, int level, const double* host, size_t count) {
    if (!e || !host) return fail(-1, "null argument");
    return e->upload(what, level, host, count);
}
int tda_engine_init(tda_engine* e, void* s) { return e ? e->init((cudaStream_t)s) : fail(-1, "null engine"); }
int tda_engine_run(tda_engine* e, int64_t it, void* s) { return e ? e->run(it, (cudaStream_t)s) : fail(-1, "null engine"); }
int tda_engine_burn(tda_engine* e, int64_t it, void* s) { return e ? e->run(it, (cudaStream_t)s, false) : fail(-1, "null engine"); }
int tda_compact_begin(tda_engine* e, int level, int64_t rec0, int64_t nrec, int first_is_full, int fields, int slot, void* s) {
    return e ? e->compact_begin(level, rec0, nrec, first_is_full, fields, slot, (cudaStream_t)s) : fail(-1, "null engine");
}
int tda_compact_rows(tda_engine* e, int slot, int64_t* n_rows) {
    if (!e || !n_rows) return fail(-1, "null argument");
    long long n = 0;
    int r = e->compact_rows(slot, &n);
    *n_rows = n;
    return r;
}
int tda_compact_fetch(tda_engine* e, int slot, int field, void* dst, size_t dst_bytes, size_t* bytes) {
    if (!e || !dst) return fail(-1, "null argument");
    return e->compact_fetch(slot, field, dst, dst_bytes, bytes);
}
int tda_compact_sync(tda_engine* e) { return e ? e->compact_sync() : fail(-1, "null engine"); }
int tda_peer_export(tda_engine* e, void* out, size_t bytes, size_t* needed) { return e ? e->peer_export(out, bytes, needed) : fail(-1, "null engine"); }
int tda_peer_import(tda_engine* e, int n_ranks, int my_rank, const void* handles, size_t bytes) {
    if (!e || !handles) return fail(-1, "null argument");
    return e->peer_import(n_ranks, my_rank, handles, bytes);
}
int tda_ess_sums(tda_engine* e, int level, int64_t rec0, int64_t nrec, int n_lag, double* sums, double* folded, void* s) {
    if (!e || !sums || !folded) return fail(-1, "null argument");
    return e->ess_sums(level, rec0, nrec, n_lag, sums, folded, (cudaStream_t)s);
}
int tda_host_alloc(size_t bytes, void** ptr) {
    if (!ptr) return fail(-1, "null argument");
    *ptr = nullptr;
    CUDA_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return 0;
}
int tda_pool_trim(void) {
    g_pool.trim();
    return 0;
}
int tda_host_free(void* ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return 0;
}
int tda_engine_sync(tda_engine* e, void* s) {
    if (!e) return fail(-1, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)s));
    return 0;
}
int tda_fetch(tda_engine* e, int level, int field, int64_t rec0, int64_t nrec, void* dst, size_t dst_bytes,
              size_t* bytes, void* s) {
    if (!e || !dst) return fail(-1, "null argument");
    return e->fetch(level, field, rec0, nrec, dst, dst_bytes, bytes, (cudaStream_t)s);
}
int tda_get(tda_engine* e, int what, int level, void* dst, size_t bytes) {
    if (!e || !dst) return fail(-1, "null argument");
    return e->get(what, level, dst, bytes);
}
int tda_set(tda_engine* e, int what, int level, const void* src, size_t bytes) {
    if (!e || !src) return fail(-1, "null argument");
    return e->set(what, level, src, bytes);
}
int tda_device_buffer(tda_engine* e, int buffer, int level, void** ptr, size_t* bytes) {
    if (!e || !ptr || !bytes) return fail(-1, "null argument");
    return e->device_buffer(buffer, level, ptr, bytes);
}
int tda_dream_slots(tda_engine* e, int64_t* slots) {
    if (!e || !slots) return fail(-1, "null argument");
    *slots = e->dream_slots;
    return 0;
}
int tda_fill_streams(tda_engine* e, double* z, int64_t nz, double* u, int64_t nu) {
    if (!e || !z || !u) return fail(-1, "null argument");
    return e->fill_streams(z, nz, u, nu);
}
int tda_tc_gemm_selftest(const float* A, const float* B, int N, float* D, int a_in_tmem, int split) {
    if (!A || !B || !D) return fail(-1, "null argument");
    std::string err;
    int r = tda::tc_gemm_selftest_host(A, B, N, D, a_in_tmem, split, err);
    g_launches++;
    if (r) return fail(r, err);
    return 0;
}
int tda_history_reset(tda_engine* e) { return e ? e->history_reset() : fail(-1, "null engine"); }
int tda_tc16_gemm_selftest(const float* A, const float* B, int N, float* D, int a_in_tmem) {
    if (!A || !B || !D) return fail(-1, "null argument");
    std::string err;
    int r = tda::tc16_gemm_selftest_host(A, B, N, D, a_in_tmem, err);
    if (r) return fail(r, err);
    g_launches++;
    return 0;
}

int tda_select_kernel(tda_engine* e, int which) { return e ? e->select_kernel(which) : fail(-1, "null engine"); }

int tda_state_size(tda_engine* e, size_t* bytes) {
    if (!e || !bytes) return fail(-1, "null argument");
    return e->state_io(nullptr, 0, bytes, 0);
}
int tda_state_save(tda_engine* e, void* host_dst, size_t dst_bytes) {
    if (!e || !host_dst) return fail(-1, "null argument");
    return e->state_io(host_dst, dst_bytes, nullptr, 0);
}
int tda_state_load(tda_engine* e, const void* host_src, size_t src_bytes) {
    if (!e || !host_src) return fail(-1, "null argument");
    return e->state_io(const_cast<void*>(host_src), src_bytes, nullptr, 1);
}

}  // extern "C"
