// Rank-normalised split-chain diagnostics on the device (see tda_post.h): what ArviZ computes on the
// reference's to_inference_data output (diagnostics.py:6-69; Vehtari et al. 2021), without the parameter
// history ever crossing PCIe.  Per parameter: pack the split chains -> radix sort (cub) -> average ranks and
// normal scores -> per-chain means -> all-lag autocovariance sums (rolling 16-lag register window, two
// coalesced loads per 16 FMAs) -> a handful of float64 sums for the host / the all-reduce.
#include <cub/device/device_radix_sort.cuh>
#include <string>

#include "tda_post.h"

namespace tda {
namespace post {

struct EssWorkspace {
    long long n_draws = 0, N = 0;
    int C = 0, Cs = 0, n_lag = 0, n_half = 0, S = 0;
    float *keys_in = nullptr, *keys_out = nullptr, *z = nullptr, *mean = nullptr;
    unsigned *idx_in = nullptr, *idx_out = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    double* sums = nullptr;      // [n_lag + 4]
};

namespace {

thread_local std::string g_eerr;

// split chain s = half * C + c, draw t: record r = t (first half) or n_draws - n_half + t (second half)
template <typename R>
__global__ void __launch_bounds__(256) ess_pack_kernel(const R* __restrict__ hist, long long stride_t, long long n_draws, int n_half, int C,
                                                       int S, float* __restrict__ keys, unsigned* __restrict__ idx, int fold, const float* __restrict__ median2) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t N = (size_t)n_half * S;
    if (i >= N) return;
    const int t = (int)(i / S), s = (int)(i - (size_t)t * S);
    const int half = s / C, c = s - half * C;
    const long long r = half ? n_draws - n_half + t : t;
    float x = (float)hist[(size_t)r * stride_t + c];
    if (fold) x = fabsf(x - 0.5f * (median2[0] + median2[1]));
    keys[i] = x;
    idx[i] = (unsigned)i;
}

// sorted position -> average rank of its run of equal keys -> normal score, scattered back to (t, s)
__global__ void __launch_bounds__(256) ess_rank_kernel(const float* __restrict__ keys, const unsigned* __restrict__ idx, size_t N, float* __restrict__ z) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const float k = keys[i];
    if (i > 0 && keys[i - 1] == k) return;                 // not the start of a run
    size_t j = i + 1;
    while (j < N && keys[j] == k) j++;
    const double avg_rank = 0.5 * ((double)(i + 1) + (double)j);       // ranks i+1 .. j
    const float zv = (float)normcdfinv((avg_rank - 0.375) / ((double)N + 0.25));
    for (size_t m = i; m < j; m++) z[idx[m]] = zv;
}

// per split chain: mean over its draws, centred in place; sums of the means and of their squares
__global__ void __launch_bounds__(256) ess_mean_kernel(float* __restrict__ z, int n_half, int S, double* __restrict__ sums, int n_lag) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    double m = 0.0;
    if (s < S) {
        for (int t = 0; t < n_half; t++) m += (double)z[(size_t)t * S + s];
        m /= (double)n_half;
        const float mf = (float)m;
        for (int t = 0; t < n_half; t++) z[(size_t)t * S + s] -= mf;
    }
    double a = (s < S) ? m : 0.0, b = (s < S) ? m * m : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sums + n_lag, a);
        atomicAdd(sums + n_lag + 1, b);
    }
}

// thread = (split chain, block of 16 lags): acc[j] = sum_t x[t] x[t + tau0 + j] with a rolling window
__global__ void __launch_bounds__(128) ess_acov_kernel(const float* __restrict__ z, int n_half, int S, int n_lag, double* __restrict__ sums) {
    const int s = blockIdx.x * 128 + threadIdx.x;
    const int tau0 = blockIdx.y * 16;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.0f;
    if (s < S) {
        const float* x = z + s;
        float win[16];
#pragma unroll
        for (int j = 0; j < 16; j++) win[j] = (tau0 + j < n_half) ? x[(size_t)(tau0 + j) * S] : 0.0f;
        const int T = n_half - tau0;
        for (int t = 0; t < T; t += 16) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const float xt = (t + u < n_half) ? x[(size_t)(t + u) * S] : 0.0f;
#pragma unroll
                for (int j = 0; j < 16; j++) acc[j] = fmaf(xt, win[(u + j) & 15], acc[j]);
                const int nx = t + u + tau0 + 16;
                win[u] = (nx < n_half) ? x[(size_t)nx * S] : 0.0f;
            }
        }
    }
    const double inv_n = 1.0 / (double)n_half;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        float v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && tau0 + j < n_lag) atomicAdd(sums + tau0 + j, (double)v * inv_n);
    }
}

int efail(const char* what, cudaError_t e) {
    g_eerr = std::string(what) + ": " + cudaGetErrorString(e);
    return -2;
}

}  // namespace

const char* ess_last_error() { return g_eerr.c_str(); }

EssWorkspace* ess_workspace_create(long long n_draws, int C, int Cs, int n_lag) {
    EssWorkspace* w = new EssWorkspace();
    w->n_draws = n_draws; w->C = C; w->Cs = Cs;
    w->n_half = (int)(n_draws / 2);
    w->S = 2 * C;
    w->n_lag = n_lag < 1 ? 1 : (n_lag > w->n_half ? w->n_half : n_lag);
    w->N = (long long)w->n_half * w->S;
    const size_t N = (size_t)w->N;
    cudaError_t e = cudaSuccess;
    if (w->n_half < 2 || N >= 0xFFFFFFFFull) { g_eerr = "ess: need at least 4 draws and fewer than 2^32 values per parameter"; delete w; return nullptr; }
    if (e == cudaSuccess) e = cudaMalloc(&w->keys_in, N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->keys_out, N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->idx_in, N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->idx_out, N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->z, N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->mean, (size_t)w->S * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->sums, ((size_t)w->n_lag + 4) * 8);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, w->cub_bytes, w->keys_in, w->keys_out, w->idx_in, w->idx_out, (int)N);
    if (e == cudaSuccess) e = cudaMalloc(&w->cub_tmp, w->cub_bytes ? w->cub_bytes : 1);
    if (e != cudaSuccess) {
        efail("ess workspace", e);
        ess_workspace_destroy(w);
        return nullptr;
    }
    return w;
}

void ess_workspace_destroy(EssWorkspace* w) {
    if (!w) return;
    cudaFree(w->keys_in); cudaFree(w->keys_out); cudaFree(w->idx_in); cudaFree(w->idx_out);
    cudaFree(w->z); cudaFree(w->mean); cudaFree(w->sums); cudaFree(w->cub_tmp);
    delete w;
}

int ess_sums(EssWorkspace* w, const void* hist, int esz, long long stride_t, double* sums_host, double* folded_host, cudaStream_t st) {
    const size_t N = (size_t)w->N;
    const unsigned gN = (unsigned)((N + 255) / 256);
    cudaError_t e = cudaSuccess;
    for (int fold = 0; fold < 2; fold++) {
        const int n_lag = fold ? 1 : w->n_lag;
        // after the first sort keys_out is sorted: its two middle elements give the median for the folded pass
        const float* med = w->keys_out + (N / 2 - 1);
        if (fold) {
            // the median must survive the repacking of keys_out's sibling buffers: copy it aside (into `mean`, free now)
            e = cudaMemcpyAsync(w->mean, med, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return efail("ess median", e);
        }
        if (esz == 4)
            ess_pack_kernel<float><<<gN, 256, 0, st>>>((const float*)hist, stride_t, w->n_draws, w->n_half, w->C, w->S, w->keys_in, w->idx_in, fold, w->mean);
        else
            ess_pack_kernel<double><<<gN, 256, 0, st>>>((const double*)hist, stride_t, w->n_draws, w->n_half, w->C, w->S, w->keys_in, w->idx_in, fold, w->mean);
        e = cub::DeviceRadixSort::SortPairs(w->cub_tmp, w->cub_bytes, w->keys_in, w->keys_out, w->idx_in, w->idx_out, (int)N, 0, 32, st);
        if (e != cudaSuccess) return efail("ess sort", e);
        ess_rank_kernel<<<gN, 256, 0, st>>>(w->keys_out, w->idx_out, N, w->z);
        e = cudaMemsetAsync(w->sums, 0, ((size_t)w->n_lag + 4) * 8, st);
        if (e != cudaSuccess) return efail("ess memset", e);
        ess_mean_kernel<<<(unsigned)((w->S + 255) / 256), 256, 0, st>>>(w->z, w->n_half, w->S, w->sums, n_lag);
        dim3 grid((unsigned)((w->S + 127) / 128), (unsigned)((n_lag + 15) / 16));
        ess_acov_kernel<<<grid, 128, 0, st>>>(w->z, w->n_half, w->S, n_lag, w->sums);
        e = cudaGetLastError();
        if (e != cudaSuccess) return efail("ess kernels", e);
        double* dst = fold ? folded_host : sums_host;
        e = cudaMemcpyAsync(dst, w->sums, ((size_t)n_lag + 2) * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return efail("ess copy", e);
        dst[n_lag + 2] = (double)w->S;
        if (!fold) dst[n_lag + 3] = (double)w->n_half;
    }
    return 0;
}

}  // namespace post
}  // namespace tda
