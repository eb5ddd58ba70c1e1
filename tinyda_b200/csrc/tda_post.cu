// Device-side post-processing of the Link history: compaction to accepted records (see tda_post.h).
// HBM-bound byte shuffling: coalesced accept-byte reads (lane = chain); the row gather stages 32-chain slabs
// through shared memory so that both its reads (chain-fastest history) and its writes (row-major rows) are
// whole 128-byte lines.
#include <string>

#include "tda_post.h"

namespace tda {
namespace post {

namespace {
thread_local std::string g_perr;
int pfail(const char* what, cudaError_t e) {
    g_perr = std::string(what) + ": " + cudaGetErrorString(e);
    return -2;
}

constexpr int CB = 256;   // chains per block of the counting pass

// pass 1: accepted records per chain, block-level exclusive scan, block totals
__global__ void __launch_bounds__(CB) count_kernel(const uint8_t* __restrict__ acc, long long nrec, int C, int Cs, int force_first,
                                                   long long* __restrict__ offsets, long long* __restrict__ block_tot) {
    __shared__ long long s_w[CB / 32];
    const int c = blockIdx.x * CB + threadIdx.x;
    long long n = 0;
    if (c < C) {
        const uint8_t* a = acc + c;
        for (long long r = 0; r < nrec; r++) n += (a[(size_t)r * Cs] != 0 || (force_first && r == 0)) ? 1 : 0;
    }
    // inclusive warp scan, then across the 8 warps
    long long v = n;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) s_w[w] = v;
    __syncthreads();
    long long base = 0;
    for (int i = 0; i < w; i++) base += s_w[i];
    if (c < C) offsets[c] = base + v - n;            // exclusive within the block
    if (threadIdx.x == CB - 1) block_tot[blockIdx.x] = base + v;
}

// pass 2: one block turns the block totals into exclusive block offsets (in place), total at [nblocks]
__global__ void __launch_bounds__(1024) scan_blocks_kernel(long long* __restrict__ block_tot, int nblocks) {
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const long long n = i < nblocks ? block_tot[i] : 0;
        long long v = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) s_w[w] = v;
        __syncthreads();
        long long base = s_carry;
        for (int k = 0; k < w; k++) base += s_w[k];
        if (i < nblocks) block_tot[i] = base + v - n;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = base + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_tot[nblocks] = s_carry;
}

// pass 3: add the block offsets; offsets[C] = total
__global__ void __launch_bounds__(CB) add_blocks_kernel(long long* __restrict__ offsets, const long long* __restrict__ block_off, int C, int nblocks) {
    const int c = blockIdx.x * CB + threadIdx.x;
    if (c < C) offsets[c] += block_off[blockIdx.x];
    if (c == 0) offsets[C] = block_off[nblocks];
}

// One block = 32 consecutive chains.  Per record the [W][32] slab of the chain-fastest history is staged
// through shared memory with coalesced loads (lane = chain, one 128-byte line per row) and the rows of the
// accepted chains leave with coalesced stores: HBM traffic = one read of the dense records + one write of
// the accepted rows, whatever the acceptance rate.
// prior / like (optional): the records' log-densities travel with the rows in the same pass
template <typename R>
__global__ void __launch_bounds__(256) gather_kernel(const R* __restrict__ src, int W, const uint8_t* __restrict__ acc, long long nrec, int C,
                                                     int Cs, int force_first, const long long* __restrict__ offsets, R* __restrict__ dst,
                                                     const R* __restrict__ prior, const R* __restrict__ like, R* __restrict__ dst_prior,
                                                     R* __restrict__ dst_like) {
    __shared__ R tile[64][33];
    __shared__ R tile_s[2][32];
    __shared__ long long rowpos[32];
    __shared__ unsigned s_mask;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32;
    if (threadIdx.x < 32) rowpos[threadIdx.x] = (c0 + threadIdx.x < C) ? offsets[c0 + threadIdx.x] : 0;
    __syncthreads();
    for (long long r = 0; r < nrec; r++) {
        if (warp == 0) {
            const bool a = (c0 + lane < C) && (acc[(size_t)r * Cs + c0 + lane] != 0 || (force_first && r == 0));
            const unsigned m = __ballot_sync(0xffffffffu, a);
            if (lane == 0) s_mask = m;
        }
        __syncthreads();
        const unsigned mask = s_mask;
        if (mask && prior) {
            if (warp == 1) tile_s[0][lane] = prior[(size_t)r * Cs + c0 + lane];
            if (warp == 2) tile_s[1][lane] = like[(size_t)r * Cs + c0 + lane];
        }
        if (mask) {
            for (int k0 = 0; k0 < W; k0 += 64) {
                const int kw = W - k0 < 64 ? W - k0 : 64;
                for (int k = warp; k < kw; k += 8) tile[k][lane] = src[((size_t)r * W + k0 + k) * Cs + c0 + lane];
                __syncthreads();
                unsigned m = mask;
                for (int jj = 0; m; jj++) {
                    const int c = __ffs(m) - 1;
                    m &= m - 1;
                    if ((jj & 7) == warp) {
                        R* o = dst + (size_t)rowpos[c] * W + k0;
                        for (int k = lane; k < kw; k += 32) o[k] = tile[k][c];
                        if (prior && k0 == 0 && lane < 2) (lane ? dst_like : dst_prior)[rowpos[c]] = tile_s[lane][c];
                    }
                }
                __syncthreads();
            }
        }
        if (threadIdx.x < 32 && ((mask >> threadIdx.x) & 1u)) rowpos[threadIdx.x] += 1;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) flags_kernel(const uint8_t* __restrict__ acc, long long nrec, int C, int Cs, int force_first,
                                                    uint8_t* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t total = (size_t)nrec * Cs;
    if (i < total) {
        const size_t r = i / Cs;
        dst[i] = (acc[i] != 0 || (force_first && r == 0)) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) max_abs_kernel(const float* __restrict__ x, size_t n, unsigned int* __restrict__ out) {
    float m = 0.0f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float v = fabsf(x[i]);
        m = (v <= 3.0e38f) ? fmaxf(m, v) : 3.0e38f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

}  // namespace

const char* last_error() { return g_perr.c_str(); }

int max_abs_f32(const float* x, size_t n, unsigned int* out_bits, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out_bits, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return pfail("max_abs", e);
    const unsigned grid = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    max_abs_kernel<<<grid ? grid : 1, 256, 0, st>>>(x, n, out_bits);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : pfail("max_abs", e);
}

int compact_offsets(const uint8_t* acc, long long nrec, int C, int Cs, int force_first, long long* offsets, long long* scratch,
                    cudaStream_t st) {
    const int nblocks = (C + CB - 1) / CB;
    count_kernel<<<nblocks, CB, 0, st>>>(acc, nrec, C, Cs, force_first, offsets, scratch);
    scan_blocks_kernel<<<1, 1024, 0, st>>>(scratch, nblocks);
    add_blocks_kernel<<<nblocks, CB, 0, st>>>(offsets, scratch, C, nblocks);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : pfail("compact_offsets", e);
}

int compact_gather(const void* src, int esz, int W, const uint8_t* acc, long long nrec, int C, int Cs, int force_first,
                   const long long* offsets, void* dst, cudaStream_t st, const void* prior, const void* like, void* dst_prior,
                   void* dst_like) {
    const unsigned grid = (unsigned)((C + 31) / 32);
    if (esz == 4)
        gather_kernel<float><<<grid, 256, 0, st>>>((const float*)src, W, acc, nrec, C, Cs, force_first, offsets, (float*)dst,
                                                   (const float*)prior, (const float*)like, (float*)dst_prior, (float*)dst_like);
    else
        gather_kernel<double><<<grid, 256, 0, st>>>((const double*)src, W, acc, nrec, C, Cs, force_first, offsets, (double*)dst,
                                                    (const double*)prior, (const double*)like, (double*)dst_prior, (double*)dst_like);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : pfail("compact_gather", e);
}

int compact_flags(const uint8_t* acc, long long nrec, int C, int Cs, int force_first, uint8_t* dst, cudaStream_t st) {
    const size_t total = (size_t)nrec * Cs;
    if (!total) return 0;
    flags_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(acc, nrec, C, Cs, force_first, dst);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : pfail("compact_flags", e);
}

}  // namespace post
}  // namespace tda
