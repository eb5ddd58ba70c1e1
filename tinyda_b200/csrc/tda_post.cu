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

// One warp = 32 consecutive chains x one segment of the records (lane = chain); warps never wait for each other.
// Per record the accept bytes of the group are one coalesced load (the next record's is already in flight).  A record
// with few accepted chains (<= GW_DIRECT, the usual case at MCMC acceptance rates) is gathered directly: lane k reads
// elements k, k + 32, ... of the accepted chain's row (one 32-byte sector each in the chain-fastest history) and
// the row leaves with coalesced stores.  A record with many accepted chains stages the [W][32] slab through the
// warp's shared-memory tile with 128-byte loads instead (one read of the slab whatever the number of rows).
// Segments other than the first start by counting the group's accepted records before their first one.
// prior / like (optional): the records' log-densities travel with the rows in the same pass
constexpr int GW_DIRECT = 4;
template <typename R>
__global__ void __launch_bounds__(256) gather_kernel(const R* __restrict__ src, int W, const uint8_t* __restrict__ acc, long long nrec, int C,
                                                     int Cs, int force_first, const long long* __restrict__ offsets, R* __restrict__ dst,
                                                     const R* __restrict__ prior, const R* __restrict__ like, R* __restrict__ dst_prior,
                                                     R* __restrict__ dst_like, int n_groups, int n_seg, long long seg_len) {
    extern __shared__ __align__(16) unsigned char gw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    R(*tile)[33] = reinterpret_cast<R(*)[33]>(gw_smem + (size_t)warp * 64 * 33 * sizeof(R));
    const long long job = (long long)blockIdx.x * wpb + warp;
    if (job >= (long long)n_groups * n_seg) return;
    // consecutive warps take consecutive chain groups of one segment: neighbouring lines of the same records
    const int seg = (int)(job / n_groups), group = (int)(job % n_groups);
    const int c0 = group * 32;
    const bool live = c0 + lane < C;
    const long long r0 = seg * seg_len, r1 = (r0 + seg_len < nrec) ? r0 + seg_len : nrec;
    long long rowpos = live ? offsets[c0 + lane] : 0;
    {
        int cnt = 0;
        long long r = 0;
        for (; r + 8 <= r0; r += 8) {
            uint8_t f[8];
#pragma unroll
            for (int u = 0; u < 8; u++) f[u] = acc[(size_t)(r + u) * Cs + c0 + lane];
#pragma unroll
            for (int u = 0; u < 8; u++) cnt += (f[u] != 0 || (force_first && r + u == 0)) ? 1 : 0;
        }
        for (; r < r0; r++) cnt += (acc[(size_t)r * Cs + c0 + lane] != 0 || (force_first && r == 0)) ? 1 : 0;
        if (live) rowpos += cnt;
    }
    uint8_t a_next = (r0 < r1) ? acc[(size_t)r0 * Cs + c0 + lane] : 0;
    for (long long r = r0; r < r1; r++) {
        const bool a = live && (a_next != 0 || (force_first && r == 0));
        if (r + 1 < r1) a_next = acc[(size_t)(r + 1) * Cs + c0 + lane];
        const unsigned mask = __ballot_sync(0xffffffffu, a);
        if (!mask) continue;
        if (prior && a) {
            dst_prior[rowpos] = prior[(size_t)r * Cs + c0 + lane];
            dst_like[rowpos] = like[(size_t)r * Cs + c0 + lane];
        }
        const R* rec = src + (size_t)r * W * Cs + c0;
        if (__popc(mask) <= GW_DIRECT) {
            // all rows' loads of a 64-column chunk in flight before the first store
            int cs[GW_DIRECT];
            long long ps[GW_DIRECT];
            unsigned m = mask;
#pragma unroll
            for (int i = 0; i < GW_DIRECT; i++) {
                cs[i] = m ? __ffs(m) - 1 : 0;
                ps[i] = __shfl_sync(0xffffffffu, rowpos, cs[i]);
                if (!m) ps[i] = -1;
                m &= m - 1;
            }
            for (int k0 = 0; k0 < W; k0 += 64) {
                R v[GW_DIRECT][2];
#pragma unroll
                for (int i = 0; i < GW_DIRECT; i++)
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int k = k0 + lane + 32 * u;
                        if (ps[i] >= 0 && k < W) v[i][u] = rec[(size_t)k * Cs + cs[i]];
                    }
#pragma unroll
                for (int i = 0; i < GW_DIRECT; i++)
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int k = k0 + lane + 32 * u;
                        if (ps[i] >= 0 && k < W) dst[(size_t)ps[i] * W + k] = v[i][u];
                    }
            }
        } else {
            for (int k0 = 0; k0 < W; k0 += 64) {
                const int kw = W - k0 < 64 ? W - k0 : 64;
                int k = 0;
                for (; k + 32 <= kw; k += 32) {            // 32 independent 128-byte loads per lane in flight
                    R v[32];
#pragma unroll
                    for (int u = 0; u < 32; u++) v[u] = rec[(size_t)(k0 + k + u) * Cs + lane];
#pragma unroll
                    for (int u = 0; u < 32; u++) tile[k + u][lane] = v[u];
                }
                for (; k + 8 <= kw; k += 8) {
                    R v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) v[u] = rec[(size_t)(k0 + k + u) * Cs + lane];
#pragma unroll
                    for (int u = 0; u < 8; u++) tile[k + u][lane] = v[u];
                }
                for (; k < kw; k++) tile[k][lane] = rec[(size_t)(k0 + k) * Cs + lane];
                __syncwarp();
                unsigned m = mask;
                while (m) {
                    const int c = __ffs(m) - 1;
                    m &= m - 1;
                    const long long pos = __shfl_sync(0xffffffffu, rowpos, c);
                    R* o = dst + (size_t)pos * W + k0;
                    for (int kk = lane; kk < kw; kk += 32) o[kk] = tile[kk][c];
                }
                __syncwarp();
            }
        }
        if (a) rowpos += 1;
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

template <typename R>
static cudaError_t gather_launch(const void* src, int W, const uint8_t* acc, long long nrec, int C, int Cs, int force_first,
                                 const long long* offsets, void* dst, cudaStream_t st, const void* prior, const void* like, void* dst_prior,
                                 void* dst_like) {
    // 64 KB of tiles per block (8 warps in float, 4 in double), three blocks per SM
    const int wpb = sizeof(R) == 4 ? 8 : 4;
    const size_t smem = (size_t)wpb * 64 * 33 * sizeof(R);
    cudaError_t ea = cudaFuncSetAttribute(gather_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea != cudaSuccess) return ea;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_groups = (C + 31) / 32;
    // enough warps for about two rounds of the machine's resident warps, at least 8 records per segment
    const long long slots = (long long)sms * 3 * wpb;
    long long n_seg = (2 * slots + n_groups - 1) / n_groups;
    if (n_seg > nrec / 8) n_seg = nrec / 8;
    if (n_seg > 64) n_seg = 64;
    if (n_seg < 1) n_seg = 1;
    const long long seg_len = (nrec + n_seg - 1) / n_seg;
    n_seg = (nrec + seg_len - 1) / seg_len;
    const long long jobs = (long long)n_groups * n_seg;
    const unsigned grid = (unsigned)((jobs + wpb - 1) / wpb);
    gather_kernel<R><<<grid, wpb * 32, smem, st>>>((const R*)src, W, acc, nrec, C, Cs, force_first, offsets, (R*)dst, (const R*)prior,
                                                   (const R*)like, (R*)dst_prior, (R*)dst_like, n_groups, (int)n_seg, seg_len);
    return cudaGetLastError();
}

int compact_gather(const void* src, int esz, int W, const uint8_t* acc, long long nrec, int C, int Cs, int force_first,
                   const long long* offsets, void* dst, cudaStream_t st, const void* prior, const void* like, void* dst_prior,
                   void* dst_like) {
    if (nrec <= 0 || C <= 0) return 0;
    const cudaError_t e = (esz == 4) ? gather_launch<float>(src, W, acc, nrec, C, Cs, force_first, offsets, dst, st, prior, like, dst_prior, dst_like)
                                     : gather_launch<double>(src, W, acc, nrec, C, Cs, force_first, offsets, dst, st, prior, like, dst_prior, dst_like);
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
