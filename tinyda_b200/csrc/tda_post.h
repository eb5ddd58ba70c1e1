// Device-side post-processing of the Link history (separate translation unit, see Makefile):
//   * compaction of the finest level's records to the ACCEPTED ones -- a rejected step re-appends the
//     same Link object in the reference (chain.py:116, :434; proposal.py:1601), so only the accept
//     byte of every record and the fields of the accepted records have to cross PCIe;
//   * Link.qoi / Link.model_output of linear models rebuilt from stored parameters (tda_engine.cu);
//   * rank-normalised split R-hat / bulk ESS of the finest level's parameter history
//     (what ArviZ computes on the reference's to_inference_data output, diagnostics.py:6-69).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace tda {
namespace post {

// ---- compaction ---------------------------------------------------------------------------------
// acc: [nrec][Cs] accept bytes of the records to compact.  force_first: record 0 is a full row whatever
// its flag says (the initial Link).  Writes counts[c] (accepted records of chain c) and, after the scan,
// offsets[C + 1] (chain-major exclusive prefix: the rows of chain c are offsets[c] .. offsets[c+1]).
// `scratch` holds (C + 255) / 256 + 1 int64 values.  total_out (device) receives offsets[C].
int compact_offsets(const uint8_t* acc, long long nrec, int C, int Cs, int force_first, long long* offsets,
                    long long* scratch, cudaStream_t st);

// dst[offsets[c] + j][0..W) = src[r_j][0..W)[c] for the j-th accepted record r_j of chain c
// (src: [nrec][W][Cs], chain fastest; dst row-major [n_rows][W]); esz = 4 (float) or 8 (double)
// prior / like != nullptr: the records' log-densities ([nrec][Cs]) are gathered in the same pass into dst_prior / dst_like
int compact_gather(const void* src, int esz, int W, const uint8_t* acc, long long nrec, int C, int Cs, int force_first,
                   const long long* offsets, void* dst, cudaStream_t st, const void* prior = nullptr, const void* like = nullptr,
                   void* dst_prior = nullptr, void* dst_like = nullptr);

// accept bytes with the forced first record made explicit: dst[r][c] (dense [nrec][Cs])
int compact_flags(const uint8_t* acc, long long nrec, int C, int Cs, int force_first, uint8_t* dst, cudaStream_t st);

// ---- diagnostics on the device ----------------------------------------------------------------------
// Rank-normalised split-chain diagnostics (Vehtari et al. 2021 -- what ArviZ computes on the reference's
// to_inference_data output, diagnostics.py:6-69) of ONE parameter's history x[t][c] = hist[t * stride_t + c]
// (t < n_draws, c < C, engine dtype esz).  Every chain is split in halves (n_half = n_draws / 2 draws each,
// the middle draw of an odd run is dropped), all 2 C n_half values are ranked on the device (radix sort,
// average ranks for ties -- a rejected step repeats its value), turned into normal scores, and the sums the
// multi-chain estimators need are accumulated:
//   sums[0 .. n_lag)  sum over the 2C split chains of the biased autocovariance at lag t
//   sums[n_lag + 0]   sum of the split-chain means      sums[n_lag + 1]  sum of their squares
//   sums[n_lag + 2]   number of split chains            sums[n_lag + 3]  n_half
// `folded` (4 values: lag-0 autocovariance sum, mean sum, mean-square sum, split chains) is the same for the
// normal scores of |x - median| (the second half of rank-normalised split R-hat).  Both are plain sums over
// chains, so ranks of a multi-GPU job all-reduce them.  Workspace (about 20 bytes per value) is allocated
// once per EssWorkspace and reused for every parameter.
struct EssWorkspace;
EssWorkspace* ess_workspace_create(long long n_draws, int C, int Cs, int n_lag);
void ess_workspace_destroy(EssWorkspace* w);
int ess_sums(EssWorkspace* w, const void* hist, int esz, long long stride_t, double* sums_host, double* folded_host, cudaStream_t st);

const char* ess_last_error();

// max |x[i]| over n floats -> *out_bits (the bits of a non-negative float; NaN / inf count as 3e38); out_bits is zeroed here
int max_abs_f32(const float* x, size_t n, unsigned int* out_bits, cudaStream_t st);

const char* last_error();

}  // namespace post
}  // namespace tda
