// Host interface of the "tcr" tcgen05 Delayed-Acceptance kernel (tda_da_tcr.cu, its own translation unit):
// two-level DA, pCN with per-chain step sizes (fixed or adaptively scaled), linear forward operators,
// Gaussian likelihoods, float32 engine -- BASELINE cfg2.  See the header comment of tda_da_tcr.cu.
#pragma once
#include <cuda_runtime.h>
#include <string>

#include "tda_common.cuh"

namespace tda {

struct DaTcrImpl;

struct DaTcrState {
    std::string err;
    bool prepared = false;
    bool w_valid = false;     // the whitened state buffer mirrors the finest level's theta
    DaTcrImpl* impl = nullptr;

    bool eligible(const tda_config& c, const Params<float>& P) const;
    // returns 1 when the problem does not fit (caller falls back to another kernel), < 0 on CUDA errors
    int prepare(const Params<float>& P, const tda_config& c);
    // prepare() if needed, then (when the whitened state buffer is stale) w = theta T^-1 of the finest level's
    // current state with the range check of the fp16 operand images; 1 = does not fit, < 0 = CUDA error
    int ready(const Params<float>& P, const tda_config& c, cudaStream_t st);
    int run(Params<float>& P, const tda_config& c, long long iterations, int sm_count, cudaStream_t st);
    void invalidate() { w_valid = false; }
    void destroy();
};

}  // namespace tda
