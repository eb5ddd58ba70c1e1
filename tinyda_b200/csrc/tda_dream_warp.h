// Host interface of the warp-per-chain DREAM(Z) / DREAM kernel (tda_dream_warp.cu, its own translation unit):
// single level, linear forward operator, isotropic / diagonal Gaussian likelihood, d <= 32, non-adaptive
// crossover -- BASELINE cfg5.  See the header comment of tda_dream_warp.cu.
#pragma once
#include <cuda_runtime.h>

#include "tda_common.cuh"

namespace tda {

inline bool dream_warp_eligible(const tda_config& c) {
    if (c.n_levels != 1 || !is_dream(c.prop_kind) || c.adaptive || c.mtm_k || c.aem) return false;
    if (c.d > 32) return false;
    const tda_level_config& lc = c.level[0];
    if (lc.model_kind != TDA_MODEL_LINEAR) return false;
    return lc.lik_kind == TDA_LIK_ISO || lc.lik_kind == TDA_LIK_DIAG;
}

// number of CTAs of a launch in which every chain has its register-resident slot and all CTAs are co-resident
// (the shared-archive variant ends each step in a grid-wide barrier); 0 = the job does not fit this kernel
template <typename R>
int dream_warp_grid(const Params<R>& P, int sm_count, int* occ);      // *occ: kernel variant (resident CTAs per SM, 2 or 3)

template <typename R>
int dream_warp_launch(Params<R>& P, int grid, int occ, cudaStream_t st);

const char* dream_warp_last_error();

}  // namespace tda
