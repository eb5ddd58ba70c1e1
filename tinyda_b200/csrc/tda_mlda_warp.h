// Host interface of the warp-per-chain MH / DA / MLDA kernel (1-D Poisson model or linear operators with at most 31
// outputs, state-independent adaptive error model; tda_mlda_warp.cu, its own translation unit) -- BASELINE cfg4 and
// the small problems of the reference's notebooks.  See the header comment there.
#pragma once
#include <cuda_runtime.h>

#include "tda_common.cuh"

namespace tda {

constexpr int MW_MAXW = 16;      // warps (= chains in flight) per CTA at most
constexpr int MW_MATW = 1024;    // elements of one warp-major matrix image (31 x 32 / 31 x 33, padded)

struct MwParams {
    void* sigw;    // [chain][MAXL][MW_MATW] bias covariances, element (i, j) at j * 32 + i
    void* liw;     // [chain][MAXL][MW_MATW] inverse Cholesky factors, element (i, j) at j * 33 + i
    const void* phiw;        // Phi images of all levels, [k][t][32] each (element = Phi[k][lane * stride + t])
    int phi_off[TDA_MAX_LEVELS];     // element offset of a level's image in phiw
    int phi_elems[TDA_MAX_LEVELS];
    int phi_smem[TDA_MAX_LEVELS];    // element offset of the level's shared-memory copy in the CTA block, -1: read from phiw
    int cta_elems;                   // CTA-shared elements in front of the per-warp blocks
    int warps;
};

bool mlda_warp_eligible(const tda_config& c);
// elements (of the engine's real type) of one image buffer for n_chains chains
size_t mlda_warp_image_elems(int n_chains);
// elements of the Phi image buffer
size_t mlda_warp_phi_elems(const tda_config& c);
// kernel launches of one mlda_warp_run call (layout conversions included)
int mlda_warp_launches(int n_levels, int aem);

// Advances every chain by P.iterations finest-level iterations.  sigw / liw: scratch images (mlda_warp_image_elems
// elements each); the error-model matrices of the lock-step layout are converted into them before the launch and
// back after it, so the engine's buffers stay the single source of truth between launches.
template <typename R>
int mlda_warp_run(Params<R>& P, void* sigw, void* liw, void* phiw, int sm_count, cudaStream_t st);

const char* mlda_warp_last_error();

}  // namespace tda
