// Tensor-core (tcgen05 / TMEM) fast path for the two-level Delayed-Acceptance hot path with a
// linear forward operator and isotropic likelihoods (BASELINE cfg2).  Placeholder state
// object until the kernel lands: never eligible, so the generic kernel runs.
#pragma once
#include <string>
#include "tda_common.cuh"

namespace tda {

template <typename R>
struct DaTcState {
    std::string err;
    bool eligible(const tda_config&, const Params<R>&) const { return false; }
    int run(Params<R>&, const tda_config&, long long, int, cudaStream_t) { err = "tensor-core path not built"; return -5; }
    void destroy() {}
};

}  // namespace tda
