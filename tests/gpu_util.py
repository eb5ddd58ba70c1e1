"""Helpers for the GPU parity tests: run the CUDA engine (through the C ABI) on a golden
fixture's problem with the fixture's injected streams."""
import numpy as np

from tinyda_b200.engine import Engine, STORE_FULL
from tinyda_b200.proposal import PROP_AM, svd_factor


def run_engine(g, dtype="float64", iterations=None, store_F=True, kernel=None):
    spec = g["spec"]
    C = g["theta0"].shape[0]
    iters = g["iterations"] if iterations is None else iterations
    kind = int(spec["proposal"]["kind"])
    store = STORE_FULL if store_F else (STORE_FULL & ~4)
    eng = Engine(spec, C, dtype=dtype, rng="injected", streams=(g["z"], g["u"]), store=store,
                 capacity_iterations=iters, archive0=g["archive0"], am_device_refactor=False)
    if kernel is not None:
        eng.select_kernel(kernel)
    eng.init(g["theta0"])
    if kind == PROP_AM:
        # parity mode for Adaptive Metropolis: the covariance factor is refreshed on the host
        # with numpy's SVD (what np.random.multivariate_normal applies), at the same steps as
        # proposal.py:509-510
        period, t0 = int(spec["proposal"]["period"]), int(spec["proposal"]["am_t0"])
        done = 0
        while done < iters:
            n = min(period - done % period, iters - done)
            eng.run(n)
            done += n
            if done % period == 0 and done >= t0:
                sig = eng.get("am_sigma")
                eng.upload_am_factors(np.stack([svd_factor(s) for s in sig]))
    else:
        eng.run(iters)
    eng.sync()
    out = []
    for l in range(spec["n_levels"]):
        h = dict(theta=np.transpose(eng.fetch(l, "theta"), (2, 0, 1)).astype(np.float64),
                 prior=eng.fetch(l, "prior").T.astype(np.float64),
                 like=eng.fetch(l, "like").T.astype(np.float64),
                 acc=eng.fetch(l, "accept").T.astype(bool))
        if store_F:
            h["F"] = np.transpose(eng.fetch(l, "output"), (2, 0, 1)).astype(np.float64)
        out.append(h)
    return out, eng


def first_divergence(acc_a, acc_b):
    """Per chain: index of the first record whose accept flag differs (or n if none)."""
    n = acc_a.shape[1]
    diff = acc_a != acc_b
    return np.where(diff.any(axis=1), diff.argmax(axis=1), n)
