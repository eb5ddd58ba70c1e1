"""Where does a cfg4 (4-level MLDA + AEM, Poisson) fine iteration spend its time?  Times variants
of the workload on one GPU: with / without the adaptive error model, per number of chains."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tinyda_b200 import lower_problem, workloads
from tinyda_b200.engine import Engine, STORE_NONE, STORE_STATS
from tinyda_b200.distributions import GaussianLogLike
from tinyda_b200.posterior import Posterior


def timed(spec, C, iters, theta0, dtype="float32"):
    L = spec["n_levels"]
    eng = Engine(spec, C, dtype=dtype, seed=1, store=[STORE_NONE] * (L - 1) + [STORE_STATS], capacity_iterations=iters)
    eng.init(theta0)
    eng.run(1)
    eng.sync()
    eng.history_reset()
    t0 = time.perf_counter()
    eng.run(iters)
    eng.sync()
    dt = time.perf_counter() - t0
    eng.close()
    return dt / iters * 1e3


w = workloads.cfg4_mlda()
kw = w["kwargs"]
for C in (4096, 16384, 32768):
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(0))
    spec = lower_problem(w["posteriors"], w["proposal"], kw["subchain_length"], kw["adaptive_error_model"])
    a = timed(spec, C, 2, theta0)
    sig2 = 1e-6
    posts = [Posterior(p.prior, GaussianLogLike(p.likelihood.data, sig2 * np.eye(p.likelihood.data.size)), p.model)
             for p in w["posteriors"]]
    spec2 = lower_problem(posts, w["proposal"], kw["subchain_length"], None)
    b = timed(spec2, C, 2, theta0)
    print("chains %6d: %.1f ms / fine iteration with AEM, %.1f ms without  (%.0f / %.0f transitions/s)"
          % (C, a, b, C / a * 1e3, C / b * 1e3), flush=True)
