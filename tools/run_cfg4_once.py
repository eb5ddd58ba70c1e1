"""One launch of the cfg4 workload (4-level MLDA on the 1-D Poisson model), with or without the adaptive error
model (for ncu / timing).  usage: python tools/run_cfg4_once.py [chains] [aem|noaem] [iterations] [dtype]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tinyda_b200 import lower_problem, workloads
from tinyda_b200.engine import Engine, STORE_NONE, STORE_STATS
from tinyda_b200.distributions import GaussianLogLike
from tinyda_b200.posterior import Posterior

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
aem = (sys.argv[2] if len(sys.argv) > 2 else "aem") == "aem"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dtype = sys.argv[4] if len(sys.argv) > 4 else "float32"
w = workloads.cfg4_mlda()
kw = w["kwargs"]
if aem:
    spec = lower_problem(w["posteriors"], w["proposal"], kw["subchain_length"], kw["adaptive_error_model"])
else:
    posts = [Posterior(p.prior, GaussianLogLike(p.likelihood.data, 1e-6 * np.eye(p.likelihood.data.size)), p.model)
             for p in w["posteriors"]]
    spec = lower_problem(posts, w["proposal"], kw["subchain_length"], None)
theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(0))
L = spec["n_levels"]
eng = Engine(spec, C, dtype=dtype, seed=1, store=[STORE_NONE] * (L - 1) + [STORE_STATS], capacity_iterations=iters * 3)
eng.init(theta0)
for rep in range(3):
    t0 = time.perf_counter(); eng.run(iters); eng.sync(); dt = time.perf_counter() - t0
    print("%d chains, %s, kernel %s: %.2f ms per fine iteration (%.3g finest transitions/s)" % (C, "AEM" if aem else "no AEM", eng.kernel(), dt / iters * 1e3, C * iters / dt), flush=True)
