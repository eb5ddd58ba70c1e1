"""One tda.sample() call of the bench's e2e shape (cfg2, float32, compacted history) for ncu launch lists.
usage: python tools/sample_call_once.py [chains] [iterations] [burn]   (burn > 0: start from burnt-in states like bench.py)"""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyda_b200 as tda
from tinyda_b200.engine import pinned_empty
from tinyda_b200.workloads import cfg2_da

C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
w = cfg2_da()
theta0 = pinned_empty((C, 64), np.float64)
theta0[:] = w["prior"].rvs(C, random_state=np.random.default_rng(1))
burn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if burn:
    from tinyda_b200 import lower_problem
    from tinyda_b200.engine import Engine, STORE_NONE
    eng = Engine(lower_problem(w["posteriors"], w["proposal"], 10), C, dtype="float32", seed=1, store=[STORE_NONE, STORE_NONE],
                 capacity_iterations=1)
    eng.init(theta0); eng.run(burn, record=False); eng.sync()
    theta0[:] = eng.get("theta", 1)
    eng.close()
with contextlib.redirect_stdout(io.StringIO()):
    res = tda.sample(w["posteriors"], w["proposal"], iters, n_chains=C, initial_parameters=theta0, subchain_length=10,
                     dtype="float32", seed=3, store_model_output=False, store_coarse_chain=False)
print("done", res["iterations"])
