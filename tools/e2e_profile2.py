"""tda.sample() at BASELINE cfg2's size from burnt-in states, with the phase timings of TDA_PROFILE=1."""
import contextlib, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TDA_PROFILE"] = "1"
import tinyda_b200 as tda
from tinyda_b200 import lower_problem
from tinyda_b200.engine import Engine, STORE_NONE, pinned_empty
from tinyda_b200.workloads import cfg2_da
C = 65536
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = cfg2_da()
spec = lower_problem(w["posteriors"], w["proposal"], 10)
eng = Engine(spec, C, dtype="float32", seed=1, store=STORE_NONE)
eng.init(w["prior"].rvs(C, random_state=np.random.default_rng(1)))
eng.run(300, record=False); eng.sync()
cur = eng.get("theta", 1); eng.close()
th = pinned_empty((C, 64), np.float64); th[:] = cur
for rep in range(4):
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        res = tda.sample(w["posteriors"], w["proposal"], iters, n_chains=C, initial_parameters=th, subchain_length=10,
                         dtype="float32", seed=3 + rep, store_model_output=False, store_coarse_chain=False)
    dt = time.perf_counter() - t0
    sys.stderr.write("== rep %d: %.1f ms -> %.1f M transitions/s\n" % (rep, dt * 1e3, C * iters / dt / 1e6))
    del res
