"""Does the cost of engine construction grow from one tda.sample() call to the next? (diagnostic)"""
import contextlib, io, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyda_b200 as tda
from tinyda_b200 import lower_problem
from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE, pinned_empty
from tinyda_b200.workloads import cfg2_da
C = 65536
w = cfg2_da()
theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
t = time.perf_counter
for rep in range(4):
    t0 = t(); spec = lower_problem(w["posteriors"], w["proposal"], 10); t1 = t()
    eng = Engine(spec, C, dtype="float32", seed=1, store=[STORE_NONE, STORE_STATS], capacity_iterations=123); t2 = t()
    eng.close(); t3 = t()
    print("bare: lower %.2f ms, Engine() %.2f ms, close %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
keep = os.environ.get("KEEP")
held = []
for rep in range(8):
    t0 = t()
    with contextlib.redirect_stdout(io.StringIO()):
        res = tda.sample(w["posteriors"], w["proposal"], 1000, n_chains=C, initial_parameters=theta0, subchain_length=10,
                         dtype="float32", seed=3 + rep, store_model_output=False, store_coarse_chain=False)
    t1 = t()
    if keep: held.append(res)
    del res
    print("sample() call %d: %.1f ms, release %.1f ms" % (rep, (t1 - t0) * 1e3, (t() - t1) * 1e3), flush=True)
