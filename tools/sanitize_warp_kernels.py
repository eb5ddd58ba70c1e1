"""Small runs of the two warp-per-chain kernels for compute-sanitizer (memcheck / racecheck).
usage: compute-sanitizer --tool memcheck python tools/sanitize_warp_kernels.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import golden_io
from tinyda_b200 import lower_problem, workloads
from tinyda_b200.engine import Engine, STORE_FULL

# cfg5 shape, shared archive, several chains per warp impossible at this size: 70 chains x 12 steps
w = workloads.cfg5_dream()
spec = lower_problem(w["posteriors"], w["proposal"])
rng = np.random.default_rng(1)
C = 70
eng = Engine(spec, C, dtype="float32", seed=1, store=STORE_FULL, capacity_iterations=12,
             archive0=w["prior"].rvs(C * 16, random_state=rng).reshape(C, 16, 32))
assert eng.kernel() == "dreamw"
eng.init(w["prior"].rvs(C, random_state=rng))
eng.run(12); eng.sync()
print("dreamw ok, accept rate %.3f" % eng.fetch(0, "accept")[1:].mean())
eng.close()

# cfg4 shape with the error model: 40 chains x 2 finest iterations (float32 and float64)
w = workloads.cfg4_mlda()
kw = w["kwargs"]
spec = lower_problem(w["posteriors"], w["proposal"], kw["subchain_length"], kw["adaptive_error_model"])
for dt in ("float32", "float64"):
    eng = Engine(spec, 40, dtype=dt, seed=2, store=STORE_FULL, capacity_iterations=2)
    assert eng.kernel() == "mldaw"
    eng.init(0.3 * w["prior"].rvs(40, random_state=rng))
    eng.run(2); eng.sync()
    print("mldaw %s ok, coarse accept rate %.3f" % (dt, eng.fetch(0, "accept").mean()))
    eng.close()

# adaptive random walk + linear operators (the window ring)
g = golden_io.load("mlda3_aem_linear")
eng = Engine(g["spec"], 33, dtype="float32", seed=3, store=STORE_FULL, capacity_iterations=30)
assert eng.kernel() == "mldaw"
eng.init(np.resize(g["theta0"], (33, g["theta0"].shape[1])))
eng.run(30); eng.sync()
print("mldaw (linear, adaptive) ok")
eng.close()
