"""cfg5's problem with per-chain archives (DREAMZ: no lock-step barrier) -- what the step costs without the grid barrier.
usage: python tools/run_dreamz_once.py [chains] [steps]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyda_b200 import lower_problem
from tinyda_b200.engine import Engine, STORE_STATS
from tinyda_b200.proposal import DREAMZ
from tinyda_b200.workloads import cfg5_dream
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = cfg5_dream()
spec = lower_problem(w["posteriors"], DREAMZ(M0=16, delta=1, nCR=3))
rng = np.random.default_rng(1)
theta0 = w["prior"].rvs(C, random_state=rng)
archive0 = w["prior"].rvs(C * 16, random_state=rng).reshape(C, 16, 32)
eng = Engine(spec, C, dtype="float32", seed=1, store=STORE_STATS, capacity_iterations=steps * 4, archive0=archive0)
eng.init(theta0)
for rep in range(3):
    t0 = time.perf_counter(); eng.run(steps); eng.sync(); dt = time.perf_counter() - t0
    print("%s, %d chains x %d steps, own archives: %.2f us per step" % (eng.kernel(), C, steps, dt / steps * 1e6))
