"""One warm launch of the cfg2 DA workload on a chosen kernel (for ncu captures).
usage: python tools/run_da_once.py <kernel> [chains] [iters]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyda_b200 import lower_problem
from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
from tinyda_b200.workloads import cfg2_da
kernel = sys.argv[1]
C = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
w = cfg2_da()
spec = lower_problem(w["posteriors"], w["proposal"], 10)
theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
eng = Engine(spec, C, dtype="float32", seed=1, store=[STORE_NONE, STORE_STATS], capacity_iterations=iters)
eng.select_kernel(kernel)
eng.init(theta0)
eng.run(100, record=False); eng.sync()
import time
for rep in range(3):
    eng.history_reset()
    t0 = time.perf_counter(); eng.run(iters); eng.sync(); dt = time.perf_counter() - t0
    print("%s: %d chains x %d iterations in %.3f ms -> %.1f M transitions/s" % (eng.kernel(), C, iters, dt * 1e3, C * iters / dt / 1e6))
if os.environ.get("NOSTORE"):
    for rep in range(3):
        t0 = time.perf_counter(); eng.run(iters, record=False); eng.sync(); dt = time.perf_counter() - t0
        print("%s, unrecorded: %d chains x %d iterations in %.3f ms -> %.1f M transitions/s" % (eng.kernel(), C, iters, dt * 1e3, C * iters / dt / 1e6))
