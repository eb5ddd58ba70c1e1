"""One warm launch of cfg2's problem with an adaptive Gaussian random walk (the 3xTF32 tcgen05 kernel), for ncu.
usage: python tools/run_rw_once.py [chains] [iters]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyda_b200 import lower_problem
from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
from tinyda_b200.workloads import cfg2_rw
C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
w = cfg2_rw()
spec = lower_problem(w["posteriors"], w["proposal"], 10)
eng = Engine(spec, C, dtype="float32", seed=1, store=[STORE_NONE, STORE_STATS], capacity_iterations=iters)
eng.init(w["prior"].rvs(C, random_state=np.random.default_rng(1)))
eng.run(100, record=False); eng.sync()
for rep in range(3):
    eng.history_reset()
    t0 = time.perf_counter(); eng.run(iters); eng.sync(); dt = time.perf_counter() - t0
    print("%s: %d chains x %d iterations in %.3f ms -> %.1f M transitions/s" % (eng.kernel(), C, iters, dt * 1e3, C * iters / dt / 1e6))
