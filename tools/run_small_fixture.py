"""Throughput of a golden fixture's problem at many chains on the auto-selected kernel and on the lock-step kernel.
usage: python tools/run_small_fixture.py <fixture> [chains] [iterations]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import golden_io
from tinyda_b200.engine import Engine, STORE_NONE, STORE_STATS
name = sys.argv[1]
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
g = golden_io.load(name)
spec = g["spec"]
L = spec["n_levels"]
rng = np.random.default_rng(0)
theta0 = np.resize(g["theta0"], (C, g["theta0"].shape[1])) + 0.01 * rng.standard_normal((C, g["theta0"].shape[1]))
arch = None if g["archive0"] is None else np.resize(g["archive0"], (C,) + g["archive0"].shape[1:])
for kern in ("auto", "generic"):
    eng = Engine(spec, C, dtype="float32", seed=3, store=[STORE_NONE] * (L - 1) + [STORE_STATS], capacity_iterations=iters, archive0=arch)
    eng.select_kernel(kern)
    eng.init(theta0)
    eng.run(iters); eng.sync(); eng.history_reset()
    t0 = time.perf_counter(); eng.run(iters); eng.sync(); dt = time.perf_counter() - t0
    print("%-18s %-8s %d chains x %d finest iterations: %.2f ms -> %.3g finest transitions/s" % (name, eng.kernel(), C, iters, dt * 1e3, C * iters / dt), flush=True)
    eng.close()
