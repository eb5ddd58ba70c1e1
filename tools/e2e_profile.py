"""Where the wall time of one tda.sample() call goes at BASELINE cfg2's size (65536 chains)."""
import contextlib, io, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyda_b200 as tda
from tinyda_b200 import lower_problem
from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE, pinned_empty
from tinyda_b200.workloads import cfg2_da

C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = cfg2_da()
t = time.perf_counter
def lap(name, t0):
    print("%-34s %8.2f ms" % (name, (t() - t0) * 1e3)); return t()
theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
for rep in range(3):
    print("--- rep", rep)
    t0 = t()
    spec = lower_problem(w["posteriors"], w["proposal"], 10); t0 = lap("lower_problem", t0)
    chunk = 123
    eng = Engine(spec, C, dtype="float32", seed=1, store=[STORE_NONE, STORE_STATS], capacity_iterations=chunk); t0 = lap("Engine()", t0)
    eng.init(theta0); eng.sync(); t0 = lap("init (H2D + initial links)", t0)
    eng.run(chunk); eng.sync(); t0 = lap("run %d (first: tc16 prepare)" % chunk, t0)
    eng.compact_begin(0, chunk + 1, True, ("theta", "stats"), 0); eng.sync(); t0 = lap("compact_begin (first: allocs)", t0)
    ch = eng.compact_collect(0); eng.compact_sync(); t0 = lap("compact_collect+sync (%d MB)" % ((ch.theta.nbytes + ch.accept.nbytes) >> 20), t0)
    eng.history_reset(); eng.run(chunk); eng.sync(); t0 = lap("run %d" % chunk, t0)
    eng.compact_begin(0, chunk, False, ("theta", "stats"), 1); eng.sync(); t0 = lap("compact_begin", t0)
    ch2 = eng.compact_collect(1); eng.compact_sync(); t0 = lap("compact_collect+sync", t0)
    eng.close(); t0 = lap("close", t0)
    del ch, ch2
    with contextlib.redirect_stdout(io.StringIO()):
        res = tda.sample(w["posteriors"], w["proposal"], iters, n_chains=C, initial_parameters=theta0, subchain_length=10,
                         dtype="float32", seed=3, store_model_output=False, store_coarse_chain=False)
    t0 = lap("tda.sample(%d iterations)" % iters, t0)
    del res
