"""Timeline probe of the tc16 kernel (GPU): per-step phase durations of CTA 0 / tile 0 from clock64 stamps."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyda_b200 import lower_problem, _lib as L
from tinyda_b200.engine import Engine, STORE_NONE, STORE_STATS
from tinyda_b200.workloads import cfg2_da
w = cfg2_da()
spec = lower_problem(w["posteriors"], w["proposal"], 10)
Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eng = Engine(spec, Cn, dtype="float32", seed=3, store=[STORE_NONE, STORE_STATS], capacity_iterations=4)
eng.select_kernel("tc16")
eng.init(w["prior"].rvs(Cn, random_state=np.random.default_rng(0)))
eng.run(4); eng.sync(); eng.history_reset()
buf = np.zeros((4, 256), dtype=np.int64)
L.check(L.lib.tda_get(eng._h, L.TDA_G_TC16_TIMELINE, 0, buf.ctypes.data_as(C.c_void_p), buf.nbytes))   # arm
eng.run(4); eng.sync()
L.check(L.lib.tda_get(eng._h, L.TDA_G_TC16_TIMELINE, 0, buf.ctypes.data_as(C.c_void_p), buf.nbytes))
t0 = buf[buf > 0].min()
rng = (buf[0, :192].reshape(64, 3) - t0)
mma = (buf[1, :240].reshape(40, 6) - t0)
r0 = (buf[2, :240].reshape(40, 6) - t0)
r1 = (buf[3, :240].reshape(40, 6) - t0)
print("RNG warp (tile 0): step: start, zfree-wait done, image done | gen cycles")
for n in range(24): print(n, rng[n], rng[n, 2] - rng[n, 1])
print("MMA warp tile 0: step: loop top, reqA ok, z ok, G1 issued, reqB ok, G2+G3 issued")
for n in range(24): print(n, mma[n], "period", mma[n, 0] - mma[n - 1, 0] if n else 0)
print("row leader tile 0: step: top, respA ok, residual done, accept known, respB ok, update done")
for n in range(24): print(n, r0[n], "| wait", r0[n, 1] - r0[n, 0], "resid", r0[n, 2] - r0[n, 1], "exch", r0[n, 3] - r0[n, 2], "wB", r0[n, 4] - r0[n, 3], "upd", r0[n, 5] - r0[n, 4])
print("row warp h=1,wq=3 tile 0")
for n in range(24): print(n, r1[n], "| wait", r1[n, 1] - r1[n, 0], "resid", r1[n, 2] - r1[n, 1], "exch", r1[n, 3] - r1[n, 2], "wB", r1[n, 4] - r1[n, 3], "upd", r1[n, 5] - r1[n, 4])
