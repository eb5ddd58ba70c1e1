"""Which kernel the engine selects for every golden fixture (float64 / float32).  usage: python tools/fixture_kernels.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io
from tinyda_b200.engine import Engine, STORE_FULL
for name in golden_io.names():
    g = golden_io.load(name)
    sel = []
    for dt in ("float64", "float32"):
        eng = Engine(g["spec"], g["theta0"].shape[0], dtype=dt, rng="injected", streams=(g["z"], g["u"]), store=STORE_FULL & ~4,
                     capacity_iterations=g["iterations"], archive0=g["archive0"], am_device_refactor=False)
        sel.append(eng.kernel())
        eng.close()
    print("%-22s float64 -> %-8s float32 -> %s" % (name, sel[0], sel[1]))
