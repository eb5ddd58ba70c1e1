"""Loader for the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import os
import glob

import numpy as np

from tinyda_b200.lowering import spec_from_flat

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name, long=False):
    """long=True: tests/golden/long/<name>.npz (make_golden.py --long): normals stored on the fp16 grid
    (float16 of 4096 z), coarse level without parameters."""
    path = os.path.join(GOLDEN_DIR, "long", name + ".npz") if long else os.path.join(GOLDEN_DIR, name + ".npz")
    flat = dict(np.load(path, allow_pickle=False))
    if "z16" in flat:
        flat["z"] = flat["z16"].astype(np.float64) / 4096.0
    spec = spec_from_flat(flat)
    L = spec["n_levels"]
    ref = []
    for l in range(L):
        pre = "ref/l%d/" % l
        ref.append({k[len(pre):]: v for k, v in flat.items() if k.startswith(pre)})
    return dict(spec=spec, theta0=flat["theta0"], z=flat["z"], u=flat["u"],
                iterations=int(flat["iterations"]), consumed=flat["consumed"],
                archive0=flat.get("archive0"), ref=ref)
