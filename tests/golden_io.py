"""Loader for the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import os
import glob

import numpy as np

from tinyda_b200.lowering import spec_from_flat

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    flat = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    spec = spec_from_flat(flat)
    L = spec["n_levels"]
    ref = []
    for l in range(L):
        pre = "ref/l%d/" % l
        ref.append({k[len(pre):]: v for k, v in flat.items() if k.startswith(pre)})
    return dict(spec=spec, theta0=flat["theta0"], z=flat["z"], u=flat["u"],
                iterations=int(flat["iterations"]), consumed=flat["consumed"],
                archive0=flat.get("archive0"), ref=ref)
