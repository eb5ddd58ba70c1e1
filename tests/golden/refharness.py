"""Harness that imports the UNMODIFIED reference tinyDA from /root/reference and runs
it under injected random streams.  TEST INFRASTRUCTURE ONLY: it exists to generate the
golden fixtures under tests/golden/*.npz (see make_golden.py) and to cross-check the
oracle in this container.  /root/reference does not exist on the GPU box, so nothing in
the gpu tests / smoke / bench imports this module.

What it does
------------
* ray / arviz / xarray are not installed and tinyDA imports them unconditionally
  (tinyDA/proposal.py:11, tinyDA/diagnostics.py:2-3): three in-memory stub modules are
  put into sys.modules before the import.  The ray stub is a synchronous shim good enough
  for tinyDA/ray.py:366-384 (ArchiveManager).
* numpy.random.{multivariate_normal, normal, uniform, random, choice, randint} are patched
  with fakes that serve from two pre-drawn per-chain streams: standard normals and
  U(0,1) uniforms.  tinyDA resolves np.random.X at call time, so patching the module
  attributes is enough.  The maps from base draws to what numpy would have returned:
    multivariate_normal(mean, C)   mean + z @ (sqrt(s)[:,None]*Vt),  U,s,Vt = svd(C)
                                   (numpy/random/mtrand.pyx, the legacy default 'svd' path)
    normal(loc, scale, size)       loc + scale*z
    uniform(lo, hi, size)          lo + (hi-lo)*u
    random()                       u
    choice(M, 2, replace=False)    r1=floor(u1*M); r2=floor(u2*(M-1)); r2 += (r2>=r1)
    choice(n, p=p)                 first k with cumsum(p)[k] > u   (k clipped to n-1)
    choice(n)                      floor(u*n)
    choice(seq[, p=p])             seq[choice(len(seq)[, p=p])]
    randint(lo, hi)                lo + floor(u*(hi-lo))
  The integer maps are OUR definition of how integers derive from uniforms (numpy's own
  integer draws are not reproducible from a uniform stream); the engine uses the same maps.
"""
import sys
import types
import contextlib
import numpy as np

REFERENCE_PATH = "/root/reference"


def _install_stubs():
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")

        class _Method:
            def __init__(self, fn):
                self._fn = fn

            def remote(self, *a, **k):
                return self._fn(*a, **k)

        class _Handle:
            def __init__(self, obj):
                object.__setattr__(self, "_obj", obj)

            def __getattr__(self, name):
                return _Method(getattr(self._obj, name))

            def __deepcopy__(self, memo):
                return self

        def remote(cls=None, **kw):
            def wrap(c):
                c.remote = classmethod(lambda kls, *a, **k: _Handle(kls(*a, **k)))
                return c
            return wrap(cls) if cls is not None else wrap

        ray.remote = remote
        ray.init = lambda *a, **k: None
        ray.get = lambda x: x
        ray.is_stub = True
        sys.modules["ray"] = ray
    for name in ("arviz", "xarray"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)


def import_reference():
    """Returns the reference package module (tinyDA), imported from /root/reference."""
    _install_stubs()
    if REFERENCE_PATH not in sys.path:
        sys.path.insert(0, REFERENCE_PATH)
    import tinyDA  # noqa
    return tinyDA


class Streams:
    """Two base streams (standard normals, uniforms) with cursors."""

    def __init__(self, normals, uniforms):
        self.z = np.asarray(normals, dtype=np.float64)
        self.u = np.asarray(uniforms, dtype=np.float64)
        self.nz = 0
        self.nu = 0

    def take_z(self, n):
        out = self.z[self.nz:self.nz + n]
        if out.shape[0] != n:
            raise RuntimeError("normal stream exhausted")
        self.nz += n
        return out.copy()

    def take_u(self, n):
        out = self.u[self.nu:self.nu + n]
        if out.shape[0] != n:
            raise RuntimeError("uniform stream exhausted")
        self.nu += n
        return out.copy()


def svd_factor(C):
    """The d x d matrix T with np.random.multivariate_normal(0, C) == z @ T."""
    C = np.atleast_2d(np.asarray(C, dtype=np.float64))
    _, s, vt = np.linalg.svd(C)
    return np.sqrt(s)[:, None] * vt


@contextlib.contextmanager
def injected(streams):
    """Patch numpy.random's legacy module-level functions to serve from `streams`."""
    S = streams
    saved = {k: getattr(np.random, k) for k in
             ("multivariate_normal", "normal", "uniform", "random", "choice", "randint")}

    def _size(size):
        if size is None:
            return None, 1
        if isinstance(size, (int, np.integer)):
            return (int(size),), int(size)
        return tuple(size), int(np.prod(size))

    def multivariate_normal(mean, cov, size=None, **kw):
        assert size is None
        mean = np.asarray(mean, dtype=np.float64)
        z = S.take_z(mean.shape[0])
        return mean + z @ svd_factor(cov)

    def normal(loc=0.0, scale=1.0, size=None):
        shp, n = _size(size)
        z = S.take_z(n)
        out = loc + scale * z
        return out[0] if shp is None else out.reshape(shp)

    def uniform(low=0.0, high=1.0, size=None):
        shp, n = _size(size)
        u = S.take_u(n)
        out = low + (high - low) * u
        return out[0] if shp is None else out.reshape(shp)

    def random(size=None):
        assert size is None
        return float(S.take_u(1)[0])

    def choice(a, size=None, replace=True, p=None):
        if not isinstance(a, (int, np.integer)):
            # choice over a sequence of objects (MultipleTry, ray.py:309-316): pick the index
            seq = list(a)
            return seq[choice(len(seq), size=size, replace=replace, p=p)]
        n = int(a)
        if size is None and p is None:
            return min(int(np.floor(S.take_u(1)[0] * n)), n - 1)
        if size is None and p is not None:
            u = S.take_u(1)[0]
            cs = np.cumsum(np.asarray(p, dtype=np.float64))
            return int(min(np.searchsorted(cs, u, side="right"), n - 1))
        assert size == 2 and replace is False and p is None
        u1, u2 = S.take_u(2)
        r1 = min(int(np.floor(u1 * n)), n - 1)
        r2 = min(int(np.floor(u2 * (n - 1))), n - 2)
        if r2 >= r1:
            r2 += 1
        return np.array([r1, r2])

    def randint(low, high=None, size=None):
        assert size is None and high is not None
        u = S.take_u(1)[0]
        return int(low + min(int(np.floor(u * (high - low))), high - low - 1))

    np.random.multivariate_normal = multivariate_normal
    np.random.normal = normal
    np.random.uniform = uniform
    np.random.random = random
    np.random.choice = choice
    np.random.randint = randint
    try:
        yield S
    finally:
        for k, v in saved.items():
            setattr(np.random, k, v)


def links_to_arrays(links):
    """List of reference Link objects -> dict of arrays."""
    out = {
        "parameters": np.array([l.parameters for l in links], dtype=np.float64),
        "prior": np.array([l.prior for l in links], dtype=np.float64),
        "likelihood": np.array([l.likelihood for l in links], dtype=np.float64),
        "model_output": np.array([l.model_output for l in links], dtype=np.float64),
    }
    return out
