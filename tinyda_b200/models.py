"""Device-resident forward models.

The reference's forward model is an arbitrary Python callable ``F(theta) -> ndarray``
(tinyDA/posterior.py:95-105) with an optional ``gradient(theta, sensitivity)``
(tinyDA/proposal.py:990-1000).  An arbitrary Python callable cannot run inside a CUDA
kernel, so the engine recognises the model CLASSES below and lowers them to a model kind
plus constant buffers (``lower()``).  Anything else raises TypeError in ``sample()`` --
there is no CPU fallback.

Each class still honours the reference's model protocol on NumPy input (``__call__``,
``gradient``) so that the very same object can be handed to the unmodified reference
(that is how the golden fixtures are generated) or to host-side utilities such as a MAP
optimiser.  The sampler never calls these NumPy methods.
"""
import numpy as np

MODEL_LINEAR = 0
MODEL_ROSENBROCK = 1
MODEL_POISSON1D = 2


class _QoiMixin:
    """Quantity of interest: the reference lets a model return ``(output, qoi)`` (posterior.py:95-105)
    and stores the second item on the Link (link.py:38-48).  Device models offer the linear functionals
    ``qoi = Q @ F(theta) + q0`` (Q: (n_qoi, m)); pass ``qoi=Q`` or ``qoi=(Q, q0)``."""

    Q = None
    q0 = None

    def _set_qoi(self, qoi):
        if qoi is None:
            return
        Q, q0 = qoi if isinstance(qoi, tuple) else (qoi, None)
        Q = np.ascontiguousarray(np.atleast_2d(np.asarray(Q, dtype=np.float64)))
        if Q.shape[1] != self.m:
            raise ValueError("qoi operator must have one column per model output")
        self.Q = Q
        self.q0 = np.zeros(Q.shape[0]) if q0 is None else np.asarray(q0, dtype=np.float64).reshape(Q.shape[0])

    def _with_qoi(self, out):
        return out if self.Q is None else (out, self.Q @ out + self.q0)

    def _lower_qoi(self, low):
        low["n_qoi"] = 0 if self.Q is None else int(self.Q.shape[0])
        if self.Q is not None:
            low["qoi_Q"] = self.Q
            low["qoi_q0"] = self.q0
        return low


class LinearModel(_QoiMixin):
    """F(theta) = G @ theta (+ offset).   G: (m, d)."""

    kind = MODEL_LINEAR

    def __init__(self, G, offset=None, qoi=None):
        G = np.ascontiguousarray(np.atleast_2d(np.asarray(G, dtype=np.float64)))
        self.G = G
        self.m, self.d = G.shape
        self.offset = None if offset is None else np.asarray(offset, dtype=np.float64).reshape(self.m)
        self._set_qoi(qoi)

    def __call__(self, parameters):
        out = self.G @ np.asarray(parameters, dtype=np.float64)
        if self.offset is not None:
            out = out + self.offset
        return self._with_qoi(out)

    def gradient(self, parameters, sensitivity):
        return self.G.T @ np.asarray(sensitivity, dtype=np.float64)

    def lower(self):
        # the kernels want G^T, row-major [d][m] (the contraction index outermost)
        off = np.zeros(self.m) if self.offset is None else self.offset
        return self._lower_qoi(dict(kind=self.kind, m=self.m, d=self.d, n_grid=0,
                                    A=np.ascontiguousarray(self.G.T), b=np.ascontiguousarray(off),
                                    scalars=np.zeros(4)))


class Rosenbrock(_QoiMixin):
    """F(x, y) = [(a-x)^2 + b (y-x^2)^2]   (examples/MALA Rosenbrock.ipynb cells 4, 8)."""

    kind = MODEL_ROSENBROCK

    def __init__(self, a=1.0, b=10.0, qoi=None):
        self.a = float(a)
        self.b = float(b)
        self.m, self.d = 1, 2
        self._set_qoi(qoi)

    def __call__(self, parameters):
        x, y = parameters[0], parameters[1]
        return self._with_qoi(np.array([(self.a - x) ** 2 + self.b * (y - x ** 2) ** 2]))

    def gradient(self, parameters, sensitivity):
        x, y = parameters[0], parameters[1]
        dFdx = -2.0 * (self.a - x) - 4.0 * self.b * x * (y - x ** 2)
        dFdy = 2.0 * self.b * (y - x ** 2)
        return np.dot(np.asarray(sensitivity, dtype=np.float64), np.array([[dFdx, dFdy]]))

    def lower(self):
        return self._lower_qoi(dict(kind=self.kind, m=1, d=2, n_grid=0, A=np.zeros(1), b=np.zeros(1),
                                    scalars=np.array([self.a, self.b, 0.0, 0.0])))


class Poisson1D(_QoiMixin):
    """-(k u')' = 1 on (0,1), u(0)=u(1)=0, log k(x) = sum_j theta_j phi_j(x),
    phi_j(x) = sqrt(2) sin((j+1) pi x)/(j+1).  Cell-centred finite differences on n cells:
    unknowns u_1..u_{n-1} at nodes i/n, k evaluated at cell centres (i+1/2)/n, symmetric
    positive-definite tridiagonal system solved with the Thomas algorithm.  The output is u
    at n_sensors equispaced interior sensors x = s/(n_sensors+1); n must be a multiple of
    n_sensors+1 so that sensors sit on nodes at every level.  
    """

    kind = MODEL_POISSON1D

    def __init__(self, n, d, n_sensors=31, qoi=None):
        n = int(n)
        if n % (n_sensors + 1) != 0:
            raise ValueError("n must be a multiple of n_sensors+1")
        self.n = n
        self.d = int(d)
        self.m = int(n_sensors)
        xc = (np.arange(n) + 0.5) / n
        j = np.arange(1, self.d + 1)
        # Phi[i, j] = phi_j(x_{i+1/2})
        self.Phi = np.sqrt(2.0) * np.sin(np.pi * xc[:, None] * j[None, :]) / j[None, :]
        self.stride = n // (n_sensors + 1)
        self._set_qoi(qoi)

    def __call__(self, parameters):
        n = self.n
        h2 = 1.0 / (n * n)
        k = np.exp(self.Phi @ np.asarray(parameters, dtype=np.float64))   # n cell values
        # row i (node i+1): -k[i] u_i + (k[i]+k[i+1]) u_{i+1} - k[i+1] u_{i+2} = h^2
        nn = n - 1
        cp = np.empty(nn)
        dp = np.empty(nn)
        diag = k[0] + k[1]
        cp[0] = -k[1] / diag
        dp[0] = h2 / diag
        for i in range(1, nn):
            a = -k[i]
            diag = (k[i] + k[i + 1]) - a * cp[i - 1]
            cp[i] = -k[i + 1] / diag
            dp[i] = (h2 - a * dp[i - 1]) / diag
        u = np.empty(nn)
        u[nn - 1] = dp[nn - 1]
        for i in range(nn - 2, -1, -1):
            u[i] = dp[i] - cp[i] * u[i + 1]
        return self._with_qoi(u[self.stride - 1::self.stride][:self.m].copy())

    def lower(self):
        return self._lower_qoi(dict(kind=self.kind, m=self.m, d=self.d, n_grid=self.n,
                                    A=np.ascontiguousarray(self.Phi.T), b=np.zeros(1),   # [d][n]
                                    scalars=np.array([float(self.stride), 0.0, 0.0, 0.0])))


def is_device_model(model):
    return isinstance(model, (LinearModel, Rosenbrock, Poisson1D))
