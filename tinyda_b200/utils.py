"""``get_MAP`` / ``get_ML`` (tinyDA/utils.py:204-269) and the device-side link evaluator behind
``Posterior.create_link``.

The reference optimises ``-posterior.create_link(x).posterior`` with scipy, one Python model call per
function evaluation.  Here a *batch* of parameter vectors is turned into Links by the CUDA engine
(one launch of the fused log-prior + forward model + log-likelihood stage that also creates the
chains' initial Links, posterior.py:78-110), so one launch yields the objective and its
central-difference gradient, or a whole differential-evolution population.  There is no CPU path.
"""
import numpy as np
from scipy.optimize import minimize, differential_evolution


class LinkEvaluator:
    """Evaluates Links (log-prior, model output, log-likelihood) for batches of parameter vectors
    on the device.  ``batch`` = the largest number of vectors per call."""

    def __init__(self, posterior, batch=1, dtype="float64", device=0):
        from .lowering import lower_problem
        from .proposal import GaussianRandomWalk
        from .engine import Engine, STORE_FULL
        self.posterior = posterior
        d = np.atleast_1d(posterior.prior.mean).shape[0]
        spec = lower_problem(posterior, GaussianRandomWalk(np.eye(d)))      # the proposal is never used
        self.d, self.batch = d, int(batch)
        self.eng = Engine(spec, self.batch, dtype=dtype, seed=0, store=STORE_FULL, capacity_iterations=0,
                          device=device)

    def __call__(self, thetas):
        """thetas [n, d] (n <= batch) -> (log_prior [n], model_output [n, m], log_likelihood [n])."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        n = thetas.shape[0]
        if n > self.batch or thetas.shape[1] != self.d:
            raise ValueError("expected at most %d parameter vectors of size %d" % (self.batch, self.d))
        full = np.zeros((self.batch, self.d))
        full[:n] = thetas
        full[n:] = thetas[0]
        self.eng.init(full)                       # initial Links of every "chain" = one create_link each
        self.eng.sync()
        prior = self.eng.fetch(0, "prior", 0, 1)[0, :n].astype(np.float64)
        like = self.eng.fetch(0, "like", 0, 1)[0, :n].astype(np.float64)
        out = self.eng.fetch(0, "output", 0, 1)[0, :, :n].T.astype(np.float64)
        return prior, out, like

    def close(self):
        self.eng.close()


def _optimise(posterior, which, kwargs):
    kwargs = dict(kwargs)
    method = kwargs.pop("method", None)
    d = np.atleast_1d(posterior.prior.mean).shape[0]
    sel = (lambda pr, lk: -(pr + lk)) if which == "posterior" else (lambda pr, lk: -lk)
    if method == "differential_evolution":                           # utils.py:228-229, :261-262
        popsize = int(kwargs.get("popsize", 15))
        ev = LinkEvaluator(posterior, batch=max(popsize * d, d + 1))
        try:
            def fun(X):                                              # X [d, S]: a whole population per launch
                X = np.atleast_2d(X.T) if X.ndim == 2 else X[None, :]
                pr, _, lk = ev(X)
                v = sel(pr, lk)
                return v if v.shape[0] > 1 else float(v[0])
            kwargs.setdefault("vectorized", True)
            kwargs.setdefault("updating", "deferred")
            return differential_evolution(fun, **kwargs)["x"]
        finally:
            ev.close()
    x0 = np.atleast_1d(np.asarray(kwargs.pop("initial_parameters", posterior.prior.rvs()), dtype=np.float64))
    ev = LinkEvaluator(posterior, batch=2 * d + 1)
    try:
        user_jac = "jac" in kwargs
        h = float(kwargs.pop("fd_step", 1e-6))

        def fun(x):
            if user_jac:
                pr, _, lk = ev(x[None, :])
                return float(sel(pr, lk)[0])
            # value and central-difference gradient from ONE batch of 2d+1 Links
            step = h * np.maximum(1.0, np.abs(x))
            pts = np.vstack([x[None, :], x[None, :] + np.diag(step), x[None, :] - np.diag(step)])
            pr, _, lk = ev(pts)
            v = sel(pr, lk)
            return float(v[0]), (v[1:d + 1] - v[d + 1:]) / (2.0 * step)

        if not user_jac:
            kwargs["jac"] = True
        return minimize(fun, x0, method=method, **kwargs)["x"]
    finally:
        ev.close()


def get_MAP(posterior, **kwargs):
    """Maximum a posteriori estimate (tinyDA/utils.py:204-235): ``initial_parameters`` (default: a
    prior draw), ``method`` and every other keyword go to ``scipy.optimize.minimize`` /
    ``differential_evolution`` exactly as in the reference; the objective is evaluated on the GPU."""
    return _optimise(posterior, "posterior", kwargs)


def get_ML(posterior, **kwargs):
    """Maximum likelihood estimate (tinyDA/utils.py:238-269)."""
    return _optimise(posterior, "likelihood", kwargs)
