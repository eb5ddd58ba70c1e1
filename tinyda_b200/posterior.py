"""Posterior descriptor (tinyDA/posterior.py:9-151).

``Posterior(prior, likelihood, model)`` keeps the reference's constructor.  In the reference
``create_link`` (posterior.py:78-110) is the per-sample "forward pass" called from the chain
loop; here the same three evaluations (log-prior, forward model, log-likelihood) are one
fused stage of the CUDA kernels and this object only carries the description.  The model must
be one of the device-resident model classes in ``tinyda_b200.models``; anything else is
rejected in ``sample()`` with TypeError (no CPU fallback).
"""
import numpy as np
import scipy.stats as stats

from .models import is_device_model


def lower_prior(prior):
    """scipy.stats.multivariate_normal (frozen) -> dict(mean, cov, LP, logconst).

    logpdf(x) = -0.5*(rank*log(2 pi) + log_pdet + |(x-mean) @ LP|^2) with scipy's own
    whitening matrix LP (scipy/stats/_covariance.py CovViaPSD), so that the device value is
    the reference's value up to summation order."""
    if not isinstance(prior, stats._multivariate.multivariate_normal_frozen):
        raise TypeError(
            "the device engine needs a scipy.stats.multivariate_normal prior "
            "(got %s); there is no CPU fallback" % type(prior).__name__)
    mean = np.atleast_1d(np.asarray(prior.mean, dtype=np.float64))
    co = prior.cov_object
    cov = np.atleast_2d(np.asarray(co.covariance, dtype=np.float64))
    LP = np.atleast_2d(np.asarray(co._LP, dtype=np.float64))
    logconst = float(co.rank * np.log(2 * np.pi) + co.log_pdet)
    return dict(mean=mean, cov=cov, LP=np.ascontiguousarray(LP), logconst=logconst)


class Posterior:
    def __init__(self, prior, likelihood, model=None):
        self.prior = prior
        self.likelihood = likelihood
        self.model = model

    def create_link(self, parameters):
        """posterior.py:78-110: the Link of one parameter vector -- log-prior, model output and
        log-likelihood -- evaluated by the CUDA engine (the fused stage that also creates the
        chains' initial Links).  The per-sample calls of the reference's chain loop never come
        here: they run inside the kernels."""
        from .utils import LinkEvaluator
        from .link import Link
        ev = getattr(self, "_evaluator", None)
        if ev is None:
            ev = self._evaluator = LinkEvaluator(self, batch=1)
        parameters = np.atleast_1d(np.asarray(parameters, dtype=np.float64))
        prior, out, like = ev(parameters[None, :])
        return Link(parameters, float(prior[0]), out[0], float(like[0]), None)

    def lower(self):
        if not is_device_model(self.model):
            raise TypeError(
                "model must be a device-resident model (tinyda_b200.LinearModel, Rosenbrock, "
                "Poisson1D); arbitrary Python callables cannot run inside the CUDA engine and "
                "there is no CPU fallback")
        if not hasattr(self.likelihood, "lower"):
            raise TypeError("likelihood must be a tinyda_b200 Gaussian log-likelihood")
        lik = self.likelihood.lower()
        model = self.model.lower()
        if lik["data"].shape[0] != model["m"]:
            raise ValueError("model output size %d does not match data size %d"
                             % (model["m"], lik["data"].shape[0]))
        return dict(lik=lik, model=model)
