"""Gaussian log-likelihood descriptors (host side).

Mirrors the constructor surface and the factory rule of the reference
(tinyDA/distributions.py:203-243 ``GaussianLogLike``; :246-329 the three Gaussian classes;
:332-449 ``AdaptiveGaussianLogLike``) so that user code keeps working unchanged.  These
objects only DESCRIBE the likelihood: the arithmetic runs in the CUDA kernels
(csrc/tda_kernels.cuh, ``loglike_*``).  ``lower()`` yields the kind + constant buffers the
engine uploads.
"""
import numpy as np

LIK_ISO, LIK_DIAG, LIK_DENSE, LIK_ADAPTIVE = 0, 1, 2, 3


def _check_cov(data, covariance):
    # same checks and messages as distributions.py:227-235 / :370-378
    if not isinstance(covariance, np.ndarray):
        raise TypeError("Covariance must be a 2-D numpy array.")
    if covariance.ndim == 2:
        if not covariance.shape[0] == data.shape[0]:
            raise ValueError("Dimensions of data and covariance do not match.")
        if not covariance.shape[0] == covariance.shape[1]:
            raise ValueError("Covariance must be an NxN array.")
    else:
        raise TypeError("Covariance must be a 2-D numpy array.")


class DefaultGaussianLogLike:
    """Dense-covariance Gaussian log-likelihood, -0.5 r^T inv(cov) r (distributions.py:246-301)."""

    kind = LIK_DENSE

    def __init__(self, data, covariance):
        self.data = np.asarray(data, dtype=np.float64)
        self.cov = np.asarray(covariance, dtype=np.float64)

    def lower(self):
        return dict(kind=self.kind, data=self.data, cov=self.cov)


class DiagonalGaussianLogLike(DefaultGaussianLogLike):
    """-0.5 sum r^2/var_i (distributions.py:304-315)."""

    kind = LIK_DIAG

    def __init__(self, data, covariance):
        self.data = np.asarray(data, dtype=np.float64)
        self.cov = np.diag(np.asarray(covariance, dtype=np.float64)).copy()

    def lower(self):
        return dict(kind=self.kind, data=self.data, var=self.cov)


class IsotropicGaussianLogLike(DefaultGaussianLogLike):
    """-0.5 |r|^2/var (distributions.py:318-329)."""

    kind = LIK_ISO

    def __init__(self, data, variance):
        self.data = np.asarray(data, dtype=np.float64)
        self.var = float(variance)

    def lower(self):
        return dict(kind=self.kind, data=self.data, var=self.var)


class AdaptiveGaussianLogLike(DefaultGaussianLogLike):
    """Bias-corrected dense Gaussian log-likelihood for the adaptive error model
    (distributions.py:332-449).  The per-chain bias mean / covariance and the re-inverted
    precision live on the device (one set per chain), not on this object."""

    kind = LIK_ADAPTIVE

    def __init__(self, data, covariance):
        data = np.asarray(data)
        _check_cov(data, covariance)
        super().__init__(data, covariance)

    def lower(self):
        return dict(kind=self.kind, data=self.data, cov=self.cov)


def GaussianLogLike(data, covariance):
    """Factory with the reference's selection rule (distributions.py:203-243): zero
    off-diagonals -> diagonal; all diagonal entries equal -> isotropic; else dense."""
    data = np.asarray(data)
    _check_cov(data, covariance)
    if np.count_nonzero(covariance - np.diag(np.diag(covariance))) == 0:
        if np.all(np.diag(covariance) == covariance[0, 0]):
            return IsotropicGaussianLogLike(data, covariance[0, 0])
        return DiagonalGaussianLogLike(data, covariance)
    return DefaultGaussianLogLike(data, covariance)


# README.md:65 of the reference calls it tda.AdaptiveLogLike
AdaptiveLogLike = AdaptiveGaussianLogLike
