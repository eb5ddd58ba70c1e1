"""Synthetic workloads of BASELINE.json's configs, built with the
package's own tinyDA-style classes.  Used by bench.py and the full-size GPU tests."""
import numpy as np
import scipy.stats as stats

from .models import LinearModel, Rosenbrock, Poisson1D
from .distributions import GaussianLogLike, AdaptiveGaussianLogLike
from .posterior import Posterior
from .proposal import CrankNicolson, GaussianRandomWalk, MALA, DREAM


def exp_cov(d, ell=0.2):
    x = np.linspace(0, 1, d)
    return np.exp(-np.abs(x[:, None] - x[None, :]) / ell)


def cfg1_linreg(seed=1):
    """README linear regression: 2-param Gaussian prior, 100 obs, adaptive RWMH."""
    rng = np.random.default_rng(seed)
    x = np.linspace(0, 1, 100)
    y = 1 + 2 * x + 0.2 * rng.standard_normal(100)
    prior = stats.multivariate_normal(np.zeros(2), np.eye(2))
    G = np.stack([np.ones_like(x), x], axis=1)
    post = Posterior(prior, GaussianLogLike(y, 0.04 * np.eye(100)), LinearModel(G))
    prop = GaussianRandomWalk(C=np.eye(2), scaling=0.1, adaptive=True)
    return dict(posteriors=[post], proposal=prop, kwargs={}, prior=prior,
                G=G, y=y, sigma2=0.04, name="cfg1: RWMH linear regression 2 params / 100 obs")


def cfg2_da(seed=2, d=64, m_f=1024, m_c=128, J=10, beta=0.05):
    """Two-level DA, pCN, linear-Gaussian inverse problem 64 params / 1024 obs (coarse = strided
    128-obs subset), subsampling_rate=10."""
    rng = np.random.default_rng(seed)
    cov = exp_cov(d)
    prior = stats.multivariate_normal(np.zeros(d), cov)
    G = rng.standard_normal((m_f, d)) / 8
    truth = prior.rvs(random_state=rng)
    y = G @ truth + 0.1 * rng.standard_normal(m_f)
    idx = np.arange(0, m_f, m_f // m_c)[:m_c]
    pc = Posterior(prior, GaussianLogLike(y[idx], 0.01 * np.eye(m_c)), LinearModel(G[idx]))
    pf = Posterior(prior, GaussianLogLike(y, 0.01 * np.eye(m_f)), LinearModel(G))
    return dict(posteriors=[pc, pf], proposal=CrankNicolson(scaling=beta), kwargs=dict(subchain_length=J),
                prior=prior, G=G, y=y, sigma2=0.01,
                name="cfg2: 2-level DA, pCN, linear-Gaussian %d params / %d obs (coarse %d), J=%d" % (d, m_f, m_c, J))


def cfg2_rw(seed=2, adaptive=True):
    """cfg2's problem with the proposal most tinyDA notebooks use: an adaptively scaled Gaussian random walk (dense
    covariance) instead of pCN -- the acceptance then needs the proposal's log-prior too."""
    from .proposal import GaussianRandomWalk
    w = cfg2_da(seed=seed)
    w["proposal"] = GaussianRandomWalk(C=exp_cov(64, 0.3), scaling=0.02, adaptive=adaptive, period=100)
    w["name"] = "cfg2 shape, 2-level DA with GaussianRandomWalk(adaptive=%s): 64 params / 1024 obs (coarse 128), J=10" % adaptive
    return w


def cfg3_mala(seed=3):
    """MALA on the 2-D Rosenbrock likelihood (examples/MALA Rosenbrock.ipynb)."""
    prior = stats.multivariate_normal(np.zeros(2), np.eye(2))
    post = Posterior(prior, GaussianLogLike(np.array([0.0]), np.eye(1)), Rosenbrock(1, 10))
    return dict(posteriors=[post], proposal=MALA(scaling=0.01, adaptive=True), kwargs={}, prior=prior,
                name="cfg3: MALA, 2-D Rosenbrock")


def cfg4_mlda(seed=4, d=16, ns=(64, 128, 256, 512), n_sensors=31, J=(10, 5, 5)):
    """4-level MLDA + state-independent AEM on the synthetic 1-D Poisson inversion."""
    rng = np.random.default_rng(seed)
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    truth = 0.5 * prior.rvs(random_state=rng)
    sig = 1e-3
    y = Poisson1D(2048, d, n_sensors)(truth) + sig * rng.standard_normal(n_sensors)
    posts = []
    for i, n in enumerate(ns):
        lk = (AdaptiveGaussianLogLike(y, sig ** 2 * np.eye(n_sensors)) if i < len(ns) - 1
              else GaussianLogLike(y, sig ** 2 * np.eye(n_sensors)))
        posts.append(Posterior(prior, lk, Poisson1D(n, d, n_sensors)))
    prop = GaussianRandomWalk(C=1e-4 * np.eye(d))
    return dict(posteriors=posts, proposal=prop,
                kwargs=dict(subchain_length=list(J), adaptive_error_model="state-independent"), prior=prior,
                name="cfg4: 4-level MLDA + AEM, 1-D Poisson grids %s, J=%s" % (list(ns), list(J)))


def cfg5_dream(seed=5, d=32, m=256):
    """DREAM(Z) with the shared archive on a 32-param linear-Gaussian problem."""
    rng = np.random.default_rng(seed)
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    G = rng.standard_normal((m, d)) / np.sqrt(d)
    truth = prior.rvs(random_state=rng)
    y = G @ truth + 0.1 * rng.standard_normal(m)
    post = Posterior(prior, GaussianLogLike(y, 0.01 * np.eye(m)), LinearModel(G))
    return dict(posteriors=[post], proposal=DREAM(M0=16, delta=1, nCR=3), kwargs={}, prior=prior,
                G=G, y=y, sigma2=0.01, name="cfg5: DREAM(Z) shared archive, %d params / %d obs" % (d, m))


def conjugate_posterior(G, y, sigma2, prior):
    """Closed-form posterior of a linear-Gaussian problem: N(mu_post, Sigma_post)."""
    Cinv = np.linalg.inv(np.atleast_2d(prior.cov))
    S = np.linalg.inv(Cinv + G.T @ G / sigma2)
    mu = S @ (G.T @ y / sigma2 + Cinv @ np.atleast_1d(prior.mean))
    return mu, S


WORKLOADS = dict(cfg1=cfg1_linreg, cfg2=cfg2_da, cfg2rw=cfg2_rw, cfg3=cfg3_mala, cfg4=cfg4_mlda, cfg5=cfg5_dream)
