"""Problem definitions shared by the golden-fixture generator, the oracle tests and the GPU
parity tests.  ``build(tda, name)`` constructs the posteriors / proposal with the classes of
the given module -- either the unmodified reference (``tinyDA``) or this repo's
``tinyda_b200`` -- from identical raw arrays; the class names and constructor arguments are
the same on both sides, which is the point of the drop-in surface.  Forward models are
always the device-resident model classes (NumPy-callable, so the reference accepts them).
"""
import numpy as np
import scipy.stats as stats

from tinyda_b200.models import LinearModel, Rosenbrock, Poisson1D


def _exp_cov(d, ell=0.2):
    x = np.linspace(0, 1, d)
    return np.exp(-np.abs(x[:, None] - x[None, :]) / ell)


def _linear_levels(rng, d, ms, sigma, prior, coarse_mode="subset", perturb=0.0):
    """Fine operator G (ms[-1] x d) and coarser ones: row subsets (strided) of the fine one,
    optionally perturbed (so that there is a model bias for the error model)."""
    m_f = ms[-1]
    G = rng.standard_normal((m_f, d)) / np.sqrt(d)
    truth = prior.rvs(random_state=rng)
    truth = np.atleast_1d(truth)
    y = G @ truth + sigma * rng.standard_normal(m_f)
    out = []
    for i, m in enumerate(ms):
        if coarse_mode == "subset":
            idx = np.arange(0, m_f, m_f // m)[:m]
        else:
            idx = np.arange(m_f)
        Gl = G[idx].copy()
        if i < len(ms) - 1 and perturb:
            Gl = Gl + perturb * (len(ms) - 1 - i) * rng.standard_normal(Gl.shape)
        out.append((Gl, y[idx].copy()))
    return out


CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


# Each case returns a dict:
#   build(tda) -> (posteriors(list), proposal, kwargs for chain/sample)
#   n_chains, iterations, seed, store_F (bool)

@case
def mh_rwmh_linreg():
    """cfg1: README linear regression (examples/Basic Sampler.ipynb), RWMH with adaptive scaling."""
    rng = np.random.default_rng(1)
    x = np.linspace(0, 1, 100)
    y = 1 + 2 * x + 0.2 * rng.standard_normal(100)
    prior = stats.multivariate_normal(np.zeros(2), np.eye(2))
    G = np.stack([np.ones_like(x), x], axis=1)

    def build(tda):
        lik = tda.GaussianLogLike(y, 0.04 * np.eye(100))
        post = tda.Posterior(prior, lik, LinearModel(G))
        prop = tda.GaussianRandomWalk(C=np.eye(2), scaling=0.1, adaptive=True)
        return [post], prop, {}
    return dict(build=build, n_chains=4, iterations=350, seed=11, prior=prior)


@case
def mh_rwmh_dense():
    """RWMH with a dense proposal covariance, non-zero prior mean, dense likelihood covariance."""
    rng = np.random.default_rng(21)
    d, m = 5, 12
    A = rng.standard_normal((d, d))
    prior = stats.multivariate_normal(rng.standard_normal(d), A @ A.T + d * np.eye(d))
    (G, y), = _linear_levels(rng, d, [m], 0.3, prior)
    B = rng.standard_normal((m, m))
    cov = 0.09 * (np.eye(m) + 0.05 * (B @ B.T))
    Cp = 0.05 * (np.eye(d) + 0.3 * np.ones((d, d)))

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, cov), LinearModel(G, offset=0.1 * np.arange(m)))
        return [post], tda.GaussianRandomWalk(C=Cp, scaling=0.7), {}
    return dict(build=build, n_chains=3, iterations=120, seed=22, prior=prior)


@case
def mh_pcn_diag():
    """pCN, diagonal (non-isotropic) likelihood."""
    rng = np.random.default_rng(31)
    d, m = 8, 16
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d))
    (G, y), = _linear_levels(rng, d, [m], 0.2, prior)
    var = 0.04 * (1 + np.arange(m) / m)

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, np.diag(var)), LinearModel(G))
        return [post], tda.CrankNicolson(scaling=0.2, adaptive=True, period=40), {}
    return dict(build=build, n_chains=3, iterations=130, seed=32, prior=prior)


class _NormalQ:
    """A multivariate normal q for IndependenceSampler whose draws go through
    np.random.multivariate_normal (so that the golden harness can inject them); the reference only
    asks for .rvs() and .logpdf(), the device lowering for .mean and .cov."""

    def __init__(self, mean, cov):
        self.mean, self.cov = np.asarray(mean, dtype=np.float64), np.asarray(cov, dtype=np.float64)
        self._frozen = stats.multivariate_normal(self.mean, self.cov)

    def rvs(self, size=1):
        return np.array([np.random.multivariate_normal(self.mean, self.cov) for _ in range(size)])

    def logpdf(self, x):
        return self._frozen.logpdf(x)


@case
def mh_independence():
    """IndependenceSampler (proposal.py:64-131) with a normal q roughly matched to the posterior."""
    rng = np.random.default_rng(47)
    d, m = 3, 8
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    (G, y), = _linear_levels(rng, d, [m], 0.4, prior)
    S = np.linalg.inv(np.eye(d) + G.T @ G / 0.16)
    mu = S @ (G.T @ y / 0.16)
    q = _NormalQ(mu + 0.1, 1.8 * S + 0.02 * np.eye(d))

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.16 * np.eye(m)), LinearModel(G))
        return [post], tda.IndependenceSampler(q), {}
    return dict(build=build, n_chains=3, iterations=150, seed=48, prior=prior)


@case
def mh_owpcn():
    """Operator-weighted pCN (proposal.py:515-605): state and noise operators from sqrtm."""
    rng = np.random.default_rng(33)
    d, m = 6, 14
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d, 0.3))
    (G, y), = _linear_levels(rng, d, [m], 0.2, prior)
    Q = rng.standard_normal((d, d))
    B = 0.5 * (np.eye(d) + 0.2 * (Q @ Q.T) / d)           # SPD with scaling*B < I

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.04 * np.eye(m)), LinearModel(G))
        return [post], tda.OperatorWeightedCrankNicolson(B, scaling=0.08), {}
    return dict(build=build, n_chains=3, iterations=130, seed=34, prior=prior)


@case
def mh_owpcn_adaptive():
    """Operator-weighted pCN with adaptive=True, the way examples/Operator-weighted pCN.ipynb uses it:
    both operators are re-derived from the adapted step size every period (proposal.py:581-591)."""
    rng = np.random.default_rng(43)
    d, m = 5, 12
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d, 0.3))
    (G, y), = _linear_levels(rng, d, [m], 0.2, prior)
    Q = rng.standard_normal((d, d))
    B = 0.6 * (np.eye(d) + 0.3 * (Q @ Q.T) / d)

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.04 * np.eye(m)), LinearModel(G))
        return [post], tda.OperatorWeightedCrankNicolson(B, scaling=0.05, adaptive=True, period=20), {}
    return dict(build=build, n_chains=3, iterations=150, seed=44, prior=prior)


@case
def da_owpcn():
    """Two-level DA with the operator-weighted pCN as the coarse proposal."""
    rng = np.random.default_rng(35)
    d = 6
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d, 0.3))
    (Gc, yc), (Gf, yf) = _linear_levels(rng, d, [6, 24], 0.15, prior)
    B = np.diag(np.linspace(0.5, 1.5, d))

    def build(tda):
        pc = tda.Posterior(prior, tda.GaussianLogLike(yc, 0.0225 * np.eye(6)), LinearModel(Gc))
        pf = tda.Posterior(prior, tda.GaussianLogLike(yf, 0.0225 * np.eye(24)), LinearModel(Gf))
        return [pc, pf], tda.OperatorWeightedCrankNicolson(B, scaling=0.05), dict(subchain_length=3)
    return dict(build=build, n_chains=3, iterations=60, seed=36, prior=prior)


@case
def mh_mtm_rwmh():
    """MultipleTry (ray.py:213-354) around a symmetric random walk: MTM(II), k = 3."""
    rng = np.random.default_rng(37)
    d, m = 4, 10
    prior = stats.multivariate_normal(0.1 * np.ones(d), np.eye(d))
    (G, y), = _linear_levels(rng, d, [m], 0.3, prior)

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.09 * np.eye(m)), LinearModel(G))
        return [post], tda.MultipleTry(tda.GaussianRandomWalk(C=0.08 * np.eye(d)), 3), {}
    return dict(build=build, n_chains=3, iterations=90, seed=38, prior=prior)


@case
def mh_mtm_pcn():
    """MultipleTry around pCN: MTM(I), the transition densities CrankNicolson.get_q enter the
    candidate and reference weights; k = 4."""
    rng = np.random.default_rng(39)
    d, m = 6, 12
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d, 0.4))
    (G, y), = _linear_levels(rng, d, [m], 0.2, prior)

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.04 * np.eye(m)), LinearModel(G))
        return [post], tda.MultipleTry(tda.CrankNicolson(scaling=0.25), 4), {}
    return dict(build=build, n_chains=3, iterations=80, seed=40, prior=prior)


@case
def da_mtm_rwmh():
    """Delayed Acceptance whose coarse proposal is a MultipleTry random walk (k = 3)."""
    rng = np.random.default_rng(41)
    d = 4
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    (Gc, yc), (Gf, yf) = _linear_levels(rng, d, [6, 24], 0.2, prior, perturb=0.02)

    def build(tda):
        pc = tda.Posterior(prior, tda.GaussianLogLike(yc, 0.04 * np.eye(6)), LinearModel(Gc))
        pf = tda.Posterior(prior, tda.GaussianLogLike(yf, 0.04 * np.eye(24)), LinearModel(Gf))
        return [pc, pf], tda.MultipleTry(tda.GaussianRandomWalk(C=0.03 * np.eye(d)), 3), dict(subchain_length=3)
    return dict(build=build, n_chains=3, iterations=40, seed=42, prior=prior)


@case
def da_pcn_small():
    """cfg2 in miniature: two-level DA, pCN, coarse = strided observation subset, J=3."""
    rng = np.random.default_rng(2)
    d = 8
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d))
    (Gc, yc), (Gf, yf) = _linear_levels(rng, d, [8, 32], 0.1, prior)

    def build(tda):
        pc = tda.Posterior(prior, tda.GaussianLogLike(yc, 0.01 * np.eye(8)), LinearModel(Gc))
        pf = tda.Posterior(prior, tda.GaussianLogLike(yf, 0.01 * np.eye(32)), LinearModel(Gf))
        return [pc, pf], tda.CrankNicolson(scaling=0.1), dict(subchain_length=3)
    return dict(build=build, n_chains=4, iterations=80, seed=42, prior=prior)


@case
def da_pcn_cfg2():
    """cfg2 at its real shape (64 params, 1024 / 128 observations, J=10), few chains."""
    rng = np.random.default_rng(2)
    d = 64
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d))
    G = rng.standard_normal((1024, d)) / 8
    truth = prior.rvs(random_state=rng)
    y = G @ truth + 0.1 * rng.standard_normal(1024)
    idx = np.arange(0, 1024, 8)

    def build(tda):
        pc = tda.Posterior(prior, tda.GaussianLogLike(y[idx], 0.01 * np.eye(128)), LinearModel(G[idx]))
        pf = tda.Posterior(prior, tda.GaussianLogLike(y, 0.01 * np.eye(1024)), LinearModel(G))
        return [pc, pf], tda.CrankNicolson(scaling=0.05), dict(subchain_length=10)
    return dict(build=build, n_chains=2, iterations=25, seed=43, prior=prior, store_F=False)


@case
def da_rwmh_adaptive():
    """DA with adaptively scaled RWMH: the adaptation window contains alignment entries
    (SURVEY.md appendix A.9)."""
    rng = np.random.default_rng(5)
    d = 4
    prior = stats.multivariate_normal(0.2 * np.ones(d), np.eye(d))
    (Gc, yc), (Gf, yf) = _linear_levels(rng, d, [6, 24], 0.2, prior, perturb=0.02)

    def build(tda):
        pc = tda.Posterior(prior, tda.GaussianLogLike(yc, 0.04 * np.eye(6)), LinearModel(Gc))
        pf = tda.Posterior(prior, tda.GaussianLogLike(yf, 0.04 * np.eye(24)), LinearModel(Gf))
        prop = tda.GaussianRandomWalk(C=0.02 * np.eye(d), scaling=1.0, adaptive=True, period=25)
        return [pc, pf], prop, dict(subchain_length=4)
    return dict(build=build, n_chains=3, iterations=90, seed=52, prior=prior)


@case
def da_aem_linear():
    """DA + state-independent adaptive error model; coarse model = perturbed fine model."""
    rng = np.random.default_rng(6)
    d, m = 4, 10
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    (Gc, yc), (Gf, yf) = _linear_levels(rng, d, [m, m], 0.1, prior, coarse_mode="same", perturb=0.05)

    def build(tda):
        pc = tda.Posterior(prior, tda.AdaptiveGaussianLogLike(yc, 0.01 * np.eye(m)), LinearModel(Gc))
        pf = tda.Posterior(prior, tda.GaussianLogLike(yf, 0.01 * np.eye(m)), LinearModel(Gf))
        prop = tda.GaussianRandomWalk(C=0.01 * np.eye(d))
        return [pc, pf], prop, dict(subchain_length=3, adaptive_error_model="state-independent")
    return dict(build=build, n_chains=3, iterations=60, seed=62, prior=prior)


@case
def da_sdaem_rwmh():
    """DA + STATE-DEPENDENT adaptive error model (chain.py:446-473, 501-522) with a symmetric
    proposal and the recommended subchain length 1."""
    rng = np.random.default_rng(61)
    d, m = 4, 10
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    (Gc, yc), (Gf, yf) = _linear_levels(rng, d, [m, m], 0.1, prior, coarse_mode="same", perturb=0.05)

    def build(tda):
        pc = tda.Posterior(prior, tda.AdaptiveGaussianLogLike(yc, 0.01 * np.eye(m)), LinearModel(Gc))
        pf = tda.Posterior(prior, tda.GaussianLogLike(yf, 0.01 * np.eye(m)), LinearModel(Gf))
        prop = tda.GaussianRandomWalk(C=0.01 * np.eye(d))
        return [pc, pf], prop, dict(subchain_length=1, adaptive_error_model="state-dependent")
    return dict(build=build, n_chains=3, iterations=120, seed=63, prior=prior)


@case
def da_sdaem_pcn():
    """State-dependent error model with the non-symmetric pCN proposal: the transition densities
    CrankNicolson.get_q (proposal.py:364-369) enter the second-stage acceptance; J = 3."""
    rng = np.random.default_rng(64)
    d, m = 6, 9
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d, 0.4))
    (Gc, yc), (Gf, yf) = _linear_levels(rng, d, [m, m], 0.15, prior, coarse_mode="same", perturb=0.04)
    B = rng.standard_normal((m, m))
    cov = 0.0225 * (np.eye(m) + 0.03 * (B @ B.T))

    def build(tda):
        pc = tda.Posterior(prior, tda.AdaptiveGaussianLogLike(yc, cov), LinearModel(Gc))
        pf = tda.Posterior(prior, tda.GaussianLogLike(yf, cov), LinearModel(Gf))
        prop = tda.CrankNicolson(scaling=0.15, adaptive=True, period=30)
        return [pc, pf], prop, dict(subchain_length=3, adaptive_error_model="state-dependent")
    return dict(build=build, n_chains=3, iterations=70, seed=65, prior=prior)


@case
def da_randomize_pcn():
    """DA with randomize_subchain_length (chain.py:369, 525-527): a uniformly chosen link of the
    coarse subchain is promoted; J = 5."""
    rng = np.random.default_rng(66)
    d = 8
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d))
    (Gc, yc), (Gf, yf) = _linear_levels(rng, d, [8, 32], 0.1, prior)

    def build(tda):
        pc = tda.Posterior(prior, tda.GaussianLogLike(yc, 0.01 * np.eye(8)), LinearModel(Gc))
        pf = tda.Posterior(prior, tda.GaussianLogLike(yf, 0.01 * np.eye(32)), LinearModel(Gf))
        return [pc, pf], tda.CrankNicolson(scaling=0.1), dict(subchain_length=5, randomize_subchain_length=True)
    return dict(build=build, n_chains=4, iterations=80, seed=67, prior=prior)


@case
def da_randomize_aem():
    """randomize_subchain_length together with the state-independent error model and an
    adaptively scaled random walk (the promoted link re-enters the coarse chain and is the one
    the bias update and the re-scoring see)."""
    rng = np.random.default_rng(68)
    d, m = 4, 10
    prior = stats.multivariate_normal(0.1 * np.ones(d), np.eye(d))
    (Gc, yc), (Gf, yf) = _linear_levels(rng, d, [m, m], 0.1, prior, coarse_mode="same", perturb=0.05)

    def build(tda):
        pc = tda.Posterior(prior, tda.AdaptiveGaussianLogLike(yc, 0.01 * np.eye(m)), LinearModel(Gc))
        pf = tda.Posterior(prior, tda.GaussianLogLike(yf, 0.01 * np.eye(m)), LinearModel(Gf))
        prop = tda.GaussianRandomWalk(C=0.01 * np.eye(d), adaptive=True, period=20)
        return [pc, pf], prop, dict(subchain_length=4, adaptive_error_model="state-independent",
                                    randomize_subchain_length=True)
    return dict(build=build, n_chains=3, iterations=60, seed=69, prior=prior)


@case
def mlda3_linear():
    """3-level MLDA without error model, J=[3,2], different output sizes per level."""
    rng = np.random.default_rng(7)
    d = 6
    prior = stats.multivariate_normal(np.zeros(d), _exp_cov(d, 0.5))
    lv = _linear_levels(rng, d, [6, 12, 24], 0.15, prior, perturb=0.02)

    def build(tda):
        posts = [tda.Posterior(prior, tda.GaussianLogLike(y, 0.0225 * np.eye(len(y))), LinearModel(G))
                 for G, y in lv]
        prop = tda.GaussianRandomWalk(C=0.03 * np.eye(d), adaptive=True, period=20)
        return posts, prop, dict(subchain_length=[3, 2])
    return dict(build=build, n_chains=3, iterations=40, seed=72, prior=prior)


@case
def mlda3_aem_linear():
    """3-level MLDA + state-independent AEM (bias sums across levels, lazy refresh order)."""
    rng = np.random.default_rng(8)
    d, m = 4, 8
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    lv = _linear_levels(rng, d, [m, m, m], 0.1, prior, coarse_mode="same", perturb=0.04)

    def build(tda):
        posts = []
        for i, (G, y) in enumerate(lv):
            lk = tda.AdaptiveGaussianLogLike(y, 0.01 * np.eye(m)) if i < 2 else tda.GaussianLogLike(y, 0.01 * np.eye(m))
            posts.append(tda.Posterior(prior, lk, LinearModel(G)))
        prop = tda.GaussianRandomWalk(C=0.01 * np.eye(d))
        return posts, prop, dict(subchain_length=[3, 2], adaptive_error_model="state-independent")
    return dict(build=build, n_chains=3, iterations=30, seed=82, prior=prior)


@case
def mlda3_dreamz():
    """3-level MLDA whose base proposal is adaptive DREAM(Z), as in the reference's
    examples/Multilevel Delayed Acceptance.ipynb (there with M0 = 1000 and an LHS archive)."""
    rng = np.random.default_rng(45)
    d = 5
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    lv = _linear_levels(rng, d, [5, 10, 20], 0.2, prior, perturb=0.02)

    def build(tda):
        posts = [tda.Posterior(prior, tda.GaussianLogLike(y, 0.04 * np.eye(len(y))), LinearModel(G))
                 for G, y in lv]
        return posts, tda.DREAMZ(M0=12, delta=1, nCR=3, adaptive=True, period=10), dict(subchain_length=[3, 2])
    return dict(build=build, n_chains=3, iterations=40, seed=46, prior=prior, archive=True)


@case
def mlda4_aem_poisson():
    """cfg4 in miniature: 4-level MLDA + state-independent AEM on the 1-D Poisson inversion."""
    rng = np.random.default_rng(4)
    d, ns, msens = 4, [32, 64, 128, 256], 31
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    truth = prior.rvs(random_state=rng)
    sig = 1e-3
    y = Poisson1D(1024, d, msens)(truth) + sig * rng.standard_normal(msens)
    models = [Poisson1D(n, d, msens) for n in ns]

    def build(tda):
        posts = []
        for i, mdl in enumerate(models):
            lk = (tda.AdaptiveGaussianLogLike(y, sig ** 2 * np.eye(msens)) if i < 3
                  else tda.GaussianLogLike(y, sig ** 2 * np.eye(msens)))
            posts.append(tda.Posterior(prior, lk, mdl))
        prop = tda.GaussianRandomWalk(C=1e-4 * np.eye(d))
        return posts, prop, dict(subchain_length=[3, 2, 2], adaptive_error_model="state-independent")
    return dict(build=build, n_chains=2, iterations=15, seed=92, prior=prior)


@case
def mala_rosenbrock():
    """cfg3: MALA on the 2-D Rosenbrock likelihood (examples/MALA Rosenbrock.ipynb)."""
    prior = stats.multivariate_normal(np.zeros(2), np.eye(2))

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(np.array([0.0]), np.eye(1)), Rosenbrock(1, 10))
        return [post], tda.MALA(scaling=0.01, adaptive=True), {}
    return dict(build=build, n_chains=4, iterations=350, seed=3, prior=prior)


@case
def mala_linear():
    """MALA on a linear-Gaussian problem with a dense prior covariance."""
    rng = np.random.default_rng(13)
    d, m = 4, 9
    prior = stats.multivariate_normal(0.1 * np.ones(d), _exp_cov(d, 0.7))
    (G, y), = _linear_levels(rng, d, [m], 0.3, prior)

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.09 * np.eye(m)), LinearModel(G))
        return [post], tda.MALA(scaling=0.15), {}
    return dict(build=build, n_chains=3, iterations=100, seed=14, prior=prior)


@case
def am_linear():
    """Adaptive Metropolis; the covariance factor is refreshed every `period` steps."""
    rng = np.random.default_rng(15)
    d, m = 4, 10
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    (G, y), = _linear_levels(rng, d, [m], 0.3, prior)

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.09 * np.eye(m)), LinearModel(G))
        return [post], tda.AdaptiveMetropolis(C0=0.05 * np.eye(d), period=20, adaptive=True), {}
    return dict(build=build, n_chains=3, iterations=110, seed=16, prior=prior)


@case
def dreamz_linear():
    """DREAM(Z) with a per-chain archive."""
    rng = np.random.default_rng(17)
    d, m = 4, 10
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    (G, y), = _linear_levels(rng, d, [m], 0.3, prior)

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.09 * np.eye(m)), LinearModel(G))
        return [post], tda.DREAMZ(M0=8, delta=2, nCR=3), {}
    return dict(build=build, n_chains=3, iterations=90, seed=18, prior=prior, archive=True)


@case
def dreamz_adaptive():
    """DREAM(Z) with adaptive=True: global scaling AND the crossover distribution pCR adapt
    (proposal.py:790-809: DeltaCR / LCR updated with the jump normalised by the archive variance)."""
    rng = np.random.default_rng(23)
    d, m = 5, 10
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    (G, y), = _linear_levels(rng, d, [m], 0.3, prior)

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.09 * np.eye(m)), LinearModel(G))
        return [post], tda.DREAMZ(M0=8, delta=1, nCR=3, adaptive=True, period=8), {}
    return dict(build=build, n_chains=3, iterations=200, seed=24, prior=prior, archive=True)


@case
def dream_shared():
    """cfg5 in miniature: DREAM with the archive shared by all chains (lock-step visibility)."""
    rng = np.random.default_rng(5)
    d, m = 6, 12
    prior = stats.multivariate_normal(np.zeros(d), np.eye(d))
    (G, y), = _linear_levels(rng, d, [m], 0.1, prior)

    def build(tda):
        post = tda.Posterior(prior, tda.GaussianLogLike(y, 0.01 * np.eye(m)), LinearModel(G))
        return [post], tda.DREAM(M0=5, delta=1, nCR=3), {}
    return dict(build=build, n_chains=4, iterations=70, seed=19, prior=prior, archive=True, shared=True)


def stream_sizes(spec, iterations):
    """Upper bounds for the number of normals / uniforms one chain consumes."""
    d = spec["d"]
    base_steps = iterations * int(np.prod(spec["J"])) if spec["J"] else iterations
    kind = spec["proposal"]["kind"]
    nz = base_steps * d
    nu = base_steps
    mult = 1
    for j in reversed(spec["J"]):      # upper-level accept tests
        nu += iterations * mult
        mult *= j
    if spec.get("randomize"):          # one index draw per fine iteration
        nu += iterations
    if kind in (4, 5):
        nu += base_steps * (2 * spec["proposal"]["delta"] + 1 + d + 1 + d)
    k = int(spec["proposal"].get("mtm_k", 0))
    if k:
        nz = base_steps * d * (2 * k - 1)
        nu += base_steps
    return nz + 8, nu + 8
