"""Proposal descriptors (host side).

Same class names, constructor arguments, defaults, validation errors and class attributes as
the reference's proposals (tinyDA/proposal.py:132-258 GaussianRandomWalk, :261-369
CrankNicolson, :372-512 AdaptiveMetropolis, :608-852 DREAMZ, :861-1005 MALA, :1627-1656
DREAM).  The reference objects carry the per-chain mutable state (scaling, k, t, AM moments,
DREAM archive) and are deep-copied per chain (tinyDA/sampler.py:176); here that state lives
in per-chain device arrays and these objects only describe the kernel to run.
``lower(prior)`` yields the POD parameters + constant buffers for the engine.
"""
import numpy as np

PROP_RWMH, PROP_PCN, PROP_AM, PROP_MALA, PROP_DREAMZ, PROP_DREAM, PROP_OWPCN, PROP_INDEP = 0, 1, 2, 3, 4, 5, 6, 7


def svd_factor(C):
    """The matrix T for which ``np.random.multivariate_normal(0, C)`` equals ``z @ T``
    with z ~ N(0, I): numpy's legacy default method factors C by SVD and returns
    z @ (sqrt(s)[:, None] * Vt).  Feeding the engine this factor (instead of a Cholesky
    factor) is what makes trajectories identical to the reference under shared streams."""
    C = np.atleast_2d(np.asarray(C, dtype=np.float64))
    _, s, vt = np.linalg.svd(C)
    return np.sqrt(s)[:, None] * vt


def _check_square(C, name):
    # proposal.py:190-196 / :444-450
    if not isinstance(C, np.ndarray):
        raise TypeError("%s must be a numpy array" % name)
    elif C.ndim == 1:
        if not C.shape[0] == 1:
            raise ValueError("%s must be an NxN array" % name)
    elif not C.shape[0] == C.shape[1]:
        raise ValueError("%s must be an NxN array" % name)


class Proposal:
    is_symmetric = False


class IndependenceSampler(Proposal):
    """Independence sampler (tinyDA/proposal.py:64-131): proposals are draws of a fixed distribution
    q, acceptance exp(pi(y) - pi(x) + log q(x) - log q(y)).  The reference takes any object with
    .rvs() / .logpdf(); the device needs a multivariate normal q (a scipy frozen
    multivariate_normal, or any object that also exposes .mean and .cov)."""

    is_symmetric = False
    adaptive = False
    kind = PROP_INDEP

    def __init__(self, q):
        self.q = q
        try:                                            # proposal.py:102-105
            self.q.logpdf(self.q.rvs(1))
        except AttributeError:
            raise

    def lower(self, prior):
        import scipy.stats as stats
        if not (hasattr(self.q, "mean") and hasattr(self.q, "cov")):
            raise TypeError("the device engine needs a multivariate normal proposal distribution q "
                            "(an object with .mean and .cov); there is no CPU fallback")
        mean = np.atleast_1d(np.asarray(self.q.mean, dtype=np.float64))
        cov = np.atleast_2d(np.asarray(self.q.cov, dtype=np.float64))
        if mean.shape[0] != prior["mean"].shape[0]:
            raise ValueError("q has a different dimension than the prior")
        co = stats.multivariate_normal(mean, cov).cov_object       # scipy's own whitening, as in logpdf
        return dict(kind=self.kind, scaling=1.0, adaptive=False, gamma=1.01, period=100, alpha_star=0.24,
                    T=svd_factor(cov), S=np.ascontiguousarray(co._LP), ow_lambda=mean, q_mean=mean, q_cov=cov)


class GaussianRandomWalk(Proposal):
    is_symmetric = True
    alpha_star = 0.24
    kind = PROP_RWMH

    def __init__(self, C, scaling=1, adaptive=False, gamma=1.01, period=100):
        _check_square(C, "C")
        self.C = C
        self.d = self.C.shape[0]
        self.scaling = scaling
        self.adaptive = adaptive
        self.gamma = gamma
        self.period = period

    def _common(self):
        return dict(kind=self.kind, scaling=float(self.scaling), adaptive=bool(self.adaptive),
                    gamma=float(self.gamma), period=int(self.period),
                    alpha_star=float(self.alpha_star))

    def lower(self, prior):
        out = self._common()
        out["T"] = svd_factor(self.C)
        out["C"] = np.atleast_2d(np.asarray(self.C, dtype=np.float64))
        return out


class CrankNicolson(GaussianRandomWalk):
    is_symmetric = False
    kind = PROP_PCN

    def __init__(self, scaling=0.1, adaptive=False, gamma=1.01, period=100):
        self.scaling = scaling
        self.adaptive = adaptive
        self.gamma = gamma
        self.period = period

    def lower(self, prior):
        out = self._common()
        out["T"] = svd_factor(prior["cov"])          # proposal.py:336-347: C <- prior covariance
        return out


class OperatorWeightedCrankNicolson(CrankNicolson):
    """Operator-weighted pCN (Law 2014), tinyDA/proposal.py:515-605:
    theta' = sqrtm(I - scaling*B) theta + sqrtm(scaling*B) xi,  xi ~ N(0, prior cov).  The two matrix
    square roots are taken once on the host (proposal.py:578-579); with adaptive=True the reference
    re-takes them every period from the adapted step size (proposal.py:581-591) and the device
    follows each chain's step size through the eigen-decomposition of a symmetric B."""

    kind = PROP_OWPCN

    def __init__(self, B, scaling=1.0, adaptive=False, gamma=1.01, period=100):
        self.B = B
        super().__init__(scaling, adaptive, gamma, period)

    def lower(self, prior):
        from scipy.linalg import sqrtm
        d = prior["cov"].shape[0]
        B = np.atleast_2d(np.asarray(self.B, dtype=np.float64))
        T_prior = svd_factor(prior["cov"])
        out = self._common()
        out["B"] = B
        out["T_prior"] = T_prior
        out["state_operator"] = np.real(sqrtm(np.eye(d) - self.scaling * B))      # proposal.py:578-579
        out["noise_operator"] = np.real(sqrtm(self.scaling * B))
        if self.adaptive:
            # the reference re-takes both matrix square roots from every chain's adapted step size
            # (proposal.py:581-591); for a symmetric B = V diag(lam) V^T they are
            # V diag(sqrt(1 - s lam)) V^T and V diag(sqrt(s lam)) V^T, which the device forms per chain
            if not np.allclose(B, B.T, rtol=1e-12, atol=1e-14):
                raise NotImplementedError("adaptive OperatorWeightedCrankNicolson needs a symmetric operator B")
            lam, V = np.linalg.eigh(B)
            out["ow_lambda"] = lam
            out["T"] = T_prior @ V                       # z @ T = V^T xi
            out["S"] = V                                 # theta @ S = V^T theta
            out["S2"] = V.T
        else:
            out["T"] = T_prior @ out["noise_operator"].T     # z @ T = noise_operator @ (z @ T_prior)
            out["S"] = out["state_operator"].T               # theta @ S = state_operator @ theta
        return out


class AdaptiveMetropolis(GaussianRandomWalk):
    kind = PROP_AM

    def __init__(self, C0, sd=None, epsilon=1e-6, t0=0, period=100, adaptive=False, gamma=1.01):
        _check_square(C0, "C0")
        self.C = C0
        self.d = self.C.shape[0]
        self.scaling = 1
        self.sd = sd if sd is not None else min(1, 2.4 ** 2 / self.d)   # proposal.py:465-468
        self.epsilon = epsilon
        self.t0 = t0
        self.period = period
        self.adaptive = adaptive
        self.gamma = gamma

    def lower(self, prior):
        out = self._common()
        out["T"] = svd_factor(self.C)
        out["C0"] = np.atleast_2d(np.asarray(self.C, dtype=np.float64))
        out["am_sd"] = float(self.sd)
        out["am_eps"] = float(self.epsilon)
        out["am_t0"] = int(self.t0)
        return out


class MALA(GaussianRandomWalk):
    is_symmetric = False
    alpha_star = 0.57
    kind = PROP_MALA

    def __init__(self, scaling=0.1, adaptive=False, gamma=1.01, period=100):
        self.scaling = scaling
        self.adaptive = adaptive
        self.gamma = gamma
        self.period = period

    def lower(self, prior):
        return self._common()


class DREAMZ(GaussianRandomWalk):
    kind = PROP_DREAMZ

    def __init__(self, M0, delta=1, b=5e-2, b_star=1e-6, Z_method="random", nCR=3,
                 adaptive=False, gamma=1.01, period=100):
        self.M = M0
        self.scaling = 1
        self.delta = delta
        self.b = b
        self.b_star = b_star
        self.Z_method = Z_method
        self.adaptive = adaptive
        self.nCR = nCR
        self.gamma = gamma
        self.period = period

    def lower(self, prior):
        if self.Z_method not in ("random", "lhs"):
            raise ValueError("Z_method must be 'random' or 'lhs'")
        out = self._common()
        out.update(M0=int(self.M), delta=int(self.delta), b=float(self.b), b_star=float(self.b_star),
                   nCR=int(self.nCR))
        return out

    def initial_archive(self, prior, rng):
        """The chain's initial archive Z (proposal.py:756-788): M0 prior draws, or for
        Z_method='lhs' a Latin hypercube pushed through the prior's marginal normal quantiles (the
        reference's fallback for a multivariate normal prior, which has no .ppf)."""
        import scipy.stats as stats
        d = np.atleast_1d(prior.mean).shape[0]
        if self.Z_method == "lhs":
            Z = stats.qmc.LatinHypercube(d=d, seed=rng).random(n=int(self.M))
            var = np.diag(np.atleast_2d(prior.cov))
            mean = np.atleast_1d(prior.mean)
            for i in range(d):
                Z[:, i] = stats.norm(loc=mean[i], scale=np.sqrt(var[i])).ppf(Z[:, i])
            return Z
        return np.atleast_2d(prior.rvs(int(self.M), random_state=rng)).reshape(int(self.M), -1)


class DREAM(DREAMZ):
    """DREAM(Z) with an archive shared by all chains (proposal.py:1627-1656).  On one GPU the archive is a single
    HBM array; across GPUs every rank holds a replica and the kernel stores each step's new rows into all of them
    over NVLink peer memory (see sampler.py, parallel.connect_dream_peers).

    ``sync_every`` (extension, default 1): how often the chains agree on what the shared archive holds.  The
    reference's archive is a Ray actor that the chains read whenever they get to it -- how stale a chain's view is
    is left to the scheduler.  Here it is deterministic: a chain at step t draws its pairs from all chains' rows
    through the last multiple of ``sync_every`` below t (1: through step t-1, the lock-step rule the golden fixtures
    were generated under).  Larger values trade a bounded staleness for fewer grid / NVLink barriers."""

    kind = PROP_DREAM

    def __init__(self, M0, delta=1, b=5e-2, b_star=1e-6, Z_method="random", nCR=3,
                 adaptive=False, gamma=1.01, period=100, sync_every=1):
        super().__init__(M0, delta, b, b_star, Z_method, nCR, adaptive, gamma, period)
        if int(sync_every) < 1:
            raise ValueError("sync_every must be a positive integer")
        self.sync_every = int(sync_every)

    def lower(self, prior):
        out = super().lower(prior)
        out["sync_every"] = int(getattr(self, "sync_every", 1))
        return out


class MultipleTry(Proposal):
    """Multiple-Try Metropolis (Liu et al. 2000) around another proposal, tinyDA/ray.py:213-354:
    k candidates from the kernel, one chosen with probability proportional to its posterior
    weight, k-1 reference points drawn from the chosen one, acceptance = ratio of the summed
    weights.  MTM(II) for symmetric kernels, MTM(I) with the kernel's transition density
    otherwise.  The reference evaluates the k links on k Ray actors; here the candidates of all
    chains are evaluated by the same lock-step kernel."""

    is_symmetric = True

    def __init__(self, kernel, k, include_current=False):
        import warnings
        self.kernel = kernel
        self.k = k
        # The reference weighs the k candidates against only k-1 reference points (ray.py:325-347);
        # Liu et al.'s MTM also counts the current state among them, and without it the chain is
        # over-dispersed (variance ratio 3.2 / 1.6 / 1.1 for k = 2 / 3 / 5 on a Gaussian target,
        # measured with the unmodified reference).  The default reproduces the reference;
        # include_current=True is the detailed-balance version.
        self.include_current = bool(include_current)
        if self.kernel.adaptive:                                   # ray.py:258-261
            warnings.warn(" Using global adaptive scaling with MultipleTry proposal can be unstable.\n")

    def lower(self, prior):
        out = self.kernel.lower(prior)
        if out["kind"] not in (PROP_RWMH, PROP_AM, PROP_PCN):
            # MTM(I) needs kernel.get_q on plain links: MALA's reads a cached gradient the candidate
            # links do not have, DREAM(Z) inherits one that returns None (proposal.py:40, :977-988)
            raise TypeError("MultipleTry needs a GaussianRandomWalk, AdaptiveMetropolis or CrankNicolson kernel")
        if int(self.k) < 2 or int(self.k) > 16:
            raise ValueError("MultipleTry is lowered for 2 <= k <= 16 tries")
        out["mtm_k"] = int(self.k)
        out["mtm_include_current"] = int(self.include_current)
        return out


def SingleDreamZ(*args, **kwargs):
    import warnings
    warnings.warn(" SingleDreamZ has been deprecated. Please use DREAMZ.")
    return DREAMZ(*args, **kwargs)
