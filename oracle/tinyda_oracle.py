"""CPU oracle for the batched-MCMC hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A plain NumPy restatement of the reference's per-chain algorithms (tinyDA, pure Python).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module, and only as the checker / the timed CPU baseline.  The product package
(tinyda_b200/) never imports it and has no CPU path.

Pinning: the reference ships no tests or golden vectors, so parity is
pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/make_golden.py runs the
unmodified reference (imported from /root/reference, with ray/arviz/xarray stubbed) under
injected normal/uniform streams and commits its trajectories as tests/golden/*.npz;
tests/test_oracle_golden.py checks this oracle against every one of them.

Everything is float64, one chain at a time, written as an explicit state machine (no object
identity, no recursion through Python objects) because that is the form the CUDA kernels use.

Reference map (file:line are into /root/reference/tinyDA/):
  log_prior            scipy.stats.multivariate_normal_frozen.logpdf, call site posterior.py:92
  eval_model           posterior.py:95 (user model; here the device-resident model classes)
  LogLike.*            distributions.py:246-449
  Moments              utils.py:9-124 (RecursiveSampleMoments)
  ZeroMeanMoments      utils.py:127-201 (ZeroMeanRecursiveSampleMoments, state-dependent AEM)
  ChainOracle.base_step    chain.py:101-125 / chain.py:415-444 / proposal.py:1583-1613
  ChainOracle.upper_step   chain.py:353-402 / chain.py:708-765 / proposal.py:1511-1578
  ChainOracle._align       proposal.py:1469-1493 (object identity replaced by saved versions)
  ChainOracle._alpha_state_dependent   chain.py:446-473, distributions.py:427-446,
                                       proposal.py:364-369 (CrankNicolson.get_q)
  randomize_subchain_length            chain.py:310-321, :369-375, :525-527
  ChainOracle._mtm_propose / _mtm_acceptance   MultipleTry ray.py:213-354
  proposals            proposal.py:132-258 (RWMH), 261-369 (pCN), 372-512 (AM),
                       64-131 (IndependenceSampler), 515-605 (operator-weighted pCN), 608-852 (DREAMZ), 861-1005 (MALA),
                       1627-1656 + ray.py:366-384 (DREAM)
"""
import numpy as np

# ---- kinds (kept numerically identical to include/tinyda_b200.h) -------------------------
PROP_RWMH, PROP_PCN, PROP_AM, PROP_MALA, PROP_DREAMZ, PROP_DREAM, PROP_OWPCN, PROP_INDEP = 0, 1, 2, 3, 4, 5, 6, 7
LIK_ISO, LIK_DIAG, LIK_DENSE, LIK_ADAPTIVE = 0, 1, 2, 3
MODEL_LINEAR, MODEL_ROSENBROCK, MODEL_POISSON1D = 0, 1, 2


# ---- leaf numerics -----------------------------------------------------------------------
def log_prior(prior, theta):
    """scipy multivariate_normal logpdf: -0.5*(rank*log(2pi) + log_pdet + |(x-mu) @ LP|^2).
    prior = dict(mean, LP, logconst) with logconst = rank*log(2pi) + log_pdet."""
    w = (theta - prior["mean"]) @ prior["LP"]
    return -0.5 * (prior["logconst"] + np.sum(np.square(w)))


def eval_model(model, theta):
    kind = model["kind"]
    if kind == MODEL_LINEAR:
        # A is G^T, [d][m]
        return theta @ model["A"] + model["b"]
    if kind == MODEL_ROSENBROCK:
        a, b = model["scalars"][0], model["scalars"][1]
        x, y = theta[0], theta[1]
        return np.array([(a - x) ** 2 + b * (y - x ** 2) ** 2])
    if kind == MODEL_POISSON1D:
        Phi_t = model["A"]                # [d][n]
        n = Phi_t.shape[1]
        stride = int(model["scalars"][0])
        m = model["m"]
        h2 = 1.0 / (n * n)
        k = np.exp(theta @ Phi_t)
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
        return u[stride - 1::stride][:m].copy()
    raise ValueError(kind)


def model_gradient(model, theta, sens):
    """model.gradient(theta, sensitivity) = J_F(theta)^T sensitivity (proposal.py:996-998)."""
    kind = model["kind"]
    if kind == MODEL_LINEAR:
        return model["A"] @ sens
    if kind == MODEL_ROSENBROCK:
        a, b = model["scalars"][0], model["scalars"][1]
        x, y = theta[0], theta[1]
        dFdx = -2.0 * (a - x) - 4.0 * b * x * (y - x ** 2)
        dFdy = 2.0 * b * (y - x ** 2)
        return sens[0] * np.array([dFdx, dFdy])
    raise ValueError("no analytic gradient for model kind %d" % kind)


class LogLike:
    """One level's Gaussian log-likelihood incl. the mutable AEM state
    (distributions.py:246-449)."""

    def __init__(self, lik):
        self.kind = lik["kind"]
        self.data = np.asarray(lik["data"], dtype=np.float64)
        m = self.data.shape[0]
        if self.kind == LIK_ISO:
            self.var = float(lik["var"])
        elif self.kind == LIK_DIAG:
            self.var = np.asarray(lik["var"], dtype=np.float64)
        else:
            self.cov = np.asarray(lik["cov"], dtype=np.float64)
            self.cov_inverse = np.linalg.inv(self.cov)          # distributions.py:280
        self.bias = np.zeros(m)                                  # distributions.py:383

    def loglike(self, F):
        if self.kind == LIK_ISO:                                 # distributions.py:326
            return -0.5 * np.linalg.norm(F - self.data) ** 2 / self.var
        if self.kind == LIK_DIAG:                                # distributions.py:312
            return -0.5 * ((F - self.data) ** 2 / self.var).sum()
        if self.kind == LIK_DENSE:                               # distributions.py:296-298
            r = F - self.data
            return -0.5 * (r @ self.cov_inverse @ r)
        r = F + self.bias - self.data                            # distributions.py:419-425
        return -0.5 * (r @ self.cov_inverse @ r)

    def set_bias(self, mu, sigma):                               # distributions.py:385-402
        self.bias = mu.copy()
        if not np.all(sigma < 1e-9):
            self.cov_inverse = np.linalg.inv(self.cov + sigma)

    def loglike_custom_bias(self, F, bias):                      # distributions.py:427-446
        r = F + bias - self.data
        return -0.5 * (r @ self.cov_inverse @ r)

    def grad_loglike(self, F):                                   # distributions.py:300-329,448
        if self.kind == LIK_ISO:
            return 1.0 / self.var * (self.data - F)
        if self.kind == LIK_DIAG:
            return 1.0 / self.var * (self.data - F)
        if self.kind == LIK_DENSE:
            return self.cov_inverse @ (self.data - F)
        return self.cov_inverse @ (self.data - (F + self.bias))


class Moments:
    """RecursiveSampleMoments, utils.py:9-124 (t starts at 1)."""

    def __init__(self, mu0, d, sd=1.0, epsilon=0.0):
        self.mu = np.array(mu0, dtype=np.float64)
        self.sigma = np.zeros((d, d))
        self.t = 1
        self.sd = sd
        self.epsilon = epsilon
        self.d = d

    def update(self, x):                                         # utils.py:113-124
        mu_prev = self.mu.copy()
        t = self.t
        self.mu = (1 / (t + 1)) * (t * mu_prev + x)
        self.sigma = (t - 1) / t * self.sigma + self.sd / t * (
            t * np.outer(mu_prev, mu_prev)
            - (t + 1) * np.outer(self.mu, self.mu)
            + np.outer(x, x)
            + self.epsilon * np.eye(self.d)
        )
        self.t += 1


class ZeroMeanMoments:
    """ZeroMeanRecursiveSampleMoments, utils.py:127-201 (t starts at 1)."""

    def __init__(self, m):
        self.sigma = np.zeros((m, m))
        self.t = 1

    def update(self, x):                                         # utils.py:190-201
        self.sigma = (self.t - 1) / self.t * self.sigma + 1 / self.t * np.outer(x, x)
        self.t += 1


def svd_factor(C):
    """T with np.random.multivariate_normal(0, C) == z @ T (numpy legacy 'svd' method)."""
    _, s, vt = np.linalg.svd(np.atleast_2d(C))
    return np.sqrt(s)[:, None] * vt


class _Stream:
    def __init__(self, z, u):
        self.z = np.asarray(z, dtype=np.float64)
        self.u = np.asarray(u, dtype=np.float64)
        self.nz = 0
        self.nu = 0

    def normals(self, n):
        out = self.z[self.nz:self.nz + n]
        assert out.shape[0] == n, "normal stream exhausted"
        self.nz += n
        return out

    def uniform(self):
        assert self.nu < self.u.shape[0], "uniform stream exhausted"
        v = self.u[self.nu]
        self.nu += 1
        return v

    def uniforms(self, n):
        out = self.u[self.nu:self.nu + n]
        assert out.shape[0] == n, "uniform stream exhausted"
        self.nu += n
        return out


class _State:
    __slots__ = ("theta", "prior", "like", "F", "sid", "grad")

    def __init__(self, theta, prior, like, F, sid, grad=None):
        self.theta, self.prior, self.like, self.F, self.sid, self.grad = theta, prior, like, F, sid, grad

    @property
    def post(self):
        return self.prior + self.like                            # link.py:48

    def copy(self):
        return _State(self.theta, self.prior, self.like, self.F, self.sid, self.grad)


class ChainOracle:
    """One (ML)DA / MH chain.  spec is the plain dict produced by
    tinyda_b200.lowering.lower_problem (or loaded from a golden fixture):

      n_levels, d, J (list, len n_levels-1),
      aem (0 none / 1 state-independent / 2 state-dependent, two levels only),
      randomize (0/1: DAChain randomize_subchain_length, two levels only),
      prior: dict(mean, LP, logconst, cov)
      levels: list of dict(lik=dict(kind, data, var|cov), model=dict(kind, A, b, scalars, m))
      proposal: dict(kind, T, scaling, adaptive, gamma, period, alpha_star,
                     am_sd, am_eps, am_t0, C0, M0, delta, b, b_star, nCR,
                     mtm_k: > 0 wraps the kernel in MultipleTry with k tries)
    """

    def __init__(self, spec, theta0, z, u, archive0=None, am_refactor="svd", svd_per_proposal=False,
                 keep_history=True):
        self.spec = spec
        self.L = int(spec["n_levels"])
        self.d = int(spec["d"])
        self.J = [int(j) for j in spec.get("J", [])]
        self.aem = int(spec.get("aem", 0))
        self.randomize = int(spec.get("randomize", 0))
        self.sub_states = []                                     # coarse links of the running subchain
        self.prior = spec["prior"]
        self.models = [lv["model"] for lv in spec["levels"]]
        self.liks = [LogLike(lv["lik"]) for lv in spec["levels"]]
        self.S = _Stream(z, u)
        self.P = dict(spec["proposal"])
        self.kind = int(self.P["kind"])
        self.am_refactor = am_refactor
        # svd_per_proposal=True re-factors the proposal covariance on every draw exactly like
        # np.random.multivariate_normal does inside the reference (proposal.py:249, :353) -- same
        # result, the reference's cost profile (used by the timed CPU baseline only)
        self.svd_per_proposal = svd_per_proposal
        self.keep_history = keep_history
        L = self.L
        self.next_sid = 1
        theta0 = np.array(theta0, dtype=np.float64)

        # initial links at every level (chain.py:70, 253-260, 624; proposal.py:1379)
        self.cur = [self._create(l, theta0, 0) for l in range(L)]
        self.accepted = [[True] for _ in range(L)]
        # saved[j][a]: latest version of level j's link whose parameters are level a's state
        self.saved = [[None] * L for _ in range(L)]
        for j in range(L):
            for a in range(j + 1, L):
                self.saved[j][a] = self.cur[j].copy()

        # proposal setup (chain.py:74, 264; proposal.py:1435)
        self.scaling = float(self.P.get("scaling", 1.0))
        self.adaptive = bool(self.P.get("adaptive", False))
        self.gamma = float(self.P.get("gamma", 1.01))
        self.period = int(self.P.get("period", 100))
        self.alpha_star = float(self.P.get("alpha_star", 0.24))
        self.k = 0
        self.t = 0
        self.mtm_k = int(self.P.get("mtm_k", 0))
        # engine extension (default off = the reference): the current state joins the reference
        # points, as Liu et al. (2000) prescribe
        self.mtm_include_current = bool(self.P.get("mtm_include_current", 0))
        if self.kind in (PROP_RWMH, PROP_PCN, PROP_AM, PROP_INDEP):
            self.T = np.array(self.P["T"], dtype=np.float64)
        if self.kind == PROP_OWPCN:                              # proposal.py:575-579
            self.state_operator = np.array(self.P["state_operator"], dtype=np.float64)
            self.noise_operator = np.array(self.P["noise_operator"], dtype=np.float64)
            self.T = np.array(self.P["T_prior"], dtype=np.float64)
        if self.kind == PROP_AM:
            self.am = Moments(theta0, self.d, sd=float(self.P["am_sd"]), epsilon=float(self.P["am_eps"]))
            self.am_t0 = int(self.P["am_t0"])
            self.n_refactor = 0
        if self.kind in (PROP_DREAMZ, PROP_DREAM):
            self.Z = np.array(archive0, dtype=np.float64)        # local archive (proposal.py:788)
            self.delta = int(self.P["delta"])
            self.b = float(self.P["b"])
            self.b_star = float(self.P["b_star"])
            self.nCR = int(self.P["nCR"])
            self.pCR = np.array(self.nCR * [1 / self.nCR])
            self.LCR = np.zeros(self.nCR)                        # proposal.py:752-754
            self.DeltaCR = np.ones(self.nCR)
            self.shared_view = None                              # set by DreamEnsemble
        if self.kind == PROP_MALA:
            self.prior_cov_inv = np.linalg.inv(self.prior["cov"])    # utils.py:275
            self.cur[0].grad = self._gradient(self.cur[0])

        # adaptive error model setup (chain.py:272-305, 644-678; proposal.py:1408-1467)
        if self.aem and L > 1:
            m = self.cur[0].F.shape[0]
            self.model_diff = [None] * L
            self.bias = [None] * L
            for l in range(1, L):
                self.model_diff[l] = self.cur[l].F - self.cur[l - 1].F
                self.bias[l] = ZeroMeanMoments(m) if self.aem == 2 else Moments(self.model_diff[l], m)
            for l in range(L - 1, 0, -1):
                self._push_bias(l)

        # histories: one record per LOCAL step per level (+ the initial link at the top level)
        self.hist = [dict(theta=[], prior=[], like=[], F=[], acc=[]) for _ in range(L)]
        self._record(L - 1, True)

    # -- links -----------------------------------------------------------------------------
    def _create(self, level, theta, sid):                        # posterior.py:78-110
        prior = log_prior(self.prior, theta)
        F = eval_model(self.models[level], theta)
        like = self.liks[level].loglike(F)
        return _State(theta, prior, like, F, sid)

    def _record(self, level, acc):
        if not self.keep_history:
            return
        h = self.hist[level]
        c = self.cur[level]
        h["theta"].append(c.theta.copy())
        h["prior"].append(c.prior)
        h["like"].append(c.like)
        h["F"].append(c.F.copy())
        h["acc"].append(bool(acc))

    # -- AEM -------------------------------------------------------------------------------
    def _push_bias(self, l):
        """Level l pushes the bias moments into level l-1's likelihood and re-scores level
        l-1's last link.  Top level: own bias only (chain.py:659-661, 753-756); lower
        levels: sums over all finer biases, current values (proposal.py:1454-1458, 1563-1569)."""
        L = self.L
        if self.aem == 2:                                        # chain.py:296-300, 517-522
            mu, sigma = self.model_diff[l], self.bias[l].sigma
        elif l == L - 1:
            mu, sigma = self.bias[l].mu, self.bias[l].sigma
        else:
            mu = np.sum([self.bias[k].mu for k in range(l, L)], axis=0)
            sigma = np.sum([self.bias[k].sigma for k in range(l, L)], axis=0)
        self.liks[l - 1].set_bias(mu, sigma)
        c = self.cur[l - 1]
        new = c.copy()                                           # posterior.py:112-134
        new.like = self.liks[l - 1].loglike(c.F)
        self.cur[l - 1] = new
        for a in range(l, L):                                    # later alignments find this version
            if self.cur[a].sid == new.sid:
                self.saved[l - 1][a] = new.copy()

    # -- proposals -------------------------------------------------------------------------
    def _gradient(self, st):                                     # proposal.py:990-1000
        g_prior = self.prior_cov_inv @ (self.prior["mean"] - st.theta)   # utils.py:272-278
        sens = self.liks[0].grad_loglike(st.F)                   # utils.py:283-285
        return g_prior + model_gradient(self.models[0], st.theta, sens)

    def _get_q(self, x, y):                                      # proposal.py:977-988
        s = self.scaling
        return -0.5 / s ** 2 * np.linalg.norm(x.theta - y.theta - 0.5 * s ** 2 * y.grad) ** 2

    def _logsumexp(self, a):
        """scipy.special.logsumexp for a 1-D array (ray.py:316, :350-352)."""
        a = np.asarray(a, dtype=np.float64)
        if a.size == 0:
            return -np.inf
        a_max = a.max()
        if not np.isfinite(a_max):
            a_max = 0.0
        with np.errstate(divide="ignore"):
            return np.log(np.sum(np.exp(a - a_max))) + a_max

    def _mtm_propose(self):                                      # ray.py:279-317
        c = self.cur[0]
        k = self.mtm_k
        links = [self._create(0, self._propose_from(c), self.next_sid) for _ in range(k)]
        if self.kind in (PROP_RWMH, PROP_AM):                    # kernel.is_symmetric -> MTM(II)
            q = np.zeros(k)
        elif self.mtm_include_current:                           # MTM(I) of Liu et al.: w(y, x) = pi(y) T(y, x)
            q = np.array([self._get_q_pcn(l, c) for l in links])
        else:                                                    # the reference: pi(y) T(x, y), ray.py:292-296
            q = np.array([self._get_q_pcn(c, l) for l in links])
        w = np.array([l.post + qi for l, qi in zip(links, q)])
        w[np.isnan(w)] = -np.inf
        self.mtm_weights = w
        u = self.S.uniform()
        if np.isinf(w).all():
            idx = min(int(np.floor(u * k)), k - 1)
        else:
            with np.errstate(over="ignore"):
                pr = np.exp(w - self._logsumexp(w))
            idx = int(min(np.searchsorted(np.cumsum(pr), u, side="right"), k - 1))
        return links[idx].theta

    def _mtm_acceptance(self, new, old):                         # ray.py:319-354
        if np.isnan(new.post) or np.isinf(self.mtm_weights).all():
            return 0.0
        refs = [self._create(0, self._propose_from(new), 0) for _ in range(self.mtm_k - 1)]
        if self.kind in (PROP_RWMH, PROP_AM):
            q = np.zeros(len(refs))
        elif self.mtm_include_current:
            q = np.array([self._get_q_pcn(r, new) for r in refs])
        else:
            q = np.array([self._get_q_pcn(new, r) for r in refs])
        wr = np.array([r.post + qi for r, qi in zip(refs, q)])
        if self.mtm_include_current:
            q_old = 0.0 if self.kind in (PROP_RWMH, PROP_AM) else self._get_q_pcn(old, new)
            wr = np.append(wr, old.post + q_old)
        wr[np.isnan(wr)] = -np.inf
        with np.errstate(over="ignore", invalid="ignore"):
            return np.exp(self._logsumexp(self.mtm_weights) - self._logsumexp(wr))

    def _propose(self):
        if self.mtm_k:
            return self._mtm_propose()
        return self._propose_from(self.cur[0])

    def _propose_from(self, c):
        d = self.d
        if self.kind in (PROP_RWMH, PROP_AM):                    # proposal.py:247-251
            T = svd_factor(self.P["C"]) if (self.svd_per_proposal and self.kind == PROP_RWMH) else self.T
            return c.theta + self.scaling * (self.S.normals(d) @ T)
        if self.kind == PROP_PCN:                                # proposal.py:349-355
            T = svd_factor(self.prior["cov"]) if self.svd_per_proposal else self.T
            return np.sqrt(1 - self.scaling ** 2) * c.theta + self.scaling * (self.S.normals(d) @ T)
        if self.kind == PROP_INDEP:                              # proposal.py:117-119: q.rvs(1)
            return np.asarray(self.P["q_mean"], dtype=np.float64) + self.S.normals(d) @ self.T
        if self.kind == PROP_OWPCN:                              # proposal.py:593-598
            return np.dot(self.state_operator, c.theta) + np.dot(self.noise_operator, self.S.normals(d) @ self.T)
        if self.kind == PROP_MALA:                               # proposal.py:948-959
            return c.theta + 0.5 * self.scaling ** 2 * c.grad + self.scaling * self.S.normals(d)
        # DREAMZ / DREAM                                         # proposal.py:811-852
        M, rowfn = self._archive()
        Z_r1 = np.zeros(d)
        Z_r2 = np.zeros(d)
        for _ in range(self.delta):
            u1, u2 = self.S.uniforms(2)
            r1 = min(int(np.floor(u1 * M)), M - 1)
            r2 = min(int(np.floor(u2 * (M - 1))), M - 2)
            if r2 >= r1:
                r2 += 1
            Z_r1 += rowfn(r1)
            Z_r2 += rowfn(r2)
        ucr = self.S.uniform()
        self.mCR = int(min(np.searchsorted(np.cumsum(self.pCR), ucr, side="right"), self.nCR - 1))
        CR = (self.mCR + 1) / self.nCR
        draw = self.S.uniforms(d)
        ind = np.zeros(d)
        ind[draw < CR] = 1
        if ind.sum() == 0:
            ind[min(int(np.floor(self.S.uniform() * d)), d - 1)] = 1
        gamma_dream = self.scaling * 2.38 / np.sqrt(2 * self.delta * ind.sum())
        e = -self.b + (self.b - (-self.b)) * self.S.uniforms(d)
        eps = 0.0 + self.b_star * self.S.normals(d)
        return c.theta + ind * ((np.ones(d) + e) * gamma_dream * (Z_r1 - Z_r2) + eps)

    def _archive(self):
        if self.kind == PROP_DREAM and self.shared_view is not None:
            return self.shared_view()
        Z = self.Z
        return Z.shape[0], (lambda r: Z[r, :])

    def _acceptance(self, new, old):
        if self.mtm_k:
            return self._mtm_acceptance(new, old)
        if self.kind == PROP_INDEP:                              # proposal.py:121-131 (no NaN guard there)
            mu, LP = np.asarray(self.P["q_mean"]), np.asarray(self.P["S"])
            logq = lambda st: -0.5 * np.sum(np.square((st.theta - mu) @ LP))     # constant cancels
            with np.errstate(over="ignore", invalid="ignore"):
                return np.exp(new.post - old.post + logq(old) - logq(new))
        if np.isnan(new.post):                                   # proposal.py:254, 358, 963
            return 0.0
        with np.errstate(over="ignore"):
            if self.kind in (PROP_PCN, PROP_OWPCN):
                return np.exp(new.like - old.like)               # proposal.py:362
            if self.kind == PROP_MALA:
                new.grad = self._gradient(new)                   # proposal.py:968-969
                q_x_y = self._get_q(old, new)
                q_y_x = self._get_q(new, old)
                return np.exp(new.post - old.post + q_x_y - q_y_x)   # proposal.py:975
            return np.exp(new.post - old.post)                   # proposal.py:258

    def _adapt(self, theta_cur, theta_prev):
        self.t += 1                                              # proposal.py:228-245
        if self.adaptive and self.t % self.period == 0:
            rate = np.mean(self.accepted[0][-self.period:])
            self.scaling = np.exp(np.log(self.scaling) + self.gamma ** -self.k * (rate - self.alpha_star))
            self.k += 1
        if self.kind == PROP_OWPCN and self.adaptive and self.t % self.period == 0:   # proposal.py:581-591
            from scipy.linalg import sqrtm
            B = np.asarray(self.P["B"], dtype=np.float64)
            self.state_operator = np.real(sqrtm(np.eye(self.d) - self.scaling * B))
            self.noise_operator = np.real(sqrtm(self.scaling * B))
        if self.kind == PROP_AM:                                 # proposal.py:502-512
            self.am.update(theta_cur)
            if self.t >= self.am_t0 and self.t % self.period == 0:
                if self.am_refactor == "svd":
                    self.T = svd_factor(self.am.sigma)
                else:
                    self.T = np.linalg.cholesky(self.am.sigma).T
                self.n_refactor += 1
        if self.kind in (PROP_DREAMZ, PROP_DREAM):               # proposal.py:790-795
            self.Z = np.vstack((self.Z, theta_cur))
            if self.adaptive and self.t % self.period == 0:      # proposal.py:797-809
                jump = theta_cur - theta_prev
                self.DeltaCR[self.mCR] = self.DeltaCR[self.mCR] + (jump ** 2 / np.var(self.Z, axis=0)).sum()
                self.LCR[self.mCR] = self.LCR[self.mCR] + 1
                if np.all(self.LCR > 0):
                    mean = self.DeltaCR / self.LCR
                    self.pCR = mean / mean.sum()

    # -- steps -----------------------------------------------------------------------------
    def base_step(self):
        old = self.cur[0]
        theta_p = self._propose()
        new = self._create(0, theta_p, self.next_sid)
        self.next_sid += 1
        alpha = self._acceptance(new, old)
        u = self.S.uniform()                                     # drawn even if alpha is 0 or >= 1
        acc = bool(u < alpha)
        if acc:
            self.cur[0] = new
        self.accepted[0].append(acc)
        self._record(0, acc)
        self.last_alpha = alpha
        self.last_u = u
        if self.randomize:
            self.sub_states.append(self.cur[0])
        self._adapt(self.cur[0].theta, old.theta)

    def _get_q_pcn(self, x, y):
        """CrankNicolson.get_q (proposal.py:364-369) without the normalising constant, which is
        the same in every term of the state-dependent acceptance and cancels exactly."""
        s = self.scaling
        w = (y.theta - np.sqrt(1 - s ** 2) * x.theta) @ self.prior["LP"]
        return -0.5 * np.sum(np.square(w)) / s ** 2

    def _alpha_state_dependent(self, l, new, below, start_below):   # chain.py:446-473
        bias_next = new.F - below.F
        biased_post = start_below.prior + self.liks[l - 1].loglike_custom_bias(start_below.F, bias_next)
        if self.kind in (PROP_RWMH, PROP_AM):                    # is_symmetric, proposal.py:168
            q_x_y = q_y_x = 0.0
        elif self.kind == PROP_PCN:
            q_x_y = self._get_q_pcn(self.cur[l], new)
            q_y_x = self._get_q_pcn(new, self.cur[l])
        else:
            raise NotImplementedError("the reference has no usable get_q for this proposal")
        return np.exp(min(new.post + q_y_x, biased_post + q_x_y)
                      - min(self.cur[l].post + q_x_y, below.post + q_y_x))

    def upper_step(self, l):
        Jb = self.J[l - 1]
        below = self.cur[l - 1]
        if sum(self.accepted[l - 1][-Jb:]) == 0:                 # chain.py:357, 711; proposal.py:1516
            acc = False
        else:
            if self.randomize:                                   # chain.py:369, 525-527
                k = min(int(np.floor(self.S.uniform() * Jb)), Jb - 1)
                below = self.sub_states[-Jb + k]
            new = self._create(l, below.theta, below.sid)
            start_below = self.saved[l - 1][l]
            with np.errstate(over="ignore", invalid="ignore"):
                if self.aem == 2:
                    alpha = self._alpha_state_dependent(l, new, below, start_below)
                else:
                    alpha = np.exp(new.post - self.cur[l].post + start_below.post - below.post)
            u = self.S.uniform()
            acc = bool(u < alpha)
            if acc:
                self.cur[l] = new
                self.cur[l - 1] = below                          # chain.py:383 (the promoted link)
        self.sub_states = []
        self.accepted[l].append(acc)
        self._record(l, acc)
        self._align(l, acc)
        if self.aem == 2:                                        # chain.py:501-522
            corrected = self.cur[l].F - (self.cur[l - 1].F + self.model_diff[l])
            self.model_diff[l] = self.cur[l].F - self.cur[l - 1].F
            self.bias[l].update(corrected)
            self._push_bias(l)
        elif self.aem:
            if acc:
                self.model_diff[l] = self.cur[l].F - self.cur[l - 1].F
            self.bias[l].update(self.model_diff[l])
            self._push_bias(l)

    def _align(self, l, acc):
        """proposal.py:1469-1493 without object identity."""
        for j in range(l - 1, -1, -1):
            if acc:
                self.saved[j][l] = self.cur[j].copy()
            else:
                s = self.saved[j][l].copy()
                s.theta = self.cur[l].theta
                s.sid = self.cur[l].sid
                self.cur[j] = s
                for a in range(j + 1, l):                        # level a now holds level l's state
                    self.saved[j][a] = self.saved[j][l].copy()
            self.accepted[j].append(acc)
        if self.kind == PROP_MALA and not acc:
            # cached gradient lives on the Link object (proposal.py:951-952): recompute is identical
            self.cur[0].grad = self._gradient(self.cur[0])

    def _run_level(self, l, n):
        for _ in range(n):
            if l == 0:
                self.base_step()
            else:
                self._run_level(l - 1, self.J[l - 1])
                self.upper_step(l)

    def run(self, iterations):
        self._run_level(self.L - 1, iterations)

    def history(self, level):
        h = self.hist[level]
        return dict(theta=np.array(h["theta"]), prior=np.array(h["prior"]), like=np.array(h["like"]),
                    F=np.array(h["F"]), acc=np.array(h["acc"], dtype=bool))


class DreamEnsemble:
    """DREAM with the shared archive (proposal.py:1627-1656, ray.py:366-384) under a deterministic
    visibility rule: a chain at step t sees all chains' rows through the last multiple of
    `sync_every` below t (sync_every = 1: through step t-1, the lock-step rule); archive order is
    chain-major as in np.concatenate(shared_archive) (ray.py:381)."""

    def __init__(self, spec, theta0, z, u, archive0):
        C = theta0.shape[0]
        self.chains = [ChainOracle(spec, theta0[c], z[c], u[c], archive0=archive0[c]) for c in range(C)]
        self.blocks = [np.array(archive0[c], dtype=np.float64) for c in range(C)]
        self.sync_every = max(1, int(spec["proposal"].get("sync_every", 1)))
        self.t = 0
        self.visible = self.blocks[0].shape[0]        # rows per chain every chain may draw from
        for ch in self.chains:
            ch.shared_view = self._view

    def _view(self):
        rows = self.visible
        M = rows * len(self.blocks)
        blocks = self.blocks
        return M, (lambda r: blocks[r // rows][r % rows, :])

    def run(self, iterations):
        for _ in range(iterations):
            for ch in self.chains:
                ch.base_step()
            self.blocks = [np.vstack((b, ch.cur[0].theta)) for b, ch in zip(self.blocks, self.chains)]
            self.t += 1
            if self.t % self.sync_every == 0:
                self.visible = self.blocks[0].shape[0]


def run_chains(spec, theta0, z, u, iterations, archive0=None):
    """Convenience driver: independent chains (or a DREAM ensemble).  Returns per-level
    histories stacked over chains: out[level][field] has shape [C, n_records, ...]."""
    C = theta0.shape[0]
    if int(spec["proposal"]["kind"]) == PROP_DREAM:
        ens = DreamEnsemble(spec, theta0, z, u, archive0)
        ens.run(iterations)
        chains = ens.chains
    else:
        chains = []
        for c in range(C):
            ch = ChainOracle(spec, theta0[c], z[c], u[c],
                             archive0=None if archive0 is None else archive0[c])
            ch.run(iterations)
            chains.append(ch)
    out = []
    for l in range(int(spec["n_levels"])):
        hs = [ch.history(l) for ch in chains]
        out.append({k: np.stack([h[k] for h in hs]) for k in hs[0]})
    return out, chains
