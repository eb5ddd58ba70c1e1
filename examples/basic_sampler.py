"""The reference's "Basic Sampler" example (examples/Basic Sampler.ipynb, README.md:100-125) with the
import swapped: linear regression y = b + m x, Gaussian prior, Gaussian likelihood, adaptive random
walk Metropolis-Hastings.  Needs a B200-class GPU; there is no CPU path.

    python examples/basic_sampler.py [n_chains] [iterations]
"""
import sys
import os

import numpy as np
import scipy.stats as stats

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyda_b200 as tda          # import tinyDA as tda

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
burnin = iterations // 6

rng = np.random.default_rng(1)
x = np.linspace(0, 1, 100)
y = 1 + 2 * x + 0.2 * rng.standard_normal(100)

prior = stats.multivariate_normal(np.zeros(2), np.eye(2))
loglike = tda.GaussianLogLike(y, 0.04 * np.eye(100))
# the forward model must live on the device: theta -> theta[0] + theta[1] * x
model = tda.LinearModel(np.stack([np.ones_like(x), x], axis=1))
posterior = tda.Posterior(prior, loglike, model)
proposal = tda.GaussianRandomWalk(C=np.eye(2), scaling=0.1, adaptive=True)

chains = tda.sample(posterior, proposal, iterations=iterations, n_chains=n_chains, seed=1)

samples = tda.get_samples(chains, "parameters", burnin=burnin)
pooled = np.concatenate([samples["chain_%d" % i] for i in range(n_chains)])
print("posterior mean  b = %.3f  m = %.3f   (truth 1, 2)" % tuple(pooled.mean(axis=0)))
print("posterior sd    b = %.3f  m = %.3f" % tuple(pooled.std(axis=0)))
par = np.stack([samples["chain_%d" % i] for i in range(min(n_chains, 64))])
print("bulk ESS (first %d chains): b %.0f, m %.0f;  R-hat b %.3f, m %.3f"
      % (par.shape[0], tda.ess_bulk(par[:, :, 0]), tda.ess_bulk(par[:, :, 1]), tda.rhat(par[:, :, 0]), tda.rhat(par[:, :, 1])))
assert abs(pooled.mean(axis=0)[0] - 1) < 0.2 and abs(pooled.mean(axis=0)[1] - 2) < 0.3
