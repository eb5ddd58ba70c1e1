"""Two-level Delayed Acceptance with pCN on a linear-Gaussian inverse problem (BASELINE.json
configs[1] shape: 64 parameters, 1024 observations, coarse model = every 8th observation), written
exactly as one would for the reference (tinyDA/sampler.py:21) with the import swapped.

    python examples/delayed_acceptance.py [n_chains] [iterations]
"""
import sys
import os
import time

import numpy as np
import scipy.stats as stats

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyda_b200 as tda          # import tinyDA as tda

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 500

rng = np.random.default_rng(2)
d, m = 64, 1024
xs = np.linspace(0, 1, d)
C = np.exp(-np.abs(xs[:, None] - xs[None, :]) / 0.2)
prior = stats.multivariate_normal(np.zeros(d), C)
G = rng.standard_normal((m, d)) / 8
truth = prior.rvs(random_state=rng)
y = G @ truth + 0.1 * rng.standard_normal(m)

coarse = tda.Posterior(prior, tda.GaussianLogLike(y[::8], 0.01 * np.eye(m // 8)), tda.LinearModel(G[::8]))
fine = tda.Posterior(prior, tda.GaussianLogLike(y, 0.01 * np.eye(m)), tda.LinearModel(G))
proposal = tda.CrankNicolson(scaling=0.05)

t0 = time.perf_counter()
chains = tda.sample([coarse, fine], proposal, iterations=iterations, n_chains=n_chains, subchain_length=10,
                    store_coarse_chain=False, dtype="float32", store_model_output=False, seed=3)
dt = time.perf_counter() - t0
print("%d chains x %d fine iterations (x 10 coarse steps each) in %.2f s = %.3g fine transitions/s incl. set-up and host copy"
      % (n_chains, iterations, dt, n_chains * iterations / dt))
like = np.stack([chains["chain_fine_%d" % i].likelihood for i in range(n_chains)])
acc = np.stack([chains["chain_fine_%d" % i].accepted[1:] for i in range(n_chains)])
print("fine-level acceptance rate %.2f; mean fine log-likelihood: initial %.0f -> after %d iterations %.0f"
      % (acc.mean(), like[:, 0].mean(), iterations, like[:, -1].mean()))
# closed-form posterior of the linear-Gaussian problem, for reference (pCN with a 0.05 step needs a few
# thousand fine iterations from a prior draw to reach it; see tests/test_gpu_sample_api.py for the
# converged 3-MCSE check)
Cinv = np.linalg.inv(C)
S = np.linalg.inv(Cinv + G.T @ G / 0.01)
mu = S @ (G.T @ y / 0.01)
last = np.stack([chains["chain_fine_%d" % i].parameters[-1] for i in range(n_chains)])
first = np.stack([chains["chain_fine_%d" % i].parameters[0] for i in range(n_chains)])
print("RMS distance of the chains to the closed-form posterior mean: %.3f at the start, %.3f at the end"
      % (np.sqrt(((first - mu) ** 2).mean()), np.sqrt(((last - mu) ** 2).mean())))
assert like[:, -1].mean() > like[:, 0].mean()
