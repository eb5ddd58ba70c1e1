"""Exploration (GPU): accept rates / convergence of cfg2-shaped DA for several pCN step sizes."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyda_b200 import lower_problem
from tinyda_b200.engine import Engine, STORE_NONE
from tinyda_b200.workloads import cfg2_da, conjugate_posterior

for kernel in ("tc", "generic"):
    for (mf, beta, burn, n) in ((256, 0.01, 1500, 3000), (256, 0.02, 1500, 3000), (256, 0.03, 1500, 3000), (1024, 0.004, 3000, 3000)):
        if kernel == "generic" and burn + n > 5000:
            continue
        w = cfg2_da(beta=beta, m_f=mf)
        mu, S = conjugate_posterior(w["G"], w["y"], w["sigma2"], w["prior"])
        sd = np.sqrt(np.diag(S))
        spec = lower_problem(w["posteriors"], w["proposal"], 10)
        C = 8192
        rng = np.random.default_rng(0)
        theta0 = rng.multivariate_normal(mu, S, size=C) + 2.0 * sd
        eng = Engine(spec, C, dtype="float32", seed=3, store=STORE_NONE)
        eng.select_kernel(kernel)
        eng.init(theta0)
        t0 = time.time()
        eng.run(burn); eng.sync()
        a0 = eng.get("accept_counts").astype(float)
        m0 = eng.get("moments")
        eng.run(n); eng.sync()
        dt = time.time() - t0
        a1 = eng.get("accept_counts").astype(float)
        m1 = eng.get("moments")
        cm = ((m1[0] - m0[0]) / n).T
        c2 = ((m1[1] - m0[1]) / n).T
        mcse = cm.std(axis=0, ddof=1) / np.sqrt(C)
        z = (cm.mean(axis=0) - mu) / mcse
        pooled_var = c2.mean(axis=0) - cm.mean(axis=0) ** 2
        print(kernel, "mf", mf, "beta", beta, "rates c/f", (a1[0] - a0[0]).mean() / (n * 10), (a1[1] - a0[1]).mean() / n,
              "max|z|", np.abs(z).max(), "max err mean (in sd)", (np.abs(cm.mean(axis=0) - mu) / sd).max(),
              "var ratio min/max", (pooled_var / np.diag(S)).min(), (pooled_var / np.diag(S)).max(), "%.2fs" % dt, flush=True)
        eng.close()
