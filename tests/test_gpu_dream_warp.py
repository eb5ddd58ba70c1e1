"""The warp-per-chain DREAM(Z) / DREAM kernel (csrc/tda_dream_warp.cu, kernel "dreamw") against the reference's
trajectories (golden fixtures), against the lock-step generic kernel on identical Philox streams, and its
own invariants (launch cuts, several chains per warp, model-output records).

Reference: proposal.py:608-852 (DREAMZ), :1627-1656 + ray.py:366-384 (DREAM, shared archive), chain.py:78-129."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu


def _cfg5(C, seed=3):
    from tinyda_b200 import lower_problem, workloads
    w = workloads.cfg5_dream()
    spec = lower_problem(w["posteriors"], w["proposal"])
    rng = np.random.default_rng(seed)
    theta0 = w["prior"].rvs(C, random_state=rng)
    archive0 = w["prior"].rvs(C * 16, random_state=rng).reshape(C, 16, 32)
    return spec, theta0, archive0


def _run(spec, theta0, archive0, iters, kernel, dtype, store, cuts=None, seed=11):
    from tinyda_b200.engine import Engine
    eng = Engine(spec, theta0.shape[0], dtype=dtype, seed=seed, store=store, capacity_iterations=iters, archive0=archive0)
    eng.select_kernel(kernel)
    assert eng.kernel() == kernel
    eng.init(theta0)
    for n in (cuts or [iters]):
        eng.run(n)
    eng.sync()
    out = dict(theta=eng.fetch(0, "theta"), like=eng.fetch(0, "like"), prior=eng.fetch(0, "prior"), acc=eng.fetch(0, "accept"),
               cursors=eng.get("cursors"), moments=eng.get("moments"), state=eng.get("theta", 0))
    if store & 4:
        out["F"] = eng.fetch(0, "output")
    eng.close()
    return out


@pytest.mark.parametrize("name", ["dreamz_linear", "dream_shared"])
def test_fixtures_select_the_warp_kernel_and_match_the_reference(name):
    """The fixtures' DREAMZ / DREAM problems run on the warp kernel by default; float64 trajectories equal the
    unmodified reference's under injected streams (decisions identical, states to 1e-10), model outputs included."""
    from gpu_util import run_engine
    g = golden_io.load(name)
    out, eng = run_engine(g, "float64", store_F=True)
    assert eng.kernel() == "dreamw"
    ref = g["ref"][0]
    assert np.array_equal(out[0]["acc"], ref["acc"])
    np.testing.assert_allclose(out[0]["theta"], ref["theta"], rtol=1e-10, atol=1e-10 * np.abs(ref["theta"]).max())
    np.testing.assert_allclose(out[0]["like"], ref["like"], rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(out[0]["prior"], ref["prior"], rtol=1e-10, atol=1e-9)
    if "F" in ref:
        np.testing.assert_allclose(out[0]["F"], ref["F"], rtol=1e-9, atol=1e-9 * np.abs(ref["F"]).max())
    assert np.array_equal(eng.get("cursors").T, g["consumed"])
    eng.close()


@pytest.mark.parametrize("C", [96, 4000])
def test_warp_kernel_equals_the_lockstep_kernel_on_philox_streams_fp64(C):
    """BASELINE cfg5's shape (d = 32, 256 observations, M0 = 16, shared archive), engine-generated Philox streams:
    same decisions, same stream consumption, states to 1e-10 -- with one chain per warp (C = 96) and with
    several chains per warp (C = 4000 in float64: 148 CTAs x 8 warps)."""
    from tinyda_b200.engine import STORE_FULL
    spec, theta0, archive0 = _cfg5(C)
    iters = 30
    a = _run(spec, theta0, archive0, iters, "dreamw", "float64", STORE_FULL)
    b = _run(spec, theta0, archive0, iters, "generic", "float64", STORE_FULL)
    assert np.array_equal(a["acc"], b["acc"])
    assert np.array_equal(a["cursors"], b["cursors"])
    np.testing.assert_allclose(a["theta"], b["theta"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(a["like"], b["like"], rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(a["F"], b["F"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(a["moments"], b["moments"], rtol=1e-9, atol=1e-9)
    assert 0.02 < a["acc"][1:].mean() < 0.98


def test_warp_kernel_fp32_agrees_with_the_lockstep_kernel_until_a_near_tie():
    from gpu_util import first_divergence
    from tinyda_b200.engine import STORE_FULL
    spec, theta0, archive0 = _cfg5(512)
    iters = 40
    a = _run(spec, theta0, archive0, iters, "dreamw", "float32", STORE_FULL & ~4)
    b = _run(spec, theta0, archive0, iters, "generic", "float32", STORE_FULL & ~4)
    fd = first_divergence(a["acc"].T.astype(bool), b["acc"].T.astype(bool))
    # the archive is shared: one flipped near-tie reaches other chains through the rows they draw later, so the
    # comparison stops at the ensemble's first flip
    k = int(fd.min())
    print("\ndreamw vs generic (float32, 512 chains x %d steps): first flipped decision at record %d" % (iters, k))
    assert k >= 10
    sc = np.abs(b["theta"][:k]).max()
    assert np.abs(a["theta"][:k] - b["theta"][:k]).max() / sc <= 1e-5


@pytest.mark.parametrize("name", ["dreamz_linear", "dream_shared"])
def test_launch_cuts_do_not_change_the_chains(name):
    """run(a); run(b) == run(a + b) bit for bit (the archive length, stream cursors and step flags carry over)."""
    from tinyda_b200.engine import STORE_FULL
    if name == "dream_shared":
        spec, theta0, archive0 = _cfg5(3000)
    else:
        g = golden_io.load(name)
        spec = g["spec"]
        C = 300
        theta0 = np.resize(g["theta0"], (C, g["theta0"].shape[1])) + 0.01 * np.arange(C)[:, None]
        archive0 = np.resize(g["archive0"], (C,) + g["archive0"].shape[1:])
        spec = dict(spec)
    iters = 24
    a = _run(spec, theta0, archive0, iters, "dreamw", "float32", STORE_FULL & ~4)
    b = _run(spec, theta0, archive0, iters, "dreamw", "float32", STORE_FULL & ~4, cuts=[1, 7, 16])
    for k in ("theta", "like", "prior", "acc", "cursors", "moments", "state"):
        assert np.array_equal(a[k], b[k]), k


def test_warp_kernel_targets_the_closed_form_posterior():
    """cfg5 is linear-Gaussian: the ensemble's pooled mean / variance against the conjugate posterior."""
    from tinyda_b200 import lower_problem, workloads
    from tinyda_b200.engine import Engine, STORE_STATS
    w = workloads.cfg5_dream()
    spec = lower_problem(w["posteriors"], w["proposal"])
    mu, S = workloads.conjugate_posterior(w["G"], w["y"], w["sigma2"], w["prior"])
    C, burn, iters = 1024, 40000, 2000          # ~9 us per lock-step step: 0.4 s; the archive grows to 5.5 GB
    rng = np.random.default_rng(0)
    theta0 = w["prior"].rvs(C, random_state=rng)
    archive0 = w["prior"].rvs(C * 16, random_state=rng).reshape(C, 16, 32)
    eng = Engine(spec, C, dtype="float32", seed=2, store=STORE_STATS, capacity_iterations=iters, archive0=archive0,
                 archive_iterations=burn + iters)
    assert eng.kernel() == "dreamw"
    eng.init(theta0)
    eng.run(burn, record=False)
    eng.run(iters)
    th = eng.fetch(0, "theta")[1:].astype(np.float64)            # [iters][d][C]
    eng.close()
    sd = np.sqrt(np.diag(S))
    pooled_mean = th.mean(axis=(0, 2))
    pooled_var = th.transpose(1, 0, 2).reshape(32, -1).var(axis=1)
    assert np.abs(pooled_mean - mu).max() < 0.1 * sd.max(), np.abs(pooled_mean - mu).max() / sd.max()
    assert np.abs(pooled_var / np.diag(S) - 1).max() < 0.2, pooled_var / np.diag(S)


@pytest.mark.parametrize("shared", [False, True])
def test_diagonal_likelihood_model_outputs_and_odd_sizes_agree_with_the_lockstep_kernel(shared):
    """d = 5 parameters, 37 observations (neither a multiple of the padding), a diagonal (non-isotropic) likelihood,
    a dense prior covariance, delta = 2 pairs, model outputs recorded: float64, decision for decision."""
    import scipy.stats as stats
    from tinyda_b200 import lower_problem
    from tinyda_b200.distributions import GaussianLogLike
    from tinyda_b200.engine import STORE_FULL
    from tinyda_b200.models import LinearModel
    from tinyda_b200.posterior import Posterior
    from tinyda_b200.proposal import DREAM, DREAMZ
    rng = np.random.default_rng(17)
    d, m, C, M0 = 5, 37, 45, 9
    A = rng.standard_normal((d, d))
    prior = stats.multivariate_normal(0.2 * np.ones(d), A @ A.T / d + 0.5 * np.eye(d))
    G = rng.standard_normal((m, d)) / np.sqrt(d)
    var = 0.02 * (1.0 + rng.random(m))
    y = G @ prior.rvs(random_state=rng) + np.sqrt(var) * rng.standard_normal(m)
    post = Posterior(prior, GaussianLogLike(y, np.diag(var)), LinearModel(G, offset=0.1 * rng.standard_normal(m)))
    prop = (DREAM if shared else DREAMZ)(M0=M0, delta=2, nCR=4)
    spec = lower_problem([post], prop)
    theta0 = prior.rvs(C, random_state=rng)
    archive0 = prior.rvs(C * M0, random_state=rng).reshape(C, M0, d)
    a = _run(spec, theta0, archive0, 35, "dreamw", "float64", STORE_FULL)
    b = _run(spec, theta0, archive0, 35, "generic", "float64", STORE_FULL)
    assert np.array_equal(a["acc"], b["acc"])
    assert np.array_equal(a["cursors"], b["cursors"])
    np.testing.assert_allclose(a["theta"], b["theta"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(a["prior"], b["prior"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(a["like"], b["like"], rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(a["F"], b["F"], rtol=1e-10, atol=1e-12)
    assert 0.02 < a["acc"][1:].mean() < 0.98


@pytest.mark.parametrize("K", [3, 8])
def test_bounded_staleness_sync_every_matches_the_oracle_and_ignores_launch_cuts(K):
    """DREAM(sync_every=K): a chain at step t draws its pairs from all chains' rows through the last multiple of K
    below t.  float64 against the oracle's DreamEnsemble under the same rule (engine streams exported), then launch
    cuts that do not fall on multiples of K bit for bit, and K = 1 against the default."""
    import problems
    from oracle import tinyda_oracle as orc
    from tinyda_b200 import lower_problem, workloads
    from tinyda_b200.engine import Engine, STORE_FULL
    from tinyda_b200.proposal import DREAM
    w = workloads.cfg5_dream()
    spec = lower_problem(w["posteriors"], DREAM(M0=16, delta=1, nCR=3, sync_every=K))
    assert int(spec["proposal"]["sync_every"]) == K
    C, iters = 24, 37
    rng = np.random.default_rng(8)
    theta0 = w["prior"].rvs(C, random_state=rng)
    archive0 = w["prior"].rvs(C * 16, random_state=rng).reshape(C, 16, 32)

    def run(cuts, dtype="float64"):
        eng = Engine(spec, C, dtype=dtype, seed=5, store=STORE_FULL, capacity_iterations=iters, archive0=archive0)
        assert eng.kernel() == "dreamw"
        eng.init(theta0)
        for n in cuts:
            eng.run(n)
        out = (np.transpose(eng.fetch(0, "theta"), (2, 0, 1)), eng.fetch(0, "accept").T.astype(bool), eng.fetch(0, "like").T,
               eng.get("cursors"))
        nz, nu = problems.stream_sizes(spec, iters)
        z, u = eng.fill_streams(nz, nu)
        eng.close()
        return out, z, u

    (th, acc, lk, cur), z, u = run([iters])
    ref, chains = orc.run_chains(spec, theta0, z, u, iters, archive0)
    assert np.array_equal(acc, ref[0]["acc"])
    np.testing.assert_allclose(th, ref[0]["theta"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(lk, ref[0]["like"], rtol=1e-9, atol=1e-8)
    assert 0.02 < acc[:, 1:].mean() < 0.98
    (th2, acc2, lk2, cur2), _, _ = run([5, 1, 14, 17])
    assert np.array_equal(th, th2) and np.array_equal(acc, acc2) and np.array_equal(cur, cur2)
    # the rule matters: the lock-step run (sync_every = 1) is a different trajectory
    spec1 = lower_problem(w["posteriors"], DREAM(M0=16, delta=1, nCR=3))
    e1 = Engine(spec1, C, dtype="float64", seed=5, store=STORE_FULL, capacity_iterations=iters, archive0=archive0)
    e1.init(theta0)
    e1.run(iters)
    th1 = np.transpose(e1.fetch(0, "theta"), (2, 0, 1))
    e1.close()
    assert not np.array_equal(th1, th)
    assert np.array_equal(th1[:, :2], th[:, :2])                  # the first step sees the initial rows under both rules


def test_sync_every_needs_the_warp_kernel():
    from tinyda_b200 import lower_problem, workloads
    from tinyda_b200._lib import EngineError
    from tinyda_b200.engine import Engine, STORE_STATS
    from tinyda_b200.proposal import DREAM
    w = workloads.cfg5_dream()
    spec = lower_problem(w["posteriors"], DREAM(M0=16, delta=1, nCR=3, sync_every=4))
    rng = np.random.default_rng(1)
    eng = Engine(spec, 8, dtype="float32", seed=1, store=STORE_STATS, capacity_iterations=4,
                 archive0=w["prior"].rvs(8 * 16, random_state=rng).reshape(8, 16, 32))
    eng.select_kernel("generic")
    eng.init(w["prior"].rvs(8, random_state=rng))
    with pytest.raises(EngineError, match="sync_every"):
        eng.run(4)
    eng.close()
