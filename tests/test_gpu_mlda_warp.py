"""The warp-per-chain MH / DA / MLDA kernel for the 1-D Poisson model with the state-independent error model
(csrc/tda_mlda_warp.cu, kernel "mldaw") against the reference's trajectories (golden fixture), against the
lock-step generic kernel on identical Philox streams at BASELINE cfg4's real shape, and its own invariants
(launch cuts, hand-over to and from the generic kernel through the shared chain-state buffers).

Reference: chain.py:680-769, proposal.py:1442-1467, :1502-1613, utils.py:113-124, distributions.py:332-449."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu


def _cfg4(C, seed=4, aem=True):
    from tinyda_b200 import lower_problem, workloads
    from tinyda_b200.distributions import GaussianLogLike
    from tinyda_b200.posterior import Posterior
    w = workloads.cfg4_mlda()
    kw = w["kwargs"]
    if aem:
        spec = lower_problem(w["posteriors"], w["proposal"], kw["subchain_length"], kw["adaptive_error_model"])
    else:
        posts = [Posterior(p.prior, GaussianLogLike(p.likelihood.data, 1e-6 * np.eye(p.likelihood.data.size)), p.model)
                 for p in w["posteriors"]]
        spec = lower_problem(posts, w["proposal"], kw["subchain_length"], None)
    theta0 = 0.3 * w["prior"].rvs(C, random_state=np.random.default_rng(seed))
    return spec, theta0


def _run(spec, theta0, plan, dtype, store=None, seed=55):
    """plan: list of (kernel, iterations) launches on one engine."""
    from tinyda_b200.engine import Engine, STORE_FULL
    iters = sum(n for _, n in plan)
    L = spec["n_levels"]
    eng = Engine(spec, theta0.shape[0], dtype=dtype, seed=seed, store=STORE_FULL if store is None else store,
                 capacity_iterations=iters, chain_offset=7)
    eng.init(theta0)
    for kernel, n in plan:
        eng.select_kernel(kernel)
        assert eng.kernel() == kernel
        eng.run(n)
    eng.sync()
    out = dict(cursors=eng.get("cursors"), moments=eng.get("moments"), counts=eng.get("accept_counts"), flags=eng.error_flags())
    for l in range(L):
        out[l] = dict(theta=eng.fetch(l, "theta"), like=eng.fetch(l, "like"), prior=eng.fetch(l, "prior"),
                      acc=eng.fetch(l, "accept"), F=eng.fetch(l, "output"))
    eng.close()
    return out


def test_fixture_selects_the_warp_kernel_and_matches_the_reference():
    """mlda4_aem_poisson (4 levels, AEM, 31 sensors): the float64 engine picks the warp kernel by itself and
    reproduces the unmodified reference decision for decision on every level."""
    from gpu_util import run_engine
    g = golden_io.load("mlda4_aem_poisson")
    out, eng = run_engine(g, "float64", store_F=True)
    assert eng.kernel() == "mldaw"
    for l in range(g["spec"]["n_levels"]):
        ref = g["ref"][l]
        assert np.array_equal(out[l]["acc"], ref["acc"]), "level %d" % l
        np.testing.assert_allclose(out[l]["theta"], ref["theta"], rtol=1e-10, atol=1e-10 * np.abs(ref["theta"]).max())
        np.testing.assert_allclose(out[l]["prior"], ref["prior"], rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(out[l]["like"], ref["like"], rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(out[l]["F"], ref["F"], rtol=1e-9, atol=1e-9 * np.abs(ref["F"]).max())
    assert np.array_equal(eng.get("cursors").T, g["consumed"])
    eng.close()


@pytest.mark.parametrize("aem", [True, False])
def test_warp_kernel_equals_the_lockstep_kernel_at_cfg4_shape_fp64(aem):
    """d = 16, grids 64/128/256/512, 31 sensors, J = [10, 5, 5]: same decisions on all four levels, same stream
    consumption, states to 1e-10, model outputs (closed-form flux sums vs the Thomas sweeps) to 1e-9."""
    spec, theta0 = _cfg4(40, aem=aem)
    a = _run(spec, theta0, [("mldaw", 2)], "float64")
    b = _run(spec, theta0, [("generic", 2)], "float64")
    assert np.array_equal(a["cursors"], b["cursors"])
    assert np.array_equal(a["counts"], b["counts"])
    for l in range(4):
        assert np.array_equal(a[l]["acc"], b[l]["acc"]), "level %d decisions" % l
        np.testing.assert_allclose(a[l]["theta"], b[l]["theta"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(a[l]["F"], b[l]["F"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(a[l]["like"], b[l]["like"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(a[l]["prior"], b[l]["prior"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(a["moments"], b["moments"], rtol=1e-10, atol=1e-12)
    assert a[0]["acc"][1:].mean() > 0.02


def test_hand_over_between_the_kernels_through_the_chain_state():
    """generic -> mldaw -> generic on one engine equals generic all the way: the warp kernel reads and leaves the
    complete chain state (links of every level, saved versions, bias moments, per-chain factors, cursors) in the
    buffers the lock-step kernel uses."""
    spec, theta0 = _cfg4(24)
    a = _run(spec, theta0, [("generic", 1), ("mldaw", 1), ("generic", 1)], "float64")
    b = _run(spec, theta0, [("generic", 3)], "float64")
    assert np.array_equal(a["cursors"], b["cursors"])
    for l in range(4):
        assert np.array_equal(a[l]["acc"], b[l]["acc"]), "level %d decisions" % l
        np.testing.assert_allclose(a[l]["theta"], b[l]["theta"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(a[l]["like"], b[l]["like"], rtol=1e-6, atol=1e-6)


def test_launch_cuts_do_not_change_the_chains():
    spec, theta0 = _cfg4(3000)
    from tinyda_b200.engine import STORE_STATS, STORE_NONE
    store = [STORE_NONE] * 3 + [STORE_STATS]
    from tinyda_b200.engine import Engine
    outs = []
    for cuts in ([3], [1, 2]):
        eng = Engine(spec, 3000, dtype="float32", seed=5, store=store, capacity_iterations=3)
        assert eng.kernel() == "mldaw"
        eng.init(theta0)
        for n in cuts:
            eng.run(n)
        outs.append((eng.fetch(3, "theta"), eng.fetch(3, "like"), eng.fetch(3, "accept"), eng.get("cursors"), eng.get("moments"),
                     eng.get("accept_counts")))
        assert eng.error_flags() == 0
        eng.close()
    for x, y in zip(*outs):
        assert np.array_equal(x, y)
    assert outs[0][2][1:].mean() > 0.02


def test_warp_kernel_fp32_follows_the_fp64_run_until_a_near_tie():
    """float32 engine (cfg4's real shape, its own Philox streams) against the float64 engine fed exactly those
    streams: identical decisions on the coarsest level up to a first near-tie, states within 1e-5 relative before
    it, finest-grid model outputs at rounding level.  (The float32 and float64 engines draw different normals
    from the same counters, hence the export; the lock-step kernel is measured next to the warp kernel.)"""
    import problems
    from gpu_util import first_divergence
    from tinyda_b200.engine import Engine, STORE_FULL
    spec, theta0 = _cfg4(64)
    iters = 2
    nz, nu = problems.stream_sizes(spec, iters)
    got = {}
    for kern in ("mldaw", "generic"):
        e32 = Engine(spec, 64, dtype="float32", seed=55, store=STORE_FULL, capacity_iterations=iters, chain_offset=7)
        e32.select_kernel(kern)
        e32.init(theta0)
        e32.run(iters)
        z, u = e32.fill_streams(nz, nu)
        a = {l: dict(theta=e32.fetch(l, "theta"), acc=e32.fetch(l, "accept"), F=e32.fetch(l, "output")) for l in range(4)}
        e32.close()
        e64 = Engine(spec, 64, dtype="float64", rng="injected", streams=(z, u), store=STORE_FULL, capacity_iterations=iters)
        e64.select_kernel(kern)
        e64.init(theta0)
        e64.run(iters)
        ref = {l: dict(theta=e64.fetch(l, "theta"), acc=e64.fetch(l, "accept"), F=e64.fetch(l, "output")) for l in range(4)}
        e64.close()
        # coarse records unaffected by any flipped decision: a flip in record r of level l >= 1 (one record per
        # S_l = J_0 ... J_{l-1} coarse steps; the finest level's record 0 is the initial link) changes the coarse
        # chain from coarse record (r + 1) S_l on (r S_l on the finest level)
        fd = first_divergence(a[0]["acc"].T.astype(bool), ref[0]["acc"].T.astype(bool))
        n = a[0]["acc"].shape[0]
        S = 1
        for l in range(1, 4):
            S *= int(spec["J"][l - 1])
            fl = first_divergence(a[l]["acc"].T.astype(bool), ref[l]["acc"].T.astype(bool))
            hit = fl < a[l]["acc"].shape[0]
            fd = np.where(hit, np.minimum(fd, (fl + (0 if l == 3 else 1)) * S), fd)
        worst = 0.0
        for c in range(fd.size):
            k = int(fd[c])
            if k:
                sc = np.abs(ref[0]["theta"][:k, :, c]).max() + 1e-30
                worst = max(worst, float(np.abs(a[0]["theta"][:k, :, c] - ref[0]["theta"][:k, :, c]).max() / sc))
        ferr = float(np.abs(a[3]["F"][0] - ref[3]["F"][0]).max() / np.abs(ref[3]["F"][0]).max())
        got[kern] = (int(fd.sum()), fd.size * n, worst, ferr)
        print("\n%s float32 vs float64 (cfg4 shape, 64 chains x 500 coarse steps): %d of %d coarse records before a first "
              "flipped near-tie, max relative state error %.2e, finest-grid model output error %.2e"
              % (kern, got[kern][0], got[kern][1], worst, ferr))
    matched, total, worst, ferr = got["mldaw"]
    assert matched >= 0.5 * total
    assert worst <= 1e-5
    assert ferr <= 2e-5


@pytest.mark.parametrize("levels,aem,msens,d,pcn", [(1, False, 7, 3, False), (2, True, 15, 5, True), (3, True, 3, 6, False),
                                                    (2, False, 31, 2, True)])
def test_other_shapes_agree_with_the_lockstep_kernel(levels, aem, msens, d, pcn):
    """Single-level MH, two-level DA and three-level MLDA on the Poisson model with fewer sensors than lanes, parameter
    counts that are not multiples of four, a dense prior covariance and both base proposals (the matvec path instead
    of the diagonal shortcut): float64, same decisions and states as the lock-step kernel."""
    import scipy.stats as stats
    from tinyda_b200 import lower_problem
    from tinyda_b200.distributions import GaussianLogLike, AdaptiveGaussianLogLike
    from tinyda_b200.models import Poisson1D
    from tinyda_b200.posterior import Posterior
    from tinyda_b200.proposal import CrankNicolson, GaussianRandomWalk
    rng = np.random.default_rng(levels * 10 + msens)
    A = rng.standard_normal((d, d))
    cov = 0.3 * (A @ A.T / d + np.eye(d))
    prior = stats.multivariate_normal(np.zeros(d), cov)
    ns = [(msens + 1) * s for s in (1, 2, 4)][:levels]
    truth = 0.5 * prior.rvs(random_state=rng)
    sig = 2e-3
    y = Poisson1D((msens + 1) * 8, d, msens)(truth) + sig * rng.standard_normal(msens)
    posts = []
    for i, n in enumerate(ns):
        lk = (AdaptiveGaussianLogLike(y, sig ** 2 * np.eye(msens)) if (aem and i < levels - 1)
              else GaussianLogLike(y, sig ** 2 * np.eye(msens)))
        posts.append(Posterior(prior, lk, Poisson1D(n, d, msens)))
    prop = CrankNicolson(scaling=0.08) if pcn else GaussianRandomWalk(C=2e-4 * cov)
    J = [4, 3][:levels - 1]
    spec = lower_problem(posts, prop, J if levels > 1 else None, "state-independent" if aem else None)
    theta0 = 0.3 * np.atleast_2d(prior.rvs(70, random_state=rng)).reshape(70, d)
    iters = 40 if levels == 1 else 6
    a = _run(spec, theta0, [("mldaw", iters)], "float64")
    b = _run(spec, theta0, [("generic", iters)], "float64")
    assert np.array_equal(a["cursors"], b["cursors"])
    for l in range(levels):
        assert np.array_equal(a[l]["acc"], b[l]["acc"]), "level %d decisions" % l
        np.testing.assert_allclose(a[l]["theta"], b[l]["theta"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(a[l]["F"], b[l]["F"], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(a[l]["like"], b[l]["like"], rtol=1e-6, atol=1e-6)
    assert 0.02 < a[0]["acc"][1:].mean() < 0.999


def test_linear_levels_diagonal_likelihood_and_adaptive_pcn_agree_with_the_lockstep_kernel():
    """Two-level DA on linear operators (11 and 23 outputs), diagonal likelihood on the fine level, the error model on
    the coarse one, adaptive pCN (the window ring, crossing several adaptation periods): float64, decision for
    decision, adapted step sizes included."""
    import scipy.stats as stats
    from tinyda_b200 import lower_problem
    from tinyda_b200.distributions import GaussianLogLike, AdaptiveGaussianLogLike
    from tinyda_b200.engine import Engine, STORE_FULL
    from tinyda_b200.models import LinearModel
    from tinyda_b200.posterior import Posterior
    from tinyda_b200.proposal import CrankNicolson
    rng = np.random.default_rng(23)
    d, m, C = 6, 23, 90
    A = rng.standard_normal((d, d))
    prior = stats.multivariate_normal(np.zeros(d), A @ A.T / d + 0.3 * np.eye(d))
    G = rng.standard_normal((m, d)) / np.sqrt(d)
    Gc = G + 0.05 * rng.standard_normal((m, d))
    var = 0.01 * (1.0 + rng.random(m))
    y = G @ prior.rvs(random_state=rng) + np.sqrt(var) * rng.standard_normal(m)
    posts = [Posterior(prior, AdaptiveGaussianLogLike(y, np.diag(var)), LinearModel(Gc)),
             Posterior(prior, GaussianLogLike(y, np.diag(var)), LinearModel(G))]
    prop = CrankNicolson(scaling=0.3, adaptive=True, period=7)
    spec = lower_problem(posts, prop, 3, "state-independent")
    theta0 = prior.rvs(C, random_state=rng)
    outs = {}
    for kern in ("mldaw", "generic"):
        eng = Engine(spec, C, dtype="float64", seed=77, store=STORE_FULL, capacity_iterations=40)
        eng.select_kernel(kern)
        eng.init(theta0)
        eng.run(17)
        eng.run(23)
        outs[kern] = dict(scaling=eng.get("scaling"), cursors=eng.get("cursors"), counts=eng.get("accept_counts"),
                          th=[eng.fetch(l, "theta") for l in range(2)], acc=[eng.fetch(l, "accept") for l in range(2)],
                          like=[eng.fetch(l, "like") for l in range(2)], F=[eng.fetch(l, "output") for l in range(2)])
        eng.close()
    a, b = outs["mldaw"], outs["generic"]
    assert np.array_equal(a["cursors"], b["cursors"]) and np.array_equal(a["counts"], b["counts"])
    np.testing.assert_allclose(a["scaling"], b["scaling"], rtol=1e-12)
    assert np.ptp(a["scaling"]) > 0 and not np.allclose(a["scaling"], 0.3)
    for l in range(2):
        assert np.array_equal(a["acc"][l], b["acc"][l]), "level %d decisions" % l
        np.testing.assert_allclose(a["th"][l], b["th"][l], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(a["F"][l], b["F"][l], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(a["like"][l], b["like"][l], rtol=1e-7, atol=1e-7)
