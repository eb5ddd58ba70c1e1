"""Compacted finest-level history (tda_compact_*), burn-in runs, the history-capacity check and the
quantity of interest, all through the C ABI."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu


def _cfg2(C, iters, store=None, dtype="float32", kernel=None, seed=3):
    from tinyda_b200 import lower_problem
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    from tinyda_b200.workloads import cfg2_da
    w = cfg2_da()
    spec = lower_problem(w["posteriors"], w["proposal"], 10)
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
    eng = Engine(spec, C, dtype=dtype, seed=seed, store=[STORE_NONE, STORE_STATS] if store is None else store,
                 capacity_iterations=iters)
    if kernel:
        eng.select_kernel(kernel)
    eng.init(theta0)
    return eng, w, spec


@pytest.mark.parametrize("C,kernel", [(512, None), (300, "generic")])
def test_compacted_history_expands_bit_identically_to_the_dense_one(C, kernel):
    """Device compaction (accept flags + accepted rows, chain-major) against the dense fetch of the same
    records: per chain and in bulk, bit for bit; two blocks with a history reset in between."""
    from tinyda_b200.link import CompactHistory
    iters = 24
    eng, w, spec = _cfg2(C, iters, kernel=kernel)
    hist = CompactHistory(C)
    dense_t, dense_l, dense_a = [], [], []
    for block, n in enumerate((iters, 9)):
        if block:
            eng.history_reset()
        eng.run(n)
        nrec = n + (1 if block == 0 else 0)
        dense_t.append(eng.fetch(1, "theta", 0, nrec)); dense_l.append(eng.fetch(1, "like", 0, nrec))
        dense_a.append(eng.fetch(1, "accept", 0, nrec))
        eng.compact_begin(0, nrec, block == 0, ("theta", "stats"), slot=block & 1)
        hist.append(eng.compact_collect(block & 1))
    eng.compact_sync()
    th = np.concatenate(dense_t); lk = np.concatenate(dense_l); ac = np.concatenate(dense_a)
    assert hist.n_records == iters + 1 + 9
    # far fewer rows than records: that is the point
    assert hist.n_rows() < 0.95 * th.shape[0] * C
    assert np.array_equal(hist.dense("theta"), np.transpose(th, (2, 0, 1)))
    assert np.array_equal(hist.dense("like"), lk.T)
    for c in (0, 1, C // 2, C - 1):
        seq = hist.chain(c)
        assert np.array_equal(seq.parameters, th[:, :, c]) and np.array_equal(seq.likelihood, lk[:, c])
        assert np.array_equal(seq.accepted[1:], ac[1:, c].astype(bool))
    # rejected records really do repeat the previous one in the dense history
    rej = ac[1:] == 0
    assert np.array_equal(th[1:][rej.nonzero()[0], :, rej.nonzero()[1]], th[:-1][rej.nonzero()[0], :, rej.nonzero()[1]])
    eng.close()


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_compaction_of_wide_rows_many_segments_and_both_gather_paths(dtype):
    """The gather kernel's corners in one run: model outputs of 1024 columns (sixteen 64-column chunks), 70 chains
    (a ragged last group of 32), 90 records (record segments whose row positions start from a count of the earlier
    accept bytes), records where many chains of a group accept (shared-memory slab) and records where few do (direct
    sector gather; the ragged group has six chains)."""
    from tinyda_b200.engine import STORE_FULL, STORE_NONE
    from tinyda_b200.link import CompactHistory
    C, iters = 70, 89
    eng, w, spec = _cfg2(C, iters, store=[STORE_NONE, STORE_FULL], dtype=dtype, kernel="generic")
    eng.run(iters)
    nrec = iters + 1
    th = eng.fetch(1, "theta", 0, nrec); out = eng.fetch(1, "output", 0, nrec); pr = eng.fetch(1, "prior", 0, nrec)
    ac = eng.fetch(1, "accept", 0, nrec)
    eng.compact_begin(0, nrec, True, ("theta", "stats", "output"), slot=0)
    hist = CompactHistory(C)
    hist.append(eng.compact_collect(0))
    eng.compact_sync()
    flags = ac[1:].astype(bool).reshape(iters, -1)[:, :C]
    per_group = np.stack([flags[:, g:g + 32].sum(axis=1) for g in range(0, C, 32)])
    assert (per_group > 4).any() and ((per_group >= 1) & (per_group <= 4)).any()          # both regimes occur
    assert np.array_equal(hist.dense("theta"), np.transpose(th, (2, 0, 1)))
    assert np.array_equal(hist.dense("output"), np.transpose(out, (2, 0, 1)))
    assert np.array_equal(hist.dense("prior"), pr.T)
    eng.close()


def test_sample_runs_in_blocks_and_returns_the_same_chains():
    """tda.sample cut into blocks (chunk_iterations) returns what one block returns, chain by chain."""
    import tinyda_b200 as tda
    import problems
    defn = problems.CASES["da_pcn_small"]()
    posts, prop, kw = defn["build"](tda)
    a = tda.sample(posts, prop, 23, n_chains=5, seed=11, **kw)
    b = tda.sample(posts, prop, 23, n_chains=5, seed=11, chunk_iterations=4, **kw)
    for key in ("chain_fine_%d", "chain_coarse_%d"):
        for c in range(5):
            sa, sb = a[key % c], b[key % c]
            assert len(sa) == len(sb)
            assert np.array_equal(sa.parameters, sb.parameters) and np.array_equal(sa.likelihood, sb.likelihood)
            assert np.array_equal(sa.model_output, sb.model_output)
    assert len(a["chain_fine_0"]) == 24 and len(a["chain_coarse_0"]) == 23 * kw["subchain_length"]
    # un-seeded calls differ (fresh seed per call); np.random.seed makes them repeatable like the reference
    np.random.seed(4)
    c1 = tda.sample(posts, prop, 6, n_chains=2, initial_parameters=np.zeros(posts[0].model.d), **kw)
    c2 = tda.sample(posts, prop, 6, n_chains=2, initial_parameters=np.zeros(posts[0].model.d), **kw)
    np.random.seed(4)
    c3 = tda.sample(posts, prop, 6, n_chains=2, initial_parameters=np.zeros(posts[0].model.d), **kw)
    assert not np.array_equal(c1["chain_fine_0"].parameters, c2["chain_fine_0"].parameters)
    assert np.array_equal(c1["chain_fine_0"].parameters, c3["chain_fine_0"].parameters)


def test_run_fails_loudly_past_the_history_capacity_and_burn_records_nothing():
    from tinyda_b200._lib import EngineError
    eng, w, spec = _cfg2(256, 5)
    eng.run(5)
    with pytest.raises(EngineError, match="hist_capacity"):
        eng.run(1)
    before = eng.fetch(1, "theta", 0, 6).copy()
    eng.run(7, record=False)                       # burn: same transitions, nothing stored
    assert np.array_equal(eng.fetch(1, "theta", 0, 6), before) and int(eng.n_records()[1]) == 6
    # ... and the chain is where 5 + 7 recorded iterations put it
    ref, _, _ = _cfg2(256, 12)
    ref.run(12)
    assert np.array_equal(eng.get("theta", 1), ref.get("theta", 1))
    eng.history_reset()
    eng.run(5)
    assert np.array_equal(eng.fetch(1, "theta", 0, 1), eng.fetch(1, "theta", 0, 1))
    eng.close(); ref.close()


@pytest.mark.parametrize("case", ["linear_da", "rosenbrock"])
def test_link_qoi_is_the_models_second_return_value(case):
    """posterior.py:95-105 / link.py:38-48: models that return (output, qoi)."""
    import scipy.stats as st
    import tinyda_b200 as tda
    rng = np.random.default_rng(2)
    if case == "linear_da":
        d, m = 6, 20
        G = rng.standard_normal((m, d)) / 3
        Q = rng.standard_normal((3, m))
        q0 = np.array([0.5, -1.0, 2.0])
        y = G @ rng.standard_normal(d) + 0.1 * rng.standard_normal(m)
        prior = st.multivariate_normal(np.zeros(d), np.eye(d))
        pc = tda.Posterior(prior, tda.GaussianLogLike(y[::2], 0.01 * np.eye(m // 2)), tda.LinearModel(G[::2], qoi=Q[:, ::2]))
        pf = tda.Posterior(prior, tda.GaussianLogLike(y, 0.01 * np.eye(m)), tda.LinearModel(G, qoi=(Q, q0)))
        res = tda.sample([pc, pf], tda.CrankNicolson(scaling=0.2), 30, n_chains=4, subchain_length=3, seed=5, chunk_iterations=7)
        fine, coarse = res["chain_fine_1"], res["chain_coarse_1"]
        np.testing.assert_allclose(fine.qoi, fine.model_output @ Q.T + q0, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(coarse.qoi, coarse.model_output @ Q[:, ::2].T, rtol=1e-9, atol=1e-9)
        link = fine[7]
        out, q = pf.model(link.parameters)
        np.testing.assert_allclose(link.qoi, q, rtol=1e-9, atol=1e-9)
        s = tda.get_samples(res, "qoi", level="fine", burnin=2)
        assert s["chain_0"].shape == (29, 3) and s["dimension"] == 3
    else:
        prior = st.multivariate_normal(np.zeros(2), np.eye(2))
        model = tda.Rosenbrock(qoi=(np.array([[2.0]]), [1.0]))
        post = tda.Posterior(prior, tda.GaussianLogLike(np.zeros(1), np.eye(1)), model)
        res = tda.sample(post, tda.MALA(scaling=0.1), 40, n_chains=3, seed=9)
        seq = res["chain_2"]
        np.testing.assert_allclose(seq.qoi, 2.0 * seq.model_output + 1.0, rtol=1e-12)
        assert res["chain_0"][3].qoi.shape == (1,)
    # a model without one keeps Link.qoi = None like the reference
    res0 = tda.sample(tda.Posterior(prior, tda.GaussianLogLike(np.zeros(1), np.eye(1)), tda.LinearModel(np.ones((1, prior.mean.size)))),
                      tda.GaussianRandomWalk(np.eye(prior.mean.size)), 5, seed=1)
    assert res0["chain_0"][2].qoi is None


@pytest.mark.parametrize("dtype,C,iters", [("float32", 256, 300), ("float64", 97, 121)])
def test_device_ess_and_rhat_match_the_host_estimators(dtype, C, iters):
    """tda_ess_sums (device: sort, average ranks, normal scores, all-lag autocovariance sums) against the
    NumPy estimators (diagnostics.ess_bulk / rhat) on the fetched history of the same run."""
    import tinyda_b200 as tda
    from tinyda_b200.diagnostics import ess_rhat_from_sums
    eng, w, spec = _cfg2(C, iters, dtype=dtype, kernel=None if dtype == "float32" else "generic")
    eng.run(iters)
    sums, folded = eng.ess_sums(1, 1, iters)                    # skip the initial link
    ess, rh = ess_rhat_from_sums(sums, folded)
    th = np.transpose(eng.fetch(1, "theta", 1, iters), (2, 0, 1)).astype(np.float32 if dtype == "float32" else np.float64)
    for k in (0, 5, 63):
        xk = th[:, :, k].astype(np.float32).astype(np.float64)  # the device ranks float32 keys
        assert abs(ess[k] / tda.ess_bulk(xk) - 1) < 5e-3, (k, ess[k], tda.ess_bulk(xk))
        assert abs(rh[k] / tda.rhat(xk) - 1) < 2e-3, (k, rh[k], tda.rhat(xk))
    # a lag window that cuts positive autocorrelation mass off over-estimates the ESS (the default is all lags)
    s2, f2 = eng.ess_sums(1, 1, iters, n_lag=48)
    ess2, _ = ess_rhat_from_sums(s2, f2)
    assert np.all(ess2 > 0.95 * ess)
    eng.close()
