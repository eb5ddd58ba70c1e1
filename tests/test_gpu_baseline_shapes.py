"""BASELINE.json's configs 3, 4 and 5 at their REAL shapes against the CPU oracle (itself pinned to the
reference by the golden fixtures): the engine's own Philox streams are exported with tda_fill_streams and
fed to the oracle, chain by chain.

  cfg3  MALA on the 2-D Rosenbrock, all 2^20 chains on the device, a random subset of 64 replayed
  cfg4  4-level MLDA + state-independent AEM, Poisson grids 64/128/256/512, d = 16, 31 sensors, J = [10,5,5]
  cfg5  DREAM with the shared archive, d = 32, 256 observations, M0 = 16
"""
import numpy as np
import pytest

import problems

pytestmark = pytest.mark.gpu


def test_cfg3_full_size_random_subset_of_chains_replayed_by_the_oracle():
    """2^20 chains (float32, register kernel, adaptive MALA); 64 of them, picked at random, are replayed by
    the float64 oracle from the exported streams of exactly those chains (chain id -> Philox key)."""
    from oracle import tinyda_oracle as orc
    from tinyda_b200 import lower_problem, workloads
    from tinyda_b200.engine import Engine, STORE_STATS
    w = workloads.cfg3_mala()
    spec = lower_problem(w["posteriors"], w["proposal"])
    C, iters, seed = 1 << 20, 120, 31                     # crosses the first adaptation boundary (period 100)
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(3))
    eng = Engine(spec, C, dtype="float32", seed=seed, store=STORE_STATS, capacity_iterations=iters)
    assert eng.kernel() == "reg"
    eng.init(theta0)
    eng.run(iters)
    acc = eng.fetch(0, "accept")                          # [iters + 1, C]
    th = eng.fetch(0, "theta")                            # [iters + 1, 2, C]
    scal = eng.get("scaling")
    eng.close()
    subset = np.random.default_rng(99).choice(C, size=64, replace=False)
    nz, nu = problems.stream_sizes(spec, iters)
    identical, steps_ok, steps = 0, 0, 0
    for c in subset:
        one = Engine(spec, 1, dtype="float32", seed=seed, store=STORE_STATS, capacity_iterations=1, chain_offset=int(c))
        z, u = one.fill_streams(nz, nu)
        one.close()
        out, chains = orc.run_chains(spec, theta0[c:c + 1], z, u, iters)
        ref_acc, ref_th = out[0]["acc"][0], out[0]["theta"][0]
        mine_acc, mine_th = acc[:, c].astype(bool), th[:, :, c].astype(np.float64)
        diff = np.nonzero(mine_acc[1:] != ref_acc[1:])[0]
        k = int(diff[0]) + 1 if diff.size else iters + 1    # records 0..k-1 agree
        steps += iters
        steps_ok += k - 1
        scale = np.abs(ref_th).max()
        assert np.abs(mine_th[:k] - ref_th[:k]).max() <= 1e-5 * scale + 1e-6, (c, k)
        if k == iters + 1:
            identical += 1
            # the adapted step size after the first period follows the same rule (proposal.py:228-245)
            np.testing.assert_allclose(scal[c], chains[0].scaling, rtol=1e-4)
    print("\ncfg3: %d / 64 replayed chains identical over %d steps; %d / %d decisions before a first near-tie flip"
          % (identical, iters, steps_ok, steps))
    assert identical >= 56 and steps_ok >= 0.95 * steps


def test_cfg4_real_shape_philox_streams_fed_to_the_oracle():
    """BASELINE cfg4: d = 16, grids 64/128/256/512, 31 sensors, J = [10, 5, 5], state-independent AEM.
    float64 engine, 3 chains x 3 finest iterations = 750 level-0 steps per chain."""
    from oracle import tinyda_oracle as orc
    from tinyda_b200 import lower_problem, workloads
    from tinyda_b200.engine import Engine, STORE_FULL
    w = workloads.cfg4_mlda()
    kw = w["kwargs"]
    spec = lower_problem(w["posteriors"], w["proposal"], kw["subchain_length"], kw["adaptive_error_model"])
    assert spec["d"] == 16 and [lv["model"]["n_grid"] for lv in spec["levels"]] == [64, 128, 256, 512]
    assert spec["J"] == [10, 5, 5] and spec["levels"][0]["model"]["m"] == 31
    C, iters = 3, 3
    theta0 = 0.3 * w["prior"].rvs(C, random_state=np.random.default_rng(4))
    eng = Engine(spec, C, dtype="float64", seed=55, store=STORE_FULL, capacity_iterations=iters, chain_offset=11)
    eng.init(theta0)
    eng.run(iters)
    nz, nu = problems.stream_sizes(spec, iters)
    z, u = eng.fill_streams(nz, nu)
    out, chains = orc.run_chains(spec, theta0, z, u, iters)
    for l in range(4):
        acc = eng.fetch(l, "accept").T.astype(bool)
        th = np.transpose(eng.fetch(l, "theta"), (2, 0, 1))
        lk = eng.fetch(l, "like").T
        F = np.transpose(eng.fetch(l, "output"), (2, 0, 1))
        assert acc.shape == out[l]["acc"].shape
        assert np.array_equal(acc, out[l]["acc"]), "level %d decisions differ" % l
        np.testing.assert_allclose(th, out[l]["theta"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(F, out[l]["F"], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(lk, out[l]["like"], rtol=1e-6, atol=1e-6)
    assert out[0]["acc"].shape[1] == iters * 250 and out[0]["acc"].mean() > 0.02
    assert np.array_equal(eng.get("cursors").T, np.array([[ch.S.nz, ch.S.nu] for ch in chains]))
    eng.close()


def test_cfg5_real_shape_shared_archive_philox_streams_fed_to_the_oracle():
    """BASELINE cfg5: d = 32, 256 observations, DREAM(M0 = 16, delta = 1, nCR = 3), shared archive with
    lock-step visibility; 8 chains x 40 steps in float64 against the oracle's DreamEnsemble."""
    from oracle import tinyda_oracle as orc
    from tinyda_b200 import lower_problem, workloads
    from tinyda_b200.engine import Engine, STORE_FULL
    w = workloads.cfg5_dream()
    spec = lower_problem(w["posteriors"], w["proposal"])
    assert spec["d"] == 32 and spec["levels"][0]["model"]["m"] == 256 and int(spec["proposal"]["M0"]) == 16
    C, iters, M0, d = 8, 40, 16, 32
    rng = np.random.default_rng(6)
    theta0 = w["prior"].rvs(C, random_state=rng)
    archive0 = w["prior"].rvs(C * M0, random_state=rng).reshape(C, M0, d)
    eng = Engine(spec, C, dtype="float64", seed=21, store=STORE_FULL, capacity_iterations=iters, archive0=archive0)
    eng.init(theta0)
    eng.run(iters)
    nz, nu = problems.stream_sizes(spec, iters)
    z, u = eng.fill_streams(nz, nu)
    out, chains = orc.run_chains(spec, theta0, z, u, iters, archive0)
    acc = eng.fetch(0, "accept").T.astype(bool)
    th = np.transpose(eng.fetch(0, "theta"), (2, 0, 1))
    assert np.array_equal(acc, out[0]["acc"])
    np.testing.assert_allclose(th, out[0]["theta"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(eng.fetch(0, "like").T, out[0]["like"], rtol=1e-9, atol=1e-8)
    assert np.array_equal(eng.get("cursors").T, np.array([[ch.S.nz, ch.S.nu] for ch in chains]))
    assert 0.02 < acc[:, 1:].mean() < 0.98
    eng.close()
