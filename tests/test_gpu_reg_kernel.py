"""GPU parity of the register-resident single-level kernel (csrc/tda_mh_reg.cuh): one thread per
chain, state in registers.  Checked against the reference's own trajectories (golden fixtures,
injected streams), against the CPU oracle fed the kernel's Philox streams, and against the
generic lock-step kernel it is interchangeable with."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu

FP32_STATE_RTOL = 1e-5      # north_star's float32 tolerance, relative to the largest |theta| of the chain

CASES = ["mh_rwmh_linreg", "mh_pcn_diag", "mala_rosenbrock", "mala_linear"]


def _store_F(name):
    return True


@pytest.mark.parametrize("name", CASES)
def test_reg_kernel_fp64_matches_reference(name):
    from gpu_util import run_engine
    g = golden_io.load(name)
    out, eng = run_engine(g, "float64", store_F=_store_F(name), kernel="reg")
    assert eng.kernel() == "reg"
    ref = g["ref"][0]
    assert np.array_equal(out[0]["acc"], ref["acc"])
    scale = np.abs(ref["theta"]).max()
    np.testing.assert_allclose(out[0]["theta"], ref["theta"], rtol=1e-10, atol=1e-10 * scale)
    np.testing.assert_allclose(out[0]["prior"], ref["prior"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(out[0]["like"], ref["like"], rtol=1e-9, atol=1e-8)
    if _store_F(name):
        np.testing.assert_allclose(out[0]["F"], ref["F"], rtol=1e-9, atol=1e-9 * np.abs(ref["F"]).max())
    assert np.array_equal(eng.get("cursors").T, g["consumed"])


@pytest.mark.parametrize("name", CASES)
def test_reg_kernel_fp32_matches_reference_until_near_tie(name):
    from gpu_util import run_engine, first_divergence
    g = golden_io.load(name)
    out, eng = run_engine(g, "float32", store_F=False, kernel="reg")
    ref = g["ref"][0]
    fd = first_divergence(out[0]["acc"], ref["acc"])
    assert fd.sum() >= 0.5 * ref["acc"].size
    worst = 0.0
    for c in range(fd.size):
        k = int(fd[c])
        if k:
            scale = np.abs(ref["theta"][c]).max() + 1e-30
            worst = max(worst, np.abs(out[0]["theta"][c, :k] - ref["theta"][c, :k]).max() / scale)
    print("\nreg kernel fp32 [%s]: %d of %d decisions before a first flip, max relative state error %.2e"
          % (name, int(fd.sum()), ref["acc"].size, worst))
    assert worst <= FP32_STATE_RTOL, worst


def _engine(g, dtype, kernel, iters, store, seed=77, C=None, offset=5):
    from tinyda_b200.engine import Engine
    spec, theta0 = g["spec"], g["theta0"]
    if C is not None:
        theta0 = np.resize(theta0, (C, theta0.shape[1]))
    eng = Engine(spec, theta0.shape[0], dtype=dtype, rng="philox", seed=seed, store=store,
                 capacity_iterations=iters, chain_offset=offset)
    eng.select_kernel(kernel)
    eng.init(theta0)
    return eng, theta0


@pytest.mark.parametrize("name", CASES)
def test_reg_kernel_philox_streams_fed_to_the_oracle(name):
    """Production RNG mode: the kernel's cached-block Philox draws are the engine's documented
    streams (tda_fill_streams) -- the oracle fed those streams reproduces the trajectories."""
    from tinyda_b200.engine import STORE_STATS
    from oracle import tinyda_oracle as orc
    g = golden_io.load(name)
    iters = g["iterations"]
    eng, theta0 = _engine(g, "float64", "reg", iters, STORE_STATS)
    eng.run(iters)
    z, u = eng.fill_streams(g["z"].shape[1], g["u"].shape[1])
    out, chains = orc.run_chains(g["spec"], theta0, z, u, iters)
    assert np.array_equal(eng.fetch(0, "accept").T.astype(bool), out[0]["acc"])
    np.testing.assert_allclose(np.transpose(eng.fetch(0, "theta"), (2, 0, 1)), out[0]["theta"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(eng.get("scaling"), [ch.scaling for ch in chains], rtol=1e-12)


@pytest.mark.parametrize("name", ["mh_rwmh_linreg", "mala_rosenbrock"])
def test_reg_and_generic_kernels_are_interchangeable_and_resumable(name):
    """fp64, 300 chains (ragged last tile): generic for a steps then reg for b steps equals reg for
    a+b steps to rounding; reg run(a); run(b) equals reg run(a+b) bit for bit; in auto mode the
    register kernel is the one selected."""
    from tinyda_b200.engine import STORE_STATS
    g = golden_io.load(name)
    a, b = 37, 45

    def run(plan):
        eng, _ = _engine(g, "float64", "auto", a + b, STORE_STATS, C=300)
        assert eng.kernel() == "reg"
        for kern, n in plan:
            eng.select_kernel(kern)
            eng.run(n)
        th, lk, sc = eng.fetch(0, "theta"), eng.fetch(0, "like"), eng.get("scaling")
        mom = eng.get("moments")
        eng.close()
        return th, lk, sc, mom

    th0, lk0, sc0, mom0 = run([("reg", a + b)])
    th1, lk1, sc1, mom1 = run([("reg", a), ("reg", b)])
    assert np.array_equal(th0, th1) and np.array_equal(lk0, lk1) and np.array_equal(sc0, sc1)
    assert np.array_equal(mom0, mom1)
    th2, lk2, sc2, mom2 = run([("generic", a), ("reg", b)])
    same = (np.abs(th2 - th0) <= 1e-9 * (1 + np.abs(th0))).all(axis=(0, 1))
    assert same.mean() > 0.98          # a near-tie may send an occasional chain elsewhere
    np.testing.assert_allclose(sc2[same], sc0[same], rtol=1e-9)
    np.testing.assert_allclose(mom2[:, :, same], mom0[:, :, same], rtol=1e-9, atol=1e-9)


def test_reg_kernel_mala_rosenbrock_full_size_properties():
    """cfg3 at BASELINE.json's size (2^20 chains, float32, Philox): the register kernel and the
    generic kernel run the same chains on the same streams -> accept counts agree for almost
    every chain and the pooled moments agree; adaptive step sizes end up near the MALA target
    acceptance rate."""
    from tinyda_b200 import lower_problem, workloads
    from tinyda_b200.engine import Engine, STORE_NONE
    w = workloads.cfg3_mala()
    spec = lower_problem(w["posteriors"], w["proposal"])
    C, iters = 1 << 20, 600
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(3))
    res = {}
    for kern in ("reg", "generic"):
        eng = Engine(spec, C, dtype="float32", seed=11, store=STORE_NONE)
        eng.select_kernel(kern)
        eng.init(theta0)
        eng.run(iters)
        res[kern] = (eng.get("accept_counts")[0], eng.get("moments"), eng.get("scaling"))
        eng.close()
    acc_r, mom_r, sc_r = res["reg"]
    acc_g, mom_g, sc_g = res["generic"]
    assert (acc_r == acc_g).mean() > 0.9
    mean_r, mean_g = mom_r[0].sum(axis=1) / (C * iters), mom_g[0].sum(axis=1) / (C * iters)
    np.testing.assert_allclose(mean_r, mean_g, atol=2e-3)
    # the last adaptation window's acceptance is pulled towards alpha* = 0.57 (proposal.py:896)
    assert np.isfinite(sc_r).all() and (sc_r > 0).all()
    assert abs(np.median(sc_r) / np.median(sc_g) - 1) < 0.05
