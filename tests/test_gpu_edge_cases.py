"""Edge cases of the hot path on the GPU: empty runs, a single chain, chain counts that do not fill
a tile, the smallest and largest shapes the tensor-core kernel accepts, and what it must refuse."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu


def test_zero_iterations_returns_only_the_initial_links():
    import tinyda_b200 as tda
    import problems
    defn = problems.CASES["da_pcn_small"]()
    posts, prop, kw = defn["build"](tda)
    res = tda.sample(posts, prop, 0, n_chains=3, seed=1, **kw)
    assert res["iterations"] == 1
    assert len(res["chain_fine_0"]) == 1 and len(res["chain_coarse_2"]) == 0
    link = res["chain_fine_1"][0]
    assert np.isfinite(link.posterior) and link.parameters.shape == (8,)


@pytest.mark.parametrize("name", ["mh_pcn_diag", "da_aem_linear", "mlda3_linear"])
@pytest.mark.parametrize("C", [1, 129, 257])
def test_ragged_chain_counts_do_not_change_any_chain(name, C):
    """Chains are independent: chain i of a C-chain engine equals chain i of any other engine fed
    the same stream, whatever the padding of the last tile (here: the golden chain 0 replicated)."""
    from tinyda_b200.engine import Engine, STORE_FULL
    g = golden_io.load(name)
    spec, iters = g["spec"], g["iterations"]
    theta0 = np.repeat(g["theta0"][:1], C, axis=0)
    z = np.repeat(g["z"][:1], C, axis=0)
    u = np.repeat(g["u"][:1], C, axis=0)
    eng = Engine(spec, C, dtype="float64", rng="injected", streams=(z, u), store=STORE_FULL,
                 capacity_iterations=iters)
    eng.init(theta0)
    eng.run(iters)
    for l in range(spec["n_levels"]):
        ref = g["ref"][l]
        th = np.transpose(eng.fetch(l, "theta"), (2, 0, 1))
        acc = eng.fetch(l, "accept").T.astype(bool)
        for c in (0, C - 1):
            assert np.array_equal(acc[c], ref["acc"][0])
            np.testing.assert_allclose(th[c], ref["theta"][0], rtol=1e-10, atol=1e-12)
    eng.close()


@pytest.mark.parametrize("m_c,m_f", [(16, 64), (128, 1920)])
def test_tc16_extreme_shapes_agree_with_the_generic_kernel(m_c, m_f):
    """Smallest and largest observation counts the fp16-split tensor-core kernel accepts, against
    the generic fp32 kernel on the same (z16) Philox streams."""
    from tinyda_b200 import lower_problem
    from tinyda_b200.engine import Engine, STORE_NONE, STORE_STATS
    from tinyda_b200.workloads import cfg2_da
    w = cfg2_da(m_f=m_f, m_c=m_c, beta=0.05 if m_f > 100 else 0.2)
    spec = lower_problem(w["posteriors"], w["proposal"], 10)
    C, iters = 256, 12
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(3))
    out = {}
    for kern in ("tc16", "generic"):
        eng = Engine(spec, C, dtype="float32", seed=5, store=[STORE_NONE, STORE_STATS], capacity_iterations=iters)
        eng.select_kernel(kern)
        if kern == "generic":
            eng.set_z_round(True)
        eng.init(theta0)
        eng.run(iters)
        out[kern] = (eng.fetch(1, "accept"), eng.fetch(1, "theta"), eng.fetch(1, "like"))
        eng.close()
    acc_a, th_a, lk_a = out["tc16"]
    acc_b, th_b, lk_b = out["generic"]
    same = acc_a == acc_b
    assert same.mean() > 0.95, same.mean()
    ok = same.all(axis=0)
    assert ok.mean() > 0.5
    scale = np.abs(th_b).max()
    assert np.abs(th_a[:, :, ok] - th_b[:, :, ok]).max() < 2e-3 * scale
    assert np.abs(lk_a[:, ok] - lk_b[:, ok]).max() < 2e-3 * np.abs(lk_b).max() + 0.05


def test_tensor_core_kernel_refuses_what_it_cannot_run():
    """Explicit kernel selection fails loudly (no silent fallback) on unsupported configurations;
    automatic selection falls back to the generic kernel."""
    from tinyda_b200._lib import EngineError
    from tinyda_b200.engine import Engine, STORE_STATS
    g = golden_io.load("da_pcn_small")            # d = 8: not a tensor-core shape
    eng = Engine(g["spec"], 4, dtype="float32", seed=1, store=STORE_STATS, capacity_iterations=4)
    eng.init(g["theta0"])
    assert eng.kernel() == "generic"
    for kern in ("tc16", "tc", "reg"):
        eng.select_kernel(kern)
        with pytest.raises(EngineError, match="does not support"):
            eng.run(1)
    eng.select_kernel("auto")
    eng.run(2)
    eng.close()
    # float64 engines never take the float32 tensor-core path
    g2 = golden_io.load("da_pcn_cfg2")
    e2 = Engine(g2["spec"], 256, dtype="float64", seed=1, store=STORE_STATS, capacity_iterations=2)
    assert e2.kernel() == "generic"
    e2.close()


@pytest.mark.parametrize("name,dtype,kernel", [("da_pcn_cfg2", "float32", "tc16"), ("mlda3_aem_linear", "float64", "generic"),
                                               ("dreamz_adaptive", "float64", "generic"), ("mala_rosenbrock", "float32", "reg"),
                                               ("am_linear", "float64", "generic")])
def test_checkpoint_resume_is_exact(name, dtype, kernel):
    """tda_state_save / tda_state_load: an engine restored from a checkpoint continues the chains
    bit for bit (constants, Links of every level, proposal and error-model state, stream cursors
    are all in the blob; the history is not)."""
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    g = golden_io.load(name)
    spec = g["spec"]
    store = [STORE_NONE] * (spec["n_levels"] - 1) + [STORE_STATS]
    C = 256
    theta0 = np.resize(g["theta0"], (C, g["theta0"].shape[1]))
    arch = None if g["archive0"] is None else np.resize(g["archive0"], (C,) + g["archive0"].shape[1:])
    a, b = 7, 9

    def make():
        e = Engine(spec, C, dtype=dtype, seed=21, store=store, capacity_iterations=a + b, archive0=arch)
        e.select_kernel(kernel)
        return e

    e1 = make()
    e1.init(theta0)
    e1.run(a)
    blob = e1.save_state()
    e1.history_reset()
    e1.run(b)
    top = spec["n_levels"] - 1
    want = (e1.fetch(top, "theta", 0, b), e1.fetch(top, "like", 0, b), e1.get("scaling"), e1.get("cursors"))
    e1.close()
    e2 = make()                       # fresh engine: never initialised, state comes from the blob
    e2.load_state(blob, iterations_done=a)
    assert e2.kernel() == kernel
    e2.run(b)
    got = (e2.fetch(top, "theta", 0, b), e2.fetch(top, "like", 0, b), e2.get("scaling"), e2.get("cursors"))
    e2.close()
    for w, x in zip(want, got):
        assert np.array_equal(w, x)
    # a blob from a different configuration is refused
    other = golden_io.load("mh_pcn_diag")
    e3 = Engine(other["spec"], C, dtype="float64", seed=1, store=STORE_STATS, capacity_iterations=2)
    from tinyda_b200._lib import EngineError
    with pytest.raises(EngineError, match="different configuration|too small"):
        e3.load_state(blob)
    e3.close()


def test_the_binding_stub_of_integration_md_runs_as_written():
    """INTEGRATION.md section B shows the ctypes stub a tinyDA maintainer would add.  This executes
    that very code block (only the library path is made absolute) on a small DA problem and
    compares its output with the package's own engine wrapper."""
    import os
    import re
    import ctypes as C
    from tinyda_b200 import _lib
    from tinyda_b200.engine import Engine, STORE_FULL
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\n# tinyDA/_b200.py.*?\n(.*?)```", text, re.S).group(1)
    block = block.replace('C.CDLL("libtinyda_b200.so")', 'C.CDLL(%r)' % _lib.LIB_PATH)
    ns = {}
    exec(block, ns)

    g = golden_io.load("da_pcn_small")
    spec, theta0, iters = g["spec"], g["theta0"], 12
    # the package's engine fills the POD config; constants are collected from its upload calls
    ups = []
    orig = Engine._up
    Engine._up = lambda self, what, level, arr: (ups.append((what, level, np.array(arr, dtype=np.float64))), orig(self, what, level, arr))[1]
    try:
        eng = Engine(spec, theta0.shape[0], dtype="float64", rng="philox", seed=3, store=STORE_FULL, capacity_iterations=iters)
    finally:
        Engine._up = orig
    eng.init(theta0)
    eng.run(iters)
    want = eng.fetch(1, "theta")
    theta = ns["sample_da_b200"](eng.cfg, ups, theta0, iters)
    eng.close()
    assert theta.shape == want.shape and np.array_equal(theta, want)


@pytest.mark.parametrize("script,args", [("basic_sampler.py", ["8", "3000"]), ("delayed_acceptance.py", ["512", "40"])])
def test_example_scripts_run(script, args):
    """The examples are the reference's notebooks with the import swapped; they must run as written."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", script)] + args, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]


def test_create_link_and_get_map_run_on_the_device():
    """Posterior.create_link (posterior.py:78-110) and get_MAP / get_ML (utils.py:204-269), used by
    three of the reference's example notebooks for `initial_parameters=MAP`: Links come from the
    CUDA engine; on a linear-Gaussian problem the MAP is the closed-form posterior mean and the ML
    estimate the least-squares solution."""
    import scipy.stats as stats
    import tinyda_b200 as tda
    from tinyda_b200.workloads import conjugate_posterior
    from oracle import tinyda_oracle as orc
    rng = np.random.default_rng(3)
    d, m, sig2 = 5, 12, 0.05
    prior = stats.multivariate_normal(0.1 * np.ones(d), np.eye(d) + 0.2)
    G = rng.standard_normal((m, d))
    y = G @ prior.rvs(random_state=rng) + np.sqrt(sig2) * rng.standard_normal(m)
    post = tda.Posterior(prior, tda.GaussianLogLike(y, sig2 * np.eye(m)), tda.LinearModel(G, offset=0.3 * np.ones(m)))
    x = prior.rvs(random_state=rng)
    link = post.create_link(x)
    np.testing.assert_allclose(link.prior, prior.logpdf(x), rtol=1e-12)
    np.testing.assert_allclose(link.model_output, G @ x + 0.3, rtol=1e-12)
    np.testing.assert_allclose(link.likelihood, -0.5 * np.sum((G @ x + 0.3 - y) ** 2) / sig2, rtol=1e-12)
    assert link.posterior == link.prior + link.likelihood
    mu, S = conjugate_posterior(G, y - 0.3, sig2, prior)
    MAP = tda.get_MAP(post, initial_parameters=np.zeros(d))
    np.testing.assert_allclose(MAP, mu, atol=2e-4)
    ML = tda.get_ML(post, initial_parameters=np.zeros(d), method="L-BFGS-B")
    np.testing.assert_allclose(ML, np.linalg.lstsq(G, y - 0.3, rcond=None)[0], atol=2e-4)
    MAP2 = tda.get_MAP(post, method="differential_evolution", bounds=[(-4, 4)] * d, seed=1, maxiter=300, tol=1e-10)
    np.testing.assert_allclose(MAP2, mu, atol=5e-3)


def test_states_outside_the_fp16_operand_range_run_on_another_kernel():
    """The fp16-split kernel's theta image is fixed-point at a scale taken from the prior (|mean| + 12 sd).
    Initial parameters far outside it would overflow to inf and freeze the chain: such a job is moved to a
    kernel that can hold them, and still moves."""
    from tinyda_b200 import lower_problem
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    from tinyda_b200.workloads import cfg2_da
    w = cfg2_da()
    spec = lower_problem(w["posteriors"], w["proposal"], 10)
    C = 256
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
    theta0[3] *= 400.0                                   # ~400 prior standard deviations out
    eng = Engine(spec, C, dtype="float32", seed=1, store=[STORE_NONE, STORE_STATS], capacity_iterations=20)
    eng.init(theta0)
    assert eng.kernel() == "tc16"                        # by configuration ...
    eng.run(20)
    assert eng.kernel() != "tc16"                        # ... but not with these states
    th = eng.fetch(1, "theta")
    assert np.isfinite(th).all() and np.isfinite(eng.fetch(1, "like")).all()
    acc = eng.get("accept_counts")
    assert acc[0][3] > 0                                 # the far-out chain accepts coarse proposals (it is not frozen)
    eng.init(w["prior"].rvs(C, random_state=np.random.default_rng(2)))
    assert eng.kernel() == "tc16"                        # fresh states: the fast kernel again
    eng.close()
