"""CPU-side tests: the host mirror of the reference's API (validation, factories, result
containers), the C-ABI library's symbol table, and the loud failure without a GPU."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest
import scipy.stats as stats

import tinyda_b200 as tda
from tinyda_b200.link import LinkSequence
from tinyda_b200.lowering import spec_to_flat, spec_from_flat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _post(d=3, m=5, prior=None):
    rng = np.random.default_rng(0)
    G = rng.standard_normal((m, d))
    prior = prior or stats.multivariate_normal(np.zeros(d), np.eye(d))
    return tda.Posterior(prior, tda.GaussianLogLike(rng.standard_normal(m), 0.1 * np.eye(m)), tda.LinearModel(G))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tinyda_b200.h")).read()
    declared = set(re.findall(r"\b(tda_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"tda_engine", "tda_config", "tda_level_config"}
    assert len(declared) >= 15
    lib = ctypes.CDLL(os.path.join(ROOT, "tinyda_b200", "libtinyda_b200.so"))
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name
    from tinyda_b200 import _lib
    assert set(_lib.EXPORTS) == declared
    assert lib.tda_abi_version() == 3


def test_config_struct_layout_matches_header(tmp_path):
    """ctypes mirror of tda_config vs the C compiler's layout of include/tinyda_b200.h."""
    import subprocess
    from tinyda_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "tinyda_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(tda_level_config), sizeof(tda_config),'
        ' offsetof(tda_config, seed), offsetof(tda_config, scaling), offsetof(tda_config, prior_logconst),'
        ' offsetof(tda_config, level)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    C = _lib.Config
    assert got == [ctypes.sizeof(_lib.LevelConfig), ctypes.sizeof(C), C.seed.offset, C.scaling.offset,
                   C.prior_logconst.offset, C.level.offset]


def test_gaussian_loglike_factory_rule():
    y = np.arange(4.0)
    assert isinstance(tda.GaussianLogLike(y, 0.5 * np.eye(4)), tda.IsotropicGaussianLogLike)
    assert isinstance(tda.GaussianLogLike(y, np.diag([1.0, 2, 3, 4])), tda.DiagonalGaussianLogLike)
    C = np.eye(4)
    C[0, 1] = C[1, 0] = 0.1
    lk = tda.GaussianLogLike(y, C)
    assert type(lk) is tda.DefaultGaussianLogLike
    with pytest.raises(TypeError):
        tda.GaussianLogLike(y, [[1.0]])
    with pytest.raises(ValueError):
        tda.GaussianLogLike(y, np.eye(3))
    with pytest.raises(TypeError):
        tda.AdaptiveGaussianLogLike(y, np.ones(4))
    assert tda.AdaptiveLogLike is tda.AdaptiveGaussianLogLike


def test_proposal_validation_matches_reference():
    with pytest.raises(TypeError):
        tda.GaussianRandomWalk([[1.0]])
    with pytest.raises(ValueError):
        tda.GaussianRandomWalk(np.ones((2, 3)))
    with pytest.raises(TypeError):
        tda.AdaptiveMetropolis(C0=1.0)
    am = tda.AdaptiveMetropolis(C0=np.eye(10))
    assert am.sd == min(1, 2.4 ** 2 / 10) and am.scaling == 1
    assert tda.GaussianRandomWalk.alpha_star == 0.24 and tda.MALA.alpha_star == 0.57
    assert tda.GaussianRandomWalk.is_symmetric and not tda.CrankNicolson.is_symmetric


def test_sample_rejects_what_the_reference_rejects():
    post = _post(prior=stats.norm(0, 1))
    with pytest.raises(TypeError, match="scipy.stats.multivariate_normal"):
        tda.sample(post, tda.CrankNicolson(), 10)
    post = _post()
    with pytest.raises(TypeError, match="list, numpy array or None"):
        tda.sample(post, tda.GaussianRandomWalk(np.eye(3)), 10, initial_parameters=(0, 0, 0))
    with pytest.raises(AssertionError):
        tda.sample(post, tda.GaussianRandomWalk(np.eye(3)), 10, n_chains=2, initial_parameters=[np.zeros(3)])
    with pytest.raises(AssertionError):
        tda.sample(post, tda.GaussianRandomWalk(np.eye(3)), 10, initial_parameters=np.zeros(4))


def test_error_model_and_randomize_options_are_checked_like_the_reference():
    """chain.py:305 (unknown error model), :310-314 (randomize needs J > 1 and the coarse chain),
    sampler.py:184-193 (warnings); plus what the device lowering cannot honour."""
    post = _post()
    rw = tda.GaussianRandomWalk(np.eye(3))
    with pytest.raises(ValueError, match="state-dependent, state-independent or None"):
        tda.sample([post, post], rw, 5, adaptive_error_model="sometimes")
    with pytest.raises(ValueError, match="subchain_length > 1"):
        tda.sample([post, post], rw, 5, subchain_length=1, randomize_subchain_length=True)
    with pytest.raises(ValueError, match="requires storing the coarse chain"):
        tda.sample([post, post], rw, 5, subchain_length=3, randomize_subchain_length=True,
                   store_coarse_chain=False)
    with pytest.raises(TypeError, match="AdaptiveGaussianLogLike"):
        with pytest.warns(UserWarning, match="not guaranteed to be ergodic"):
            tda.sample([post, post], rw, 5, subchain_length=3, adaptive_error_model="state-dependent")
    # lowering: aem code 2, randomize flag; the state-dependent model is two-level only
    prior = stats.multivariate_normal(np.zeros(3), np.eye(3))
    G = np.arange(12.0).reshape(4, 3) / 10
    pc = tda.Posterior(prior, tda.AdaptiveGaussianLogLike(np.zeros(4), 0.1 * np.eye(4)), tda.LinearModel(G))
    pf = tda.Posterior(prior, tda.GaussianLogLike(np.zeros(4), 0.1 * np.eye(4)), tda.LinearModel(G))
    spec = tda.lower_problem([pc, pf], tda.CrankNicolson(0.2), 3, "state-dependent", True)
    assert spec["aem"] == 2 and spec["randomize"] == 1
    assert tda.lower_problem([pc, pf], rw, 3, "state-independent")["aem"] == 1
    with pytest.raises(ValueError, match="two-level"):
        tda.lower_problem([pc, pc, pf], rw, [2, 2], "state-dependent")
    with pytest.raises(ValueError, match="two-level"):
        tda.lower_problem([pc, pc, pf], rw, [2, 2], None, True)


def test_python_callable_model_is_rejected_not_run_on_cpu():
    prior = stats.multivariate_normal(np.zeros(2), np.eye(2))
    post = tda.Posterior(prior, tda.GaussianLogLike(np.zeros(3), np.eye(3)), lambda th: np.zeros(3))
    with pytest.raises(TypeError, match="no CPU fallback"):
        tda.sample(post, tda.GaussianRandomWalk(np.eye(2)), 10)


def test_sample_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tinyda_b200._lib import EngineError
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(EngineError, match="no CPU path"):
            tda.sample(_post(), tda.GaussianRandomWalk(np.eye(3)), 5, n_chains=2, subsampling_rate=3)


def test_subsampling_rate_warns():
    import torch
    post = _post()
    with pytest.warns(UserWarning, match="subsampling_rate has been deprecated"):
        try:
            tda.sample([post, post], tda.GaussianRandomWalk(np.eye(3)), 2, subsampling_rate=2)
        except Exception:
            if torch.cuda.is_available():
                raise


def test_lowering_roundtrip_and_checks():
    p = _post()
    spec = tda.lower_problem([p, p, p], tda.GaussianRandomWalk(0.1 * np.eye(3)), [3, 2])
    assert spec["n_levels"] == 3 and spec["J"] == [3, 2] and spec["d"] == 3
    back = spec_from_flat(spec_to_flat(spec))
    assert back["J"] == [3, 2]
    np.testing.assert_array_equal(back["levels"][1]["model"]["A"], spec["levels"][1]["model"]["A"])
    with pytest.raises(TypeError):      # AEM needs adaptive likelihoods on the coarse levels
        tda.lower_problem([p, p], tda.GaussianRandomWalk(np.eye(3)), 2, "state-independent")
    with pytest.raises(ValueError):
        tda.lower_problem([p, p], tda.GaussianRandomWalk(np.eye(3)), [2, 2])
    # scipy's own whitening reproduces scipy's logpdf
    prior = stats.multivariate_normal(np.array([0.3, -0.2, 0.1]), np.array([[2, .3, 0], [.3, 1, .1], [0, .1, .5]]))
    lp = tda.posterior.lower_prior(prior)
    x = np.array([0.5, 0.1, -1.0])
    val = -0.5 * (lp["logconst"] + np.sum(((x - lp["mean"]) @ lp["LP"]) ** 2))
    assert abs(val - prior.logpdf(x)) < 1e-12


def test_link_sequence_behaves_like_a_list_of_links():
    n, d, m = 7, 3, 4
    rng = np.random.default_rng(1)
    seq = LinkSequence(rng.standard_normal((n, d)), rng.standard_normal(n), rng.standard_normal(n),
                       rng.standard_normal((n, m)), np.ones(n, bool))
    assert len(seq) == n
    l = seq[-1]
    assert l.posterior == l.prior + l.likelihood and l.parameters.shape == (d,) and l.qoi is None
    tail = seq[2:]
    assert len(tail) == n - 2 and np.array_equal(tail[0].parameters, seq[2].parameters)
    both = seq[5:] + seq[:1]
    assert len(both) == 3
    assert np.array([lk.parameters for lk in both]).shape == (3, d)
    with pytest.raises(IndexError):
        seq[n]


def test_get_samples_matches_reference_layout():
    n, d, m = 9, 2, 3
    rng = np.random.default_rng(2)
    mk = lambda: LinkSequence(rng.standard_normal((n, d)), rng.standard_normal(n), rng.standard_normal(n),
                              rng.standard_normal((n, m)), np.ones(n, bool))
    res = {"sampler": "DA", "n_chains": 2, "iterations": n, "subchain_length": 3,
           "chain_fine_0": mk(), "chain_fine_1": mk(), "chain_coarse_0": mk(), "chain_coarse_1": mk()}
    s = tda.get_samples(res, "parameters", "fine", burnin=2)
    assert s["iterations"] == n - 2 and s["dimension"] == d and s["chain_1"].shape == (n - 2, d)
    st = tda.get_samples(res, "stats", "coarse")
    assert st["dimension"] == 3
    np.testing.assert_allclose(st["chain_0"][:, 2], st["chain_0"][:, 0] + st["chain_0"][:, 1])
    mo = tda.get_samples(res, "model_output")
    assert mo["dimension"] == m


def test_to_inference_data_builds_the_reference_groups(monkeypatch):
    """diagnostics.py:6-69: groups posterior / posterior_predictive / qoi / sample_stats with the
    variables x{i}, obs_{i}, qoi_{i}, prior / likelihood / posterior and dims (chain, draw).
    arviz and xarray are not installed in this image: two recording stand-ins take their place."""
    import sys
    import types
    xr = types.ModuleType("xarray")
    az = types.ModuleType("arviz")

    class Dataset:
        def __init__(self, data_vars, coords):
            self.data_vars, self.coords = data_vars, coords

    class InferenceData:
        def __init__(self, **groups):
            self.groups = groups

    xr.Dataset, az.InferenceData = Dataset, InferenceData
    monkeypatch.setitem(sys.modules, "xarray", xr)
    monkeypatch.setitem(sys.modules, "arviz", az)
    n, d, m = 12, 2, 3
    rng = np.random.default_rng(4)
    mk = lambda: LinkSequence(rng.standard_normal((n, d)), rng.standard_normal(n), rng.standard_normal(n),
                              rng.standard_normal((n, m)), np.ones(n, bool))
    res = {"sampler": "MH", "n_chains": 3, "iterations": n, "chain_0": mk(), "chain_1": mk(), "chain_2": mk()}
    idata = tda.to_inference_data(res, burnin=4, parameter_names=["b", "m"])
    assert set(idata.groups) == {"posterior", "posterior_predictive", "qoi", "sample_stats"}
    post = idata.groups["posterior"]
    assert list(post.data_vars) == ["b", "m"]
    dims, arr = post.data_vars["m"]
    assert dims == ["chain", "draw"] and arr.shape == (3, n - 4)
    np.testing.assert_array_equal(arr[1], res["chain_1"].parameters[4:, 1])
    assert list(idata.groups["posterior_predictive"].data_vars) == ["obs_0", "obs_1", "obs_2"]
    assert list(idata.groups["sample_stats"].data_vars) == ["prior", "likelihood", "posterior"]
    assert list(idata.groups["qoi"].data_vars) == ["qoi_0"]
    assert post.coords["draw"][1] == list(range(n - 4))


def test_ess_and_rhat_on_ar1_chains():
    rng = np.random.default_rng(3)
    n, m, rho = 20000, 4, 0.8
    x = np.zeros((m, n))
    e = rng.standard_normal((m, n))
    for i in range(1, n):
        x[:, i] = rho * x[:, i - 1] + e[:, i]
    expected = m * n * (1 - rho) / (1 + rho)
    assert abs(tda.ess_bulk(x) / expected - 1) < 0.2
    assert abs(tda.rhat(x) - 1) < 0.01
    x[0] += 3.0
    assert tda.rhat(x) > 1.2


def test_model_classes_keep_the_reference_model_protocol():
    mdl = tda.Poisson1D(64, 4, 31)
    out = mdl(np.array([0.1, -0.2, 0.05, 0.0]))
    assert out.shape == (31,) and np.all(out > 0)
    # constant conductivity k=1: u = x(1-x)/2 exactly at the nodes
    u = tda.Poisson1D(32, 2, 31)(np.zeros(2))
    x = np.arange(1, 32) / 32
    np.testing.assert_allclose(u, x * (1 - x) / 2, rtol=1e-12)
    r = tda.Rosenbrock(1, 10)
    th = np.array([0.3, -0.4])
    g = r.gradient(th, np.array([1.0]))
    h = 1e-6
    fd = np.array([(r(th + h * np.eye(2)[i])[0] - r(th - h * np.eye(2)[i])[0]) / (2 * h) for i in range(2)])
    np.testing.assert_allclose(g, fd, rtol=1e-6)


def _compact_numpy(theta, prior, like, acc, bounds):
    """NumPy statement of the device compaction (tda_compact_*): dense [nrec, C(, d)] history -> chunks."""
    from tinyda_b200.engine import CompactChunk
    chunks = []
    for k, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
        fl = acc[a:b].copy()
        if k == 0:
            fl[0] = 1
        ch = CompactChunk(b - a)
        ch.accept = fl.astype(np.uint8)
        counts = fl.sum(axis=0)
        ch.offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        rows_t, rows_p, rows_l = [], [], []
        for c in range(acc.shape[1]):
            sel = np.nonzero(fl[:, c])[0] + a
            rows_t.append(theta[sel, c]); rows_p.append(prior[sel, c]); rows_l.append(like[sel, c])
        ch.theta = np.concatenate(rows_t); ch.prior = np.concatenate(rows_p); ch.like = np.concatenate(rows_l)
        chunks.append(ch)
    return chunks


def test_compact_history_expands_to_the_dense_history():
    """A rejected step repeats the previous Link (chain.py:116, :434): the compacted form (accept flags +
    accepted rows, block by block) must expand to the dense history bit for bit, per chain and in bulk."""
    from tinyda_b200.link import CompactHistory, SampleResult, LinkSequence
    rng = np.random.default_rng(5)
    nrec, C, d = 41, 7, 3
    acc = (rng.random((nrec, C)) < 0.3).astype(np.uint8)
    acc[:, 2] = 0                                  # a chain that never moves
    acc[:, 4] = 1                                  # one that always does
    theta = np.zeros((nrec, C, d)); prior = np.zeros((nrec, C)); like = np.zeros((nrec, C))
    cur_t, cur_p, cur_l = rng.standard_normal((C, d)), rng.standard_normal(C), rng.standard_normal(C)
    for r in range(nrec):
        mv = acc[r].astype(bool) | (r == 0)
        cur_t = np.where(mv[:, None], rng.standard_normal((C, d)), cur_t)
        cur_p = np.where(mv, rng.standard_normal(C), cur_p)
        cur_l = np.where(mv, rng.standard_normal(C), cur_l)
        theta[r], prior[r], like[r] = cur_t, cur_p, cur_l
    for bounds in ([0, nrec], [0, 1, 9, 10, 30, nrec]):
        hist = CompactHistory(C)
        for ch in _compact_numpy(theta, prior, like, acc, bounds):
            hist.append(ch)
        assert hist.n_records == nrec
        for c in range(C):
            seq = hist.chain(c)
            assert np.array_equal(seq.parameters, theta[:, c]) and np.array_equal(seq.prior, prior[:, c])
            assert np.array_equal(seq.likelihood, like[:, c])
            assert np.array_equal(seq.accepted[1:], acc[1:, c].astype(bool)) and not seq.accepted[0]
        assert np.array_equal(hist.dense("theta"), np.swapaxes(theta, 0, 1))
        assert np.array_equal(hist.dense("like", burnin=4), like[4:].T)
    # the result dict materialises a chain when its key is first read
    res = SampleResult({"sampler": "MH", "n_chains": C, "iterations": nrec})
    res.add_chains("chain_{}", 0, C, hist.chain)
    res.history, res.local_chains = hist, (0, C)
    assert not dict.__contains__(res, "chain_3") and "chain_3" in res and "chain_7" not in res and "chain_x" not in res
    assert isinstance(res["chain_3"], LinkSequence) and dict.__contains__(res, "chain_3")
    assert len(res) == 3 + C and list(res)[:3] == ["sampler", "n_chains", "iterations"] and list(res)[-1] == "chain_%d" % (C - 1)
    assert all(isinstance(v, LinkSequence) for k, v in res.items() if k.startswith("chain_"))
    with pytest.raises(KeyError):
        res["chain_99"]
    import tinyda_b200 as tda
    s = tda.get_samples(res, "parameters", burnin=2)
    assert np.array_equal(s["chain_5"], theta[2:, 5]) and s["iterations"] == nrec - 2
    st = tda.get_samples(res, "stats")
    assert np.array_equal(st["chain_1"][:, 2], prior[:, 1] + like[:, 1])


def test_models_return_output_and_qoi_like_the_reference_protocol():
    """posterior.py:95-105: a model may return (output, qoi)."""
    import tinyda_b200 as tda
    G = np.arange(12.0).reshape(4, 3)
    Q = np.array([[1.0, 0, 0, -1.0], [0.5, 0.5, 0.5, 0.5]])
    m = tda.LinearModel(G, qoi=(Q, [1.0, 2.0]))
    out, q = m(np.array([1.0, 2.0, 3.0]))
    assert np.allclose(q, Q @ out + [1.0, 2.0])
    low = m.lower()
    assert low["n_qoi"] == 2 and low["qoi_Q"].shape == (2, 4)
    assert tda.LinearModel(G).lower()["n_qoi"] == 0 and isinstance(tda.LinearModel(G)(np.zeros(3)), np.ndarray)


def _ess_sums_numpy(x, n_lag=None):
    """NumPy statement of tda_ess_sums for one parameter: x [n_chains, n_draws] -> (sums, folded)."""
    from tinyda_b200.diagnostics import _split, _rank_normalise, _autocov
    def sums_of(z, n_lag):
        m, n = z.shape
        ac = _autocov(z).sum(axis=0)[:n_lag]
        means = z.mean(axis=1)
        return np.concatenate([ac, [means.sum(), (means ** 2).sum(), m, n]])
    xs = _split(np.asarray(x, dtype=np.float64))
    n_half = xs.shape[1]
    n_lag = n_half if n_lag is None else n_lag
    s = sums_of(_rank_normalise(xs), n_lag)
    f = sums_of(_rank_normalise(np.abs(xs - np.median(xs))), 1)[[0, 1, 2, 3]]
    return s, f


def test_ess_and_rhat_from_sums_equal_the_array_estimators():
    """diagnostics.ess_rhat_from_sums (fed by the device kernels) against ess_bulk / rhat on host arrays."""
    import tinyda_b200 as tda
    from tinyda_b200.diagnostics import ess_rhat_from_sums
    rng = np.random.default_rng(3)
    m, n, rho = 16, 400, 0.8
    x = np.zeros((m, n))
    for t in range(1, n):
        x[:, t] = rho * x[:, t - 1] + np.sqrt(1 - rho ** 2) * rng.standard_normal(m)
    x[:, 100:140] = x[:, [100]]                      # a stretch of rejections: ties
    s, f = _ess_sums_numpy(x)
    ess, rh = ess_rhat_from_sums(s[None, :], f[None, :])
    assert abs(ess[0] / tda.ess_bulk(x) - 1) < 1e-9 and abs(rh[0] / tda.rhat(x) - 1) < 1e-9
    # truncated lags: same answer once the window covers the positive part of the autocorrelation
    s2, f2 = _ess_sums_numpy(x, n_lag=120)
    ess2, _ = ess_rhat_from_sums(s2[None, :], f2[None, :])
    assert abs(ess2[0] / ess[0] - 1) < 0.05


def test_dream_sync_every_extension_is_validated_and_lowered():
    """DREAM(..., sync_every=K) (bounded staleness of the shared archive): the reference's signature is
    unchanged by default, a non-positive value is refused, the value reaches the lowered spec, DREAMZ has no such
    option."""
    import scipy.stats as stats
    import tinyda_b200 as tda
    from tinyda_b200.lowering import lower_problem
    prior = stats.multivariate_normal(np.zeros(3), np.eye(3))
    post = tda.Posterior(prior, tda.GaussianLogLike(np.zeros(4), 0.1 * np.eye(4)), tda.LinearModel(np.ones((4, 3))))
    assert int(lower_problem([post], tda.DREAM(M0=5))["proposal"]["sync_every"]) == 1
    assert int(lower_problem([post], tda.DREAM(M0=5, sync_every=6))["proposal"]["sync_every"]) == 6
    with pytest.raises(ValueError):
        tda.DREAM(M0=5, sync_every=0)
    with pytest.raises(TypeError):
        tda.DREAMZ(M0=5, sync_every=2)
    assert "sync_every" not in lower_problem([post], tda.DREAMZ(M0=5))["proposal"]
