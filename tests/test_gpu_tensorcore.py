"""tcgen05 tensor-core path: self-test GEMM (descriptor / TMEM conventions, 3xTF32 accuracy)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(A, B, a_in_tmem, split):
    from tinyda_b200._lib import lib, check
    N = B.shape[1]
    D = np.zeros((128, N), dtype=np.float32)
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    check(lib.tda_tc_gemm_selftest(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), N,
                                   D.ctypes.data_as(C.c_void_p), a_in_tmem, split))
    return D


@pytest.mark.parametrize("a_in_tmem", [1, 0])
@pytest.mark.parametrize("N", [64, 128, 8, 256])
def test_tf32x3_gemm_matches_fp64(a_in_tmem, N):
    rng = np.random.default_rng(N + a_in_tmem)
    A = rng.standard_normal((128, 64)).astype(np.float32)
    B = (rng.standard_normal((64, N)) / 8).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    D = _gemm(A, B, a_in_tmem, 1)
    err = np.abs(D - ref).max()
    scale = np.abs(ref).max()
    assert err < 2e-6 * scale + 1e-6, (err, scale)
    # a single TF32 pass on B is ~1e-3 relative: the split matters
    D1 = _gemm(A, B, a_in_tmem, 0)
    err1 = np.abs(D1 - ref).max()
    assert err1 < 5e-3 * scale and err1 > 5 * err


def _gemm16(A, B, a_in_tmem):
    from tinyda_b200._lib import lib, check
    N = B.shape[1]
    D = np.zeros((128, N), dtype=np.float32)
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    check(lib.tda_tc16_gemm_selftest(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), N,
                                     D.ctypes.data_as(C.c_void_p), a_in_tmem))
    return D


@pytest.mark.parametrize("a_in_tmem", [1, 0])
@pytest.mark.parametrize("N", [64, 128, 16, 256])
def test_fp16_split_gemm_matches_fp64(a_in_tmem, N):
    """kind::f16 MMAs on two-term fp16 splits at power-of-two scales: fp32-grade accuracy, with the
    A operand packed in TMEM (theta) or canonical in shared memory (the normals)."""
    rng = np.random.default_rng(100 + N + a_in_tmem)
    A = rng.standard_normal((128, 64)).astype(np.float32)
    A[:, 3] *= 1e-3                                  # a column far below the matrix scale
    B = (rng.standard_normal((64, N)) / 8).astype(np.float32)
    B[5, :] *= 1e-4
    ref = A.astype(np.float64) @ B.astype(np.float64)
    D = _gemm16(A, B, a_in_tmem)
    err = np.abs(D - ref).max()
    scale = np.abs(ref).max()
    assert err < 2e-6 * scale + 1e-6, (err, scale)


# ---- tensor-core Delayed-Acceptance kernels --------------------------------------------------
KERNELS = ["tc", "tc16", "tcr"]


def _cfg2_engine(C, kernel, seed=9, rng="philox", streams=None, iters=30, theta0=None, chain_offset=0, store=None):
    from tinyda_b200 import lower_problem
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    from tinyda_b200.workloads import cfg2_da
    w = cfg2_da()
    spec = lower_problem(w["posteriors"], w["proposal"], 10)
    if theta0 is None:
        theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
    eng = Engine(spec, C, dtype="float32", rng=rng, seed=seed, streams=streams,
                 store=[STORE_NONE, STORE_STATS] if store is None else store,
                 capacity_iterations=iters, chain_offset=chain_offset)
    eng.select_kernel(kernel)
    eng.init(theta0)
    return eng, w


@pytest.mark.parametrize("kernel", KERNELS)
def test_tc_kernel_matches_reference_trajectory_until_near_tie(kernel):
    """cfg2 golden fixture (unmodified reference, injected streams) vs the tcgen05 kernels."""
    import golden_io
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    g = golden_io.load("da_pcn_cfg2")
    C, iters = g["theta0"].shape[0], g["iterations"]
    eng = Engine(g["spec"], C, dtype="float32", rng="injected", streams=(g["z"], g["u"]),
                 store=[STORE_NONE, STORE_STATS], capacity_iterations=iters)
    eng.select_kernel(kernel)
    eng.init(g["theta0"])
    eng.run(iters)
    acc = eng.fetch(1, "accept").T.astype(bool)
    th = np.transpose(eng.fetch(1, "theta"), (2, 0, 1)).astype(np.float64)
    lk = eng.fetch(1, "like").T.astype(np.float64)
    ref = g["ref"][1]
    diff = acc != ref["acc"]
    first = np.where(diff.any(axis=1), diff.argmax(axis=1), acc.shape[1])
    assert first.min() >= 10, first            # long common prefix on both chains
    for c in range(C):
        k = int(first[c])
        scale = np.abs(ref["theta"][c]).max()
        assert np.abs(th[c, :k] - ref["theta"][c, :k]).max() <= 1e-5 * scale        # north_star's fp32 tolerance
        np.testing.assert_allclose(lk[c, :k], ref["like"][c, :k], rtol=1e-5, atol=1e-5 * np.abs(ref["like"][c]).max())


@pytest.mark.parametrize("kernel", KERNELS)
def test_tc_kernel_agrees_with_generic_fp32_kernel(kernel):
    C, iters = 512, 30
    a, _ = _cfg2_engine(C, kernel, iters=iters)
    b, _ = _cfg2_engine(C, "generic", iters=iters)
    if kernel in ("tc16", "tcr"):
        b.set_z_round(True)            # same normal stream (fp16 grid) for the generic kernel
    a.run(iters)
    b.run(iters)
    acc_a, acc_b = a.fetch(1, "accept"), b.fetch(1, "accept")
    th_a, th_b = a.fetch(1, "theta"), b.fetch(1, "theta")
    same = (acc_a == acc_b)
    assert same.mean() > 0.97, same.mean()
    # chains whose decisions all agree follow the same trajectory to float32 accuracy
    ok = same.all(axis=0)
    assert ok.mean() > 0.6
    scale = np.abs(th_b).max()
    assert np.abs(th_a[:, :, ok] - th_b[:, :, ok]).max() < 2e-3 * scale
    assert np.array_equal(a.get("cursors")[0], b.get("cursors")[0])
    ca, cb = a.get("accept_counts"), b.get("accept_counts")
    assert abs(ca[0].mean() - cb[0].mean()) < 0.05 * cb[0].mean() + 1


@pytest.mark.parametrize("kernel", KERNELS)
def test_tc_kernel_resume_and_sharding_exact(kernel):
    C = 512
    def run(lo, hi, splits):
        from tinyda_b200.workloads import cfg2_da
        theta0 = cfg2_da()["prior"].rvs(C, random_state=np.random.default_rng(1))
        eng, _ = _cfg2_engine(hi - lo, kernel, iters=sum(splits), theta0=theta0[lo:hi], chain_offset=lo)
        for s in splits:
            eng.run(s)
        return eng.fetch(1, "theta"), eng.fetch(1, "like")
    th_a, lk_a = run(0, C, [12])
    th_b, lk_b = run(0, C, [5, 7])
    assert np.array_equal(th_a, th_b) and np.array_equal(lk_a, lk_b)
    th_c, lk_c = run(256, 512, [12])
    assert np.array_equal(th_a[:, :, 256:512], th_c)


@pytest.mark.parametrize("kernel", ["tc16", "tcr"])
def test_tc16_iteration_blocks_hand_chains_between_sms_exactly(monkeypatch, kernel):
    """tc16 cuts a launch into (tile pair, iteration block) units dealt round-robin over the SMs; with
    more pairs than SMs consecutive blocks of a pair run on different SMs and the state travels through
    global memory.  The result must not depend on the block count, bit for bit."""
    import os
    from tinyda_b200.workloads import cfg2_da
    C, iters = 150 * 256 + 40, 7                      # 151 tile pairs > 148 SMs, ragged last pair
    theta0 = cfg2_da()["prior"].rvs(C, random_state=np.random.default_rng(4))
    outs = []
    for blocks in ("1", "3", "7"):
        monkeypatch.setenv("TDA_TC16_BLOCKS", blocks)
        eng, _ = _cfg2_engine(C, kernel, iters=iters, theta0=theta0)
        eng.run(3)
        eng.run(iters - 3)
        outs.append((eng.fetch(1, "theta"), eng.fetch(1, "like"), eng.fetch(1, "accept"), eng.get("accept_counts"),
                     eng.get("cursors"), eng.get("moments")))
        eng.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    assert outs[0][2][1:].mean() > 0.3


@pytest.mark.parametrize("kernel", ["tc16", "tcr"])
def test_tc16_philox_streams_fed_to_the_oracle(kernel):
    """Production mode of the fp16-split kernel: its z16 / uniform streams exported with
    tda_fill_streams and fed to the CPU oracle reproduce the accept decisions and (to float32
    accuracy) the states, chain by chain until the first near-tie."""
    from oracle import tinyda_oracle as orc
    C, iters = 64, 12
    eng, w = _cfg2_engine(C, kernel, iters=iters, chain_offset=7)
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
    eng.run(iters)
    assert eng.kernel() == kernel
    z, u = eng.fill_streams(iters * 10 * 64, iters * 11)
    zh = (z * 4096).astype(np.float16).astype(np.float64) / 4096
    assert np.array_equal(zh, z)                                  # the stream lies on the fp16 grid
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    from tinyda_b200 import lower_problem
    spec = lower_problem(w["posteriors"], w["proposal"], 10)
    out, _ = orc.run_chains(spec, theta0, z, u, iters)
    acc = eng.fetch(1, "accept").T.astype(bool)
    th = np.transpose(eng.fetch(1, "theta"), (2, 0, 1)).astype(np.float64)
    ref_acc, ref_th = out[1]["acc"], out[1]["theta"]
    assert acc.shape == ref_acc.shape
    diff = acc != ref_acc
    assert diff.mean() < 0.02, diff.mean()
    ok = ~diff.any(axis=1)
    assert ok.mean() > 0.7
    scale = np.abs(ref_th).max()
    assert np.abs(th[ok] - ref_th[ok]).max() < 2e-4 * scale


@pytest.mark.parametrize("kernel", KERNELS)
def test_tc_kernel_conjugate_posterior_full_shape(kernel):
    """cfg2 at its real shape (64 params, 1024/128 observations, J=10, 8192 chains) on the tcgen05
    kernel with a pCN step tuned for stationarity (beta = 0.004): mean within MCSE and variance
    within 5% of the closed-form posterior, starting 2 sd away in every coordinate."""
    from test_gpu_sample_api import _da_conjugate_check
    rc, rf = _da_conjugate_check(kernel, 1024, 0.004, 3000, 3000)
    rc2, rf2 = _da_conjugate_check(kernel, 256, 0.02, 1500, 3000)
    assert rc > rc2          # the smaller step accepts more often


# ---- tc16 with the reference's default storage: chain_coarse_i and Link.model_output ---------------
def _link_fields(eng, level):
    f = {k: eng.fetch(level, k).astype(np.float64) for k in ("theta", "prior", "like", "output")}
    f["accept"] = eng.fetch(level, "accept")
    return f


def test_tc16_full_storage_is_consistent_and_does_not_change_the_chain():
    """tda.sample's defaults (store_coarse_chain=True, Link.model_output kept, sampler.py:421-436) on the
    fp16-split kernel: the kernel records the coarse parameters, log-likelihood and accept flag of every
    coarse step; Link.prior of the coarse records and Link.model_output of both levels are rebuilt from
    the recorded parameters when fetched.  Every record must be a consistent Link (posterior.py:78-110),
    and storing more must not move the chain by a bit."""
    from tinyda_b200.engine import STORE_FULL
    C, iters, J = 512, 12, 10
    a, w = _cfg2_engine(C, "tc16", iters=iters, store=[STORE_FULL, STORE_FULL])
    b, _ = _cfg2_engine(C, "tc16", iters=iters)
    assert a.kernel() == "tc16"
    a.run(5); a.run(7)                       # two launches: the pending-record range must extend
    b.run(iters)
    assert np.array_equal(a.fetch(1, "theta"), b.fetch(1, "theta"))
    assert np.array_equal(a.fetch(1, "like"), b.fetch(1, "like"))
    assert np.array_equal(a.fetch(1, "accept"), b.fetch(1, "accept"))
    assert list(a.n_records()[:2]) == [J * iters, iters + 1]
    G, y, s2, prior = w["G"], w["y"], w["sigma2"], w["prior"]
    idx = np.arange(0, 1024, 8)
    for level, (Gl, yl) in enumerate([(G[idx], y[idx]), (G, y)]):
        f = _link_fields(a, level)
        th = np.transpose(f["theta"], (0, 2, 1))                    # [rec, chain, d]
        F = np.transpose(f["output"], (0, 2, 1))                    # [rec, chain, m]
        F_ref = th @ Gl.T
        assert np.abs(F - F_ref).max() < 1e-5 * np.abs(F_ref).max()
        like_ref = -0.5 * ((F_ref - yl) ** 2).sum(axis=2) / s2
        np.testing.assert_allclose(f["like"], like_ref, rtol=2e-4, atol=2e-2)
        prior_ref = prior.logpdf(th.reshape(-1, 64)).reshape(th.shape[:2])
        np.testing.assert_allclose(f["prior"], prior_ref, rtol=2e-4, atol=2e-2)
    # coarse accept flags describe the recorded coarse states: a rejected step repeats the previous record
    f0 = _link_fields(a, 0)
    rej = f0["accept"][1:] == 0
    same = (f0["theta"][1:] == f0["theta"][:-1]).all(axis=1)
    within = (np.arange(1, J * iters) % J != 0)[:, None]             # not the first step of a subchain
    assert np.array_equal((rej & within), (same & within))
    acc = a.get("accept_counts")
    assert np.array_equal(acc[0], f0["accept"].sum(axis=0))


def test_tc16_full_storage_agrees_with_generic_kernel_and_alternates_with_it():
    from tinyda_b200.engine import STORE_FULL
    C, iters, J = 512, 10, 10
    a, w = _cfg2_engine(C, "tc16", iters=2 * iters, store=[STORE_FULL, STORE_FULL])
    b, _ = _cfg2_engine(C, "generic", iters=2 * iters, store=[STORE_FULL, STORE_FULL])
    b.set_z_round(True)
    a.run(iters)
    b.run(iters)
    fa, fb = _link_fields(a, 0), _link_fields(b, 0)
    ok = (fa["accept"] == fb["accept"]).all(axis=0) & (a.fetch(1, "accept") == b.fetch(1, "accept")).all(axis=0)
    assert ok.mean() > 0.6
    for k, tol in (("theta", 2e-3), ("output", 2e-3)):
        scale = np.abs(fb[k]).max()
        assert np.abs(fa[k][:, :, ok] - fb[k][:, :, ok]).max() < tol * scale, k
    for k in ("prior", "like"):
        np.testing.assert_allclose(fa[k][:, ok], fb[k][:, ok], rtol=1e-3, atol=0.3)
    # the generic kernel continues the tc16 run on the same buffers: its records carry the model
    # output of the current state, which tc16 does not keep -- the engine rebuilds it before the launch
    a.select_kernel("generic")
    a.set_z_round(True)
    a.run(iters)
    G = w["G"]
    idx = np.arange(0, 1024, 8)
    for level, Gl in enumerate([G[idx], G]):
        th = np.transpose(a.fetch(level, "theta").astype(np.float64), (0, 2, 1))
        F = np.transpose(a.fetch(level, "output").astype(np.float64), (0, 2, 1))
        F_ref = th @ Gl.T
        assert np.abs(F - F_ref).max() < 1e-5 * np.abs(F_ref).max(), level


# ---- tc16 on shapes that are not the kernel's native 64 / 16k / 64k: zero-padded operands -------
@pytest.mark.parametrize("d,m_c,m_f", [(32, 100, 1000), (48, 128, 1024), (16, 7, 70), (64, 120, 1900)])
def test_tc16_padded_shapes_agree_with_generic_kernel(d, m_c, m_f):
    """d in {16, 32, 48, 64}, any m_c <= 128, any m_f <= 1920: the host pads the operators with zero
    rows / columns (prepare), the kernel only touches the d parameter rows that exist, and the z16
    stream is consumed d normals per coarse step -- the same positions the generic kernel reads."""
    from tinyda_b200 import lower_problem
    from tinyda_b200.engine import Engine, STORE_FULL
    from tinyda_b200.workloads import cfg2_da
    w = cfg2_da(d=d, m_f=m_f, m_c=m_c)
    J = 10
    spec = lower_problem(w["posteriors"], w["proposal"], J)
    C, iters = 512, 20
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1)).reshape(C, d)
    engs = []
    for kernel in ("tc16", "generic"):
        e = Engine(spec, C, dtype="float32", rng="philox", seed=11, store=[STORE_FULL, STORE_FULL], capacity_iterations=iters)
        e.select_kernel(kernel)
        e.init(theta0)
        if kernel == "generic":
            e.set_z_round(True)
        e.run(iters)
        engs.append(e)
    a, b = engs
    assert a.kernel() == "tc16"
    acc_a, acc_b = a.fetch(1, "accept"), b.fetch(1, "accept")
    cacc_a, cacc_b = a.fetch(0, "accept"), b.fetch(0, "accept")
    assert (acc_a == acc_b).mean() > 0.97 and (cacc_a == cacc_b).mean() > 0.97
    ok = (acc_a == acc_b).all(axis=0) & (cacc_a == cacc_b).all(axis=0)
    assert ok.mean() > 0.6, ok.mean()
    for level in (0, 1):
        for k, tol in (("theta", 2e-3), ("output", 2e-3)):
            xa, xb = a.fetch(level, k), b.fetch(level, k)
            assert xa.shape == xb.shape
            assert np.abs(xa[:, :, ok] - xb[:, :, ok]).max() < tol * np.abs(xb).max(), (level, k)
        for k in ("prior", "like"):
            np.testing.assert_allclose(a.fetch(level, k)[:, ok], b.fetch(level, k)[:, ok], rtol=1e-3, atol=0.3)
    assert np.array_equal(a.get("cursors")[0], b.get("cursors")[0])
    # the links are consistent in float64 as well
    idx = np.arange(0, m_f, m_f // m_c)[:m_c]
    th = np.transpose(a.fetch(0, "theta").astype(np.float64), (0, 2, 1))
    F_ref = th @ w["G"][idx].T
    like_ref = -0.5 * ((F_ref - w["y"][idx]) ** 2).sum(axis=2) / w["sigma2"]
    np.testing.assert_allclose(a.fetch(0, "like"), like_ref, rtol=2e-4, atol=2e-2)
    for e in engs:
        e.close()


@pytest.mark.parametrize("kind", ["diagonal", "dense"])
def test_tc16_diagonal_and_dense_likelihoods_folded_into_the_operators(kind):
    """DiagonalGaussianLogLike / DefaultGaussianLogLike (distributions.py:304-315, :246-301) on the
    tensor-core kernel: prepare() whitens the linear models and the data with the Cholesky factor of the
    precision, the kernel then scores an isotropic unit-variance residual.  Same decisions as the generic
    kernel (which evaluates r^T prec r directly) and log-likelihoods that agree with float64."""
    from scipy import stats
    import tinyda_b200 as tda
    from tinyda_b200 import lower_problem
    from tinyda_b200.engine import Engine, STORE_FULL
    from tinyda_b200.workloads import exp_cov
    from tinyda_b200.models import LinearModel
    rng = np.random.default_rng(21)
    d, m_f, m_c, J = 64, 512, 64, 10
    prior = stats.multivariate_normal(np.zeros(d), exp_cov(d))
    G = rng.standard_normal((m_f, d)) / 8
    y = G @ prior.rvs(random_state=rng) + 0.1 * rng.standard_normal(m_f)
    sd = 0.1 * (1 + 0.5 * rng.random(m_f))
    if kind == "diagonal":
        cov = np.diag(sd ** 2)
    else:
        i = np.arange(m_f)
        cov = np.outer(sd, sd) * (0.4 ** np.abs(i[:, None] - i[None, :]))      # AR(1)-correlated noise
    idx = np.arange(0, m_f, m_f // m_c)
    posts = [tda.Posterior(prior, tda.GaussianLogLike(y[idx], cov[np.ix_(idx, idx)]), LinearModel(G[idx])),
             tda.Posterior(prior, tda.GaussianLogLike(y, cov), LinearModel(G))]
    spec = lower_problem(posts, tda.CrankNicolson(scaling=0.05), J)
    C, iters = 256, 12
    theta0 = prior.rvs(C, random_state=np.random.default_rng(1))
    engs = []
    for kernel in ("tc16", "generic"):
        e = Engine(spec, C, dtype="float32", rng="philox", seed=5, store=[STORE_FULL, STORE_FULL], capacity_iterations=iters)
        e.select_kernel(kernel)
        e.init(theta0)
        if kernel == "generic":
            e.set_z_round(True)
        e.run(iters)
        engs.append(e)
    a, b = engs
    assert a.kernel() == "tc16"
    ok = (a.fetch(1, "accept") == b.fetch(1, "accept")).all(axis=0) & (a.fetch(0, "accept") == b.fetch(0, "accept")).all(axis=0)
    assert ok.mean() > 0.6, ok.mean()
    prec = [np.linalg.inv(cov[np.ix_(idx, idx)]), np.linalg.inv(cov)]
    for level, (Gl, yl) in enumerate([(G[idx], y[idx]), (G, y)]):
        th_a, th_b = a.fetch(level, "theta"), b.fetch(level, "theta")
        assert np.abs(th_a[:, :, ok] - th_b[:, :, ok]).max() < 2e-3 * np.abs(th_b).max()
        r = np.transpose(th_a.astype(np.float64), (0, 2, 1)) @ Gl.T - yl
        like_ref = -0.5 * np.einsum("rcm,mn,rcn->rc", r, prec[level], r)
        np.testing.assert_allclose(a.fetch(level, "like"), like_ref, rtol=3e-4, atol=3e-2)
        F = np.transpose(a.fetch(level, "output").astype(np.float64), (0, 2, 1))
        assert np.abs(F - (r + yl)).max() < 1e-5 * np.abs(r + yl).max()       # Link.model_output is the unwhitened F
    for e in engs:
        e.close()


@pytest.mark.parametrize("kernel", ["tc16", "tc"])
def test_tc16_coarse_chain_matches_reference_trajectory_until_near_tie(kernel):
    """chain_coarse_i of the unmodified reference (cfg2 golden fixture, injected streams) against the
    coarse Links the tc16 (and, since round 2, the 3xTF32 tc) kernel records: parameters, log-likelihood, accept
    flags, and the log-prior rebuilt at fetch time, up to the first decision that differs (float32 near-tie)."""
    import golden_io
    from tinyda_b200.engine import Engine, STORE_STATS
    g = golden_io.load("da_pcn_cfg2")
    C, iters, J = g["theta0"].shape[0], g["iterations"], 10
    eng = Engine(g["spec"], C, dtype="float32", rng="injected", streams=(g["z"], g["u"]),
                 store=[STORE_STATS, STORE_STATS], capacity_iterations=iters)
    eng.select_kernel(kernel)
    eng.init(g["theta0"])
    eng.run(iters)
    ref = g["ref"][0]
    acc = eng.fetch(0, "accept").T.astype(bool)
    th = np.transpose(eng.fetch(0, "theta"), (2, 0, 1)).astype(np.float64)
    lk = eng.fetch(0, "like").T.astype(np.float64)
    pr = eng.fetch(0, "prior").T.astype(np.float64)
    assert acc.shape == ref["acc"].shape == (C, J * iters)
    diff = acc != ref["acc"]
    first = np.where(diff.any(axis=1), diff.argmax(axis=1), acc.shape[1])
    assert first.min() >= 10 * J, first                # at least ten fine iterations in common
    for c in range(C):
        k = int(first[c])
        scale = np.abs(ref["theta"][c]).max()
        assert np.abs(th[c, :k] - ref["theta"][c, :k]).max() <= 1e-5 * scale        # north_star's fp32 tolerance
        np.testing.assert_allclose(lk[c, :k], ref["like"][c, :k], rtol=1e-5, atol=1e-5 * np.abs(ref["like"][c]).max())
        np.testing.assert_allclose(pr[c, :k], ref["prior"][c, :k], rtol=2e-4, atol=2e-2)
    eng.close()


# ---- tcr: per-chain / adaptively scaled pCN steps, coarse chain through the whitened records -----------
def _adaptive_cfg2(C, kernel, iters, period=20, store=None, seed=13):
    from tinyda_b200 import lower_problem, CrankNicolson
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    from tinyda_b200.workloads import cfg2_da
    w = cfg2_da()
    spec = lower_problem(w["posteriors"], CrankNicolson(scaling=0.05, adaptive=True, period=period), 10)
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
    eng = Engine(spec, C, dtype="float32", seed=seed, store=[STORE_NONE, STORE_STATS] if store is None else store,
                 capacity_iterations=iters)
    if kernel:
        eng.select_kernel(kernel)
    eng.init(theta0)
    return eng


def test_tcr_runs_adaptive_pcn_and_agrees_with_the_generic_kernel():
    """CrankNicolson(adaptive=True) -- what the reference's notebooks use -- selects the tensor-core kernel
    automatically; step sizes follow proposal.py:228-245 with the window semantics of the generic kernel
    (coarse decisions + one alignment entry per fine iteration)."""
    C, iters = 512, 30                                  # 300 base steps: 15 adaptations with period 20
    a = _adaptive_cfg2(C, None, iters)
    assert a.kernel() == "tcr"
    b = _adaptive_cfg2(C, "generic", iters)
    b.set_z_round(True)
    a.run(13); a.run(iters - 13)                        # a launch boundary inside an adaptation period
    b.run(iters)
    acc_a, acc_b = a.fetch(1, "accept"), b.fetch(1, "accept")
    th_a, th_b = a.fetch(1, "theta"), b.fetch(1, "theta")
    sa, sb = a.get("scaling"), b.get("scaling")
    same = (acc_a == acc_b).all(axis=0) & (a.get("accept_counts")[0] == b.get("accept_counts")[0])
    assert same.mean() > 0.5, same.mean()
    # chains whose every decision agrees adapted identically and follow the same trajectory to float32 accuracy
    np.testing.assert_allclose(sa[same], sb[same], rtol=1e-5)
    scale = np.abs(th_b).max()
    assert np.abs(th_a[:, :, same] - th_b[:, :, same]).max() < 2e-3 * scale
    assert sa.std() > 0                                 # the steps moved away from the common start
    assert np.array_equal(a.get("cursors"), b.get("cursors"))
    a.close(); b.close()


def test_tcr_per_chain_step_sizes():
    """tda_set(TDA_G_SCALING): every chain its own pCN step; the kernel reads them per chain."""
    C, iters = 512, 10
    from tinyda_b200 import lower_problem
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    from tinyda_b200.workloads import cfg2_da
    w = cfg2_da()
    spec = lower_problem(w["posteriors"], w["proposal"], 10)
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
    steps = np.where(np.arange(C) % 2 == 0, 0.02, 0.11)
    out = {}
    for kern in ("tcr", "generic"):
        eng = Engine(spec, C, dtype="float32", seed=3, store=[STORE_NONE, STORE_STATS], capacity_iterations=iters)
        eng.select_kernel(kern)
        if kern == "generic":
            eng.set_z_round(True)
        eng.init(theta0)
        eng.set_scaling(steps)                          # after init: init restores the proposal's own step
        eng.run(iters)
        out[kern] = (eng.fetch(1, "accept"), eng.fetch(1, "theta"), eng.get("accept_counts"))
        eng.close()
    same = (out["tcr"][0] == out["generic"][0]).all(axis=0) & (out["tcr"][2][0] == out["generic"][2][0])
    assert same.mean() > 0.8
    scale = np.abs(out["generic"][1]).max()
    assert np.abs(out["tcr"][1][:, :, same] - out["generic"][1][:, :, same]).max() < 1e-3 * scale
    # smaller steps are accepted more often on the coarse level
    cnt = out["tcr"][2][0]
    assert cnt[0::2].mean() > 1.15 * cnt[1::2].mean()


def test_tcr_coarse_chain_records_agree_with_the_generic_kernel():
    """store_coarse_chain=True on the tcr kernel: coarse Links are recorded whitened and turned into
    theta / log-prior / model output when first fetched."""
    from tinyda_b200.engine import STORE_FULL
    C, iters = 256, 6
    a = _adaptive_cfg2(C, "tcr", iters, store=[STORE_FULL, STORE_FULL])
    b = _adaptive_cfg2(C, "generic", iters, store=[STORE_FULL, STORE_FULL])
    b.set_z_round(True)
    a.run(iters); b.run(iters)
    fa, fb = _link_fields(a, 0), _link_fields(b, 0)
    same = (fa["accept"] == fb["accept"]).all(axis=0) & (a.fetch(1, "accept") == b.fetch(1, "accept")).all(axis=0)
    assert same.mean() > 0.7
    for k in ("theta", "output"):
        scale = np.abs(fb[k]).max()
        assert np.abs(fa[k][:, :, same] - fb[k][:, :, same]).max() < 1e-3 * scale, k
    for k in ("prior", "like"):
        np.testing.assert_allclose(fa[k][:, same], fb[k][:, same], rtol=2e-4, atol=0.05)
    assert fa["theta"].shape[0] == iters * 10
    a.close(); b.close()


def test_tc16_single_tile_modes_are_bit_identical_to_the_paired_mode(monkeypatch):
    """With fewer tile pairs than half the SMs every CTA advances ONE 128-chain tile, with fewer tiles than half
    the SMs one 64-chain half tile (the strong-scaling case: 8192 chains per GPU use 128 SMs instead of 32);
    the chains do not notice."""
    from tinyda_b200.workloads import cfg2_da
    C, iters = 2048 + 100, 9
    theta0 = cfg2_da()["prior"].rvs(C, random_state=np.random.default_rng(4))
    outs = []
    for env in ({"TDA_TC16_NO_SOLO": "1"}, {"TDA_TC16_NO_HALF": "1"}, {}):
        for k in ("TDA_TC16_NO_SOLO", "TDA_TC16_NO_HALF"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng, _ = _cfg2_engine(C, "tc16", iters=iters, theta0=theta0)
        eng.run(4); eng.run(iters - 4)
        outs.append((eng.fetch(1, "theta"), eng.fetch(1, "like"), eng.fetch(1, "accept"), eng.get("accept_counts"), eng.get("cursors")))
        eng.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)


def _cfg2_rwmh(C, kernel, iters, dtype="float32", rng="philox", streams=None, seed=9, adaptive=False):
    """cfg2's problem with a GaussianRandomWalk proposal (dense covariance) instead of pCN."""
    from tinyda_b200 import lower_problem
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    from tinyda_b200.proposal import GaussianRandomWalk
    from tinyda_b200.workloads import cfg2_da, exp_cov
    w = cfg2_da()
    prop = GaussianRandomWalk(C=exp_cov(64, 0.3), scaling=0.02, adaptive=adaptive, period=15)
    spec = lower_problem(w["posteriors"], prop, 10)
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1))
    eng = Engine(spec, C, dtype=dtype, rng=rng, seed=seed, streams=streams, store=[STORE_NONE, STORE_STATS],
                 capacity_iterations=iters)
    eng.select_kernel(kernel)
    eng.init(theta0)
    return eng, spec, theta0


def _prefix_agreement(th_a, acc_a, th_r, acc_r, jump=1e-4):
    """Per chain: records before the first flipped decision -- a differing fine-level accept flag, or a state that
    differs by more than `jump` relative (a flipped COARSE decision inside a subchain moves the state by a proposal
    step without touching the fine-level flags; the coarse chain is not recorded on this kernel).  Returns (records
    before a flip, all records, largest relative state error before the flip)."""
    n, _, C = th_a.shape
    matched, worst = 0, 0.0
    for c in range(C):
        sc = np.abs(th_r[:, :, c]).max() + 1e-30
        err = np.abs(th_a[:, :, c] - th_r[:, :, c]).max(axis=1) / sc
        bad = (acc_a[:, c] != acc_r[:, c]) | (err > jump)
        k = int(np.argmax(bad)) if bad.any() else n
        matched += k
        if k:
            worst = max(worst, float(err[:k].max()))
    return matched, n * C, worst


def test_random_walk_delayed_acceptance_runs_on_the_tensor_cores():
    """GaussianRandomWalk-based two-level DA at cfg2's shape (proposal.py:132-258: theta' = theta + s xi, the
    acceptance is the full posterior ratio): selected automatically onto the 3xTF32 tcgen05 kernel (a third job per
    coarse step, theta' @ LP, gives the proposal's log-prior).  Against the lock-step float32 kernel on the same
    Philox streams, and against the float64 engine fed those streams: identical trajectories up to a first near-tie,
    states within 1e-5 relative before it (north_star's float32 tolerance)."""
    import problems
    C, iters = 512, 25
    a, spec, theta0 = _cfg2_rwmh(C, "auto", iters)
    assert a.kernel() == "tc"
    b, _, _ = _cfg2_rwmh(C, "generic", iters)
    a.run(iters)
    b.run(iters)
    acc_a, acc_b = a.fetch(1, "accept"), b.fetch(1, "accept")
    th_a, th_b = a.fetch(1, "theta"), b.fetch(1, "theta")
    assert (acc_a == acc_b).mean() > 0.97
    assert np.array_equal(a.get("cursors")[0], b.get("cursors")[0])
    ca, cb = a.get("accept_counts"), b.get("accept_counts")
    assert ca[0].mean() > 0.02 * iters * 10 and abs(ca[0].mean() - cb[0].mean()) < 0.05 * cb[0].mean() + 1
    m_g, tot, w_g = _prefix_agreement(th_a, acc_a, th_b, acc_b)
    nz, nu = problems.stream_sizes(spec, iters)
    z, u = a.fill_streams(nz, nu)
    ref, _, _ = _cfg2_rwmh(C, "generic", iters, dtype="float64", rng="injected", streams=(z, u))
    ref.run(iters)
    m_r, _, w_r = _prefix_agreement(th_a, acc_a, ref.fetch(1, "theta"), ref.fetch(1, "accept"))
    print("\ntc (random walk, float32): %d of %d fine records before a first flip against the lock-step float32 kernel "
          "(max relative state error %.2e), %d against the float64 engine on the same streams (%.2e)" % (m_g, tot, w_g, m_r, w_r))
    assert m_g >= 0.8 * tot and m_r >= 0.8 * tot
    assert w_g <= 1e-5 and w_r <= 1e-5
    for e in (a, b, ref):
        e.close()


def test_random_walk_tensor_core_kernel_resume_is_exact():
    outs = []
    for cuts in ([12], [5, 7]):
        eng, _, _ = _cfg2_rwmh(512, "tc", sum(cuts))
        for n in cuts:
            eng.run(n)
        outs.append((eng.fetch(1, "theta"), eng.fetch(1, "like"), eng.fetch(1, "prior"), eng.get("cursors")))
        eng.close()
    for x, y in zip(*outs):
        assert np.array_equal(x, y)


def test_adaptive_random_walk_delayed_acceptance_on_the_tensor_cores():
    """GaussianRandomWalk(adaptive=True) -- what the reference's notebooks use -- at cfg2's shape: the accept window
    (coarse decisions + the alignment entry of every fine iteration) and the step-size rule of proposal.py:228-245
    inside the tcgen05 kernel, across several adaptation periods (period 15, 275 window entries).  Against the
    float64 engine on the same streams: trajectories and adapted step sizes."""
    import problems
    C, iters = 512, 25
    a, spec, theta0 = _cfg2_rwmh(C, "auto", iters, adaptive=True)
    assert a.kernel() == "tc"
    a.run(11)
    a.run(iters - 11)                                   # the window and the counters carry over launches
    acc_a, th_a, sc_a = a.fetch(1, "accept"), a.fetch(1, "theta"), a.get("scaling")
    nz, nu = problems.stream_sizes(spec, iters)
    z, u = a.fill_streams(nz, nu)
    ref, _, _ = _cfg2_rwmh(C, "generic", iters, dtype="float64", rng="injected", streams=(z, u), adaptive=True)
    ref.run(iters)
    acc_r, th_r, sc_r = ref.fetch(1, "accept"), ref.fetch(1, "theta"), ref.get("scaling")
    m, tot, w = _prefix_agreement(th_a, acc_a, th_r, acc_r)
    print("\ntc (adaptive random walk, float32) vs float64: %d of %d fine records before a first flip, max relative state "
          "error %.2e" % (m, tot, w))
    assert m >= 0.8 * tot and w <= 1e-5
    assert np.array_equal(a.get("cursors")[0], ref.get("cursors")[0])
    assert np.ptp(sc_r) > 0 and not np.allclose(sc_r, 0.02)
    # chains that never flipped adapted their step exactly like the float64 run
    n = th_a.shape[0]
    clean = np.array([(acc_a[:, c] == acc_r[:, c]).all() and
                      (np.abs(th_a[:, :, c] - th_r[:, :, c]).max() <= 1e-4 * np.abs(th_r[:, :, c]).max()) for c in range(C)])
    assert clean.mean() > 0.7
    np.testing.assert_allclose(sc_a[clean], sc_r[clean], rtol=2e-5)
    a.close()
    ref.close()


@pytest.mark.parametrize("d,adaptive,m_f,m_c", [(32, True, 256, 64), (48, False, 300, 50), (16, True, 200, 25), (64, True, 1000, 100)])
def test_random_walk_tensor_core_kernel_with_fewer_than_64_parameters(d, adaptive, m_f, m_c):
    """d in {16, 32, 48}: operators zero-padded to 64 rows on the host, the padded parameter columns draw no normals
    (stream positions t d + k as on the lock-step kernel); observation counts that are not multiples of 16 / 64 are
    zero-padded too.  Two-level random-walk DA against the float64 engine on the same streams."""
    import problems
    import scipy.stats as stats
    from tinyda_b200 import lower_problem
    from tinyda_b200.distributions import GaussianLogLike
    from tinyda_b200.engine import Engine, STORE_STATS
    from tinyda_b200.models import LinearModel
    from tinyda_b200.posterior import Posterior
    from tinyda_b200.proposal import GaussianRandomWalk
    from tinyda_b200.workloads import exp_cov
    rng = np.random.default_rng(d)
    C, iters = 256, 20
    prior = stats.multivariate_normal(np.zeros(d), exp_cov(d))
    G = rng.standard_normal((m_f, d)) / np.sqrt(d)
    y = G @ prior.rvs(random_state=rng) + 0.1 * rng.standard_normal(m_f)
    idx = np.arange(0, m_f, m_f // m_c)[:m_c]
    posts = [Posterior(prior, GaussianLogLike(y[idx], 0.01 * np.eye(m_c)), LinearModel(G[idx])),
             Posterior(prior, GaussianLogLike(y, 0.01 * np.eye(m_f)), LinearModel(G))]
    prop = GaussianRandomWalk(C=exp_cov(d, 0.3), scaling=0.03, adaptive=adaptive, period=12)
    spec = lower_problem(posts, prop, 5)
    theta0 = prior.rvs(C, random_state=rng)
    a = Engine(spec, C, dtype="float32", seed=4, store=[STORE_STATS, STORE_STATS], capacity_iterations=iters)
    assert a.kernel() == "tc"
    a.init(theta0)
    a.run(iters)
    nz, nu = problems.stream_sizes(spec, iters)
    z, u = a.fill_streams(nz, nu)
    ref = Engine(spec, C, dtype="float64", rng="injected", streams=(z, u), store=[STORE_STATS, STORE_STATS], capacity_iterations=iters)
    ref.select_kernel("generic")
    ref.init(theta0)
    ref.run(iters)
    m, tot, w = _prefix_agreement(a.fetch(1, "theta"), a.fetch(1, "accept"), ref.fetch(1, "theta"), ref.fetch(1, "accept"))
    mc_, totc, wc_ = _prefix_agreement(a.fetch(0, "theta"), a.fetch(0, "accept"), ref.fetch(0, "theta"), ref.fetch(0, "accept"))
    print("\ntc, d = %d (adaptive=%s): %d of %d fine and %d of %d coarse records before a first flip, max relative state error "
          "%.2e / %.2e" % (d, adaptive, m, tot, mc_, totc, w, wc_))
    assert m >= 0.8 * tot and mc_ >= 0.8 * totc
    assert w <= 1e-5 and wc_ <= 1e-5
    assert np.array_equal(a.get("cursors"), ref.get("cursors"))
    assert a.fetch(0, "accept").mean() > 0.02
    a.close()
    ref.close()


def test_random_walk_tensor_core_kernel_folds_diagonal_and_dense_likelihoods():
    """A dense Gaussian likelihood on the coarse level and a diagonal one on the fine level (distributions.py:246-315)
    are folded into the operator images on the host (Cholesky of the precision); the recorded Links still carry the
    un-whitened model output.  Random-walk DA, d = 32, against the float64 engine on the same streams."""
    import problems
    import scipy.stats as stats
    from tinyda_b200 import lower_problem
    from tinyda_b200.distributions import GaussianLogLike
    from tinyda_b200.engine import Engine, STORE_FULL
    from tinyda_b200.models import LinearModel
    from tinyda_b200.posterior import Posterior
    from tinyda_b200.proposal import GaussianRandomWalk
    from tinyda_b200.workloads import exp_cov
    rng = np.random.default_rng(77)
    d, m_f, m_c, C, iters = 32, 200, 40, 256, 16
    prior = stats.multivariate_normal(np.zeros(d), exp_cov(d))
    G = rng.standard_normal((m_f, d)) / np.sqrt(d)
    var = 0.01 * (1.0 + rng.random(m_f))
    y = G @ prior.rvs(random_state=rng) + np.sqrt(var) * rng.standard_normal(m_f)
    idx = np.arange(0, m_f, m_f // m_c)[:m_c]
    B = rng.standard_normal((m_c, m_c))
    cov_c = 0.01 * (np.eye(m_c) + 0.3 * B @ B.T / m_c)                       # dense coarse covariance
    posts = [Posterior(prior, GaussianLogLike(y[idx], cov_c), LinearModel(G[idx])),
             Posterior(prior, GaussianLogLike(y, np.diag(var)), LinearModel(G))]
    spec = lower_problem(posts, GaussianRandomWalk(C=exp_cov(d, 0.3), scaling=0.03), 5)
    assert [int(lv["lik"]["kind"]) for lv in spec["levels"]] == [2, 1]          # dense, diagonal
    theta0 = prior.rvs(C, random_state=rng)
    a = Engine(spec, C, dtype="float32", seed=4, store=[STORE_FULL, STORE_FULL], capacity_iterations=iters)
    assert a.kernel() == "tc"
    a.init(theta0)
    a.run(iters)
    nz, nu = problems.stream_sizes(spec, iters)
    z, u = a.fill_streams(nz, nu)
    ref = Engine(spec, C, dtype="float64", rng="injected", streams=(z, u), store=[STORE_FULL, STORE_FULL], capacity_iterations=iters)
    ref.select_kernel("generic")
    ref.init(theta0)
    ref.run(iters)
    for l in (0, 1):
        m, tot, w = _prefix_agreement(a.fetch(l, "theta"), a.fetch(l, "accept"), ref.fetch(l, "theta"), ref.fetch(l, "accept"))
        assert m >= 0.8 * tot and w <= 1e-5, (l, m, tot, w)
    # records of chains that never flipped: log-likelihoods and (un-whitened) model outputs of both levels
    for l in (0, 1):
        th_a, th_r = a.fetch(l, "theta"), ref.fetch(l, "theta")
        clean = np.abs(th_a - th_r).max(axis=(0, 1)) <= 1e-4 * np.abs(th_r).max()      # this level's chain never flipped
        assert clean.mean() > 0.7
        la, lr = a.fetch(l, "like")[:, clean], ref.fetch(l, "like")[:, clean]
        np.testing.assert_allclose(la, lr, rtol=2e-4, atol=2e-4 * np.abs(lr).max())
        Fa, Fr = a.fetch(l, "output")[:, :, clean], ref.fetch(l, "output")[:, :, clean]
        np.testing.assert_allclose(Fa, Fr, rtol=1e-4, atol=1e-5 * np.abs(Fr).max())
    a.close()
    ref.close()
