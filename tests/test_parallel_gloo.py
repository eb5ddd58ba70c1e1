"""world_size-2 tests of the multi-GPU host logic on the gloo backend (CPU tensors): chain
sharding, the in-place all-gather of one DREAM archive slot, and the R-hat moment reduction."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tinyda_b200 import parallel


def test_shard_ranges_partition_the_chains():
    for n, w in [(65536, 8), (10, 4), (7, 2), (3, 8)]:
        r = [parallel.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1
    assert parallel.archive_row_to_chain_slot(17, 5) == (3, 2)


def _ar1_chains():
    rng = np.random.default_rng(5)
    C, n, d = 6, 4000, 2
    y = np.zeros((C, n, d))
    e = rng.standard_normal((C, n, d))
    for t in range(1, n):
        y[:, t, 0] = 0.7 * y[:, t - 1, 0] + e[:, t, 0]
        y[:, t, 1] = 0.95 * y[:, t - 1, 1] + e[:, t, 1]
    return y


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.rank_world() == (rank, world)
        for n_chains in (8, 7):
            d = 3
            full = torch.arange(n_chains * d, dtype=torch.float64).reshape(n_chains, d)
            lo, hi = parallel.shard_range(n_chains, rank, world)
            rows = torch.zeros(n_chains, d, dtype=torch.float64)
            rows[lo:hi] = full[lo:hi]
            parallel.allgather_rows(rows, lo, hi, world)
            assert torch.equal(rows, full)
        rng = np.random.default_rng(0)
        C, d, n = 6, 4, 50
        x = rng.standard_normal((C, n, d)) + np.arange(C)[:, None, None] * 0.1
        lo, hi = parallel.shard_range(C, rank, world)
        s1 = x[lo:hi].sum(axis=1).T
        s2 = (x[lo:hi] ** 2).sum(axis=1).T
        out = parallel.allreduce_chain_moments(s1, s2, n)
        # ESS over both ranks' chains: AR(1) series, 3 chains per rank
        y = _ar1_chains()
        ess = parallel.allreduce_ess(y[lo:hi])
        q.put((rank, out["rhat"], out["mean"], out["n_chains"], ess))
    finally:
        dist.destroy_process_group()


def test_two_rank_allgather_and_moment_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    C, d, n = 6, 4, 50
    x = rng.standard_normal((C, n, d)) + np.arange(C)[:, None, None] * 0.1
    W = x.var(axis=1, ddof=1).mean(axis=0)
    B_over_n = x.mean(axis=1).var(axis=0, ddof=1)
    rhat = np.sqrt(((n - 1) / n * W + B_over_n) / W)
    from tinyda_b200.diagnostics import _ess_plain
    y = _ar1_chains()
    ess_one = np.array([_ess_plain(y[:, :, k]) for k in range(y.shape[2])])
    for rank, r, mean, m, ess in res:
        assert m == C
        np.testing.assert_allclose(r, rhat, rtol=1e-12)
        np.testing.assert_allclose(mean, x.mean(axis=(0, 1)), rtol=1e-12)
        np.testing.assert_allclose(ess, ess_one, rtol=1e-9)          # sharded == single process
    # and the estimate is right: ESS/N of an AR(1) chain is (1 - rho) / (1 + rho)
    N = y.shape[0] * y.shape[1]
    assert abs(ess_one[0] / N - 0.3 / 1.7) < 0.03 and abs(ess_one[1] / N - 0.05 / 1.95) < 0.01
