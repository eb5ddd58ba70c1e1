"""Worker for tests/test_gpu_multi.py: launched by torchrun with one process per GPU (NCCL).
Checks, on real GPUs, that sharding a job over ranks changes nothing:
  * DREAM with the shared archive all-gathered over NCCL after every step equals the same job on
    one engine bit for bit (lock-step archive visibility makes the run deterministic);
  * a Delayed-Acceptance job sharded by tda.sample() equals the single-engine run bit for bit;
  * the R-hat moment reduction over NCCL equals the one computed from all chains.
Each rank writes "ok" (or the failure) to <outdir>/rank<k>.txt."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(outdir):
    import torch
    import torch.distributed as dist
    import tinyda_b200 as tda
    from tinyda_b200 import parallel, workloads
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    from tinyda_b200.lowering import lower_problem

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- DREAM, shared archive ----------------------------------------------------------------
    w = workloads.cfg5_dream()
    C, iters, M0, d = 512, 40, 16, 32
    rng = np.random.default_rng(5)
    theta0 = w["prior"].rvs(C, random_state=rng)
    archive0 = w["prior"].rvs(C * M0, random_state=rng).reshape(C, M0, d)
    res = tda.sample(w["posteriors"][0], w["proposal"], iters, n_chains=C, initial_parameters=[t for t in theta0],
                     seed=9, initial_archive=archive0, store_model_output=False, dtype="float32")
    lo, hi = res["local_chains"]
    assert (lo, hi) == parallel.shard_range(C, rank, world)
    # the step's new archive rows travel inside the persistent kernel, through NVLink peer memory
    assert res.dream_exchange == "peer-memory", res.dream_exchange
    mine = np.stack([res["chain_%d" % c].parameters for c in range(lo, hi)])          # [n_local, iters+1, d]
    spec = lower_problem(w["posteriors"], w["proposal"])
    eng = Engine(spec, C, dtype="float32", seed=9, store=STORE_STATS, capacity_iterations=iters,
                 device=local, archive0=archive0)
    eng.init(theta0)
    eng.run(iters)
    one = np.transpose(eng.fetch(0, "theta"), (2, 0, 1))
    eng.close()
    assert np.array_equal(mine, one[lo:hi]), "sharded DREAM differs from the single-engine run"
    assert np.abs(np.diff(mine, axis=1)).max() > 0

    # ---- the same exchange on the lock-step kernel (flag handshake instead of arrival counters; it serves the
    # adaptive crossover distribution and d > 32), engines driven directly ----
    lo_g, hi_g = parallel.shard_range(C, rank, world)
    eg = Engine(spec, hi_g - lo_g, dtype="float32", seed=9, store=STORE_STATS, capacity_iterations=iters, device=local,
                chain_offset=lo_g, n_chains_global=C, archive0=archive0)
    eg.select_kernel("generic")
    assert parallel.connect_dream_peers(eg, rank, world)
    eg.init(theta0[lo_g:hi_g])
    parallel.run_dream_shared(eg, iters, rank, world)
    mine_g = np.transpose(eg.fetch(0, "theta"), (2, 0, 1))
    dist.barrier()
    eg.close()
    eng = Engine(spec, C, dtype="float32", seed=9, store=STORE_STATS, capacity_iterations=iters, device=local, archive0=archive0)
    eng.select_kernel("generic")
    eng.init(theta0)
    eng.run(iters)
    one_g = np.transpose(eng.fetch(0, "theta"), (2, 0, 1))
    eng.close()
    assert np.array_equal(mine_g, one_g[lo_g:hi_g]), "sharded DREAM (lock-step kernel) differs from the single-engine run"

    # ---- bounded staleness: DREAM(sync_every=5), the ranks meet every fifth step only; launch cut off the grid ----
    from tinyda_b200.proposal import DREAM
    spec5 = lower_problem(w["posteriors"], DREAM(M0=16, delta=1, nCR=3, sync_every=5))
    e5 = Engine(spec5, hi_g - lo_g, dtype="float32", seed=9, store=STORE_STATS, capacity_iterations=iters, device=local,
                chain_offset=lo_g, n_chains_global=C, archive0=archive0)
    assert e5.kernel() == "dreamw" and parallel.connect_dream_peers(e5, rank, world)
    e5.init(theta0[lo_g:hi_g])
    parallel.run_dream_shared(e5, 13, rank, world)
    parallel.run_dream_shared(e5, iters - 13, rank, world)
    mine_5 = np.transpose(e5.fetch(0, "theta"), (2, 0, 1))
    dist.barrier()
    e5.close()
    eng = Engine(spec5, C, dtype="float32", seed=9, store=STORE_STATS, capacity_iterations=iters, device=local, archive0=archive0)
    eng.init(theta0)
    eng.run(iters)
    one_5 = np.transpose(eng.fetch(0, "theta"), (2, 0, 1))
    eng.close()
    assert np.array_equal(mine_5, one_5[lo_g:hi_g]), "sharded DREAM(sync_every=5) differs from the single-engine run"
    assert not np.array_equal(one_5, one), "sync_every=5 should not reproduce the lock-step trajectory"

    # ---- Delayed Acceptance, chains sharded by sample() ------------------------------------------
    w2 = workloads.cfg2_da()
    C2, it2 = 1024, 6
    th2 = w2["prior"].rvs(C2, random_state=np.random.default_rng(2))
    res2, e2 = tda.sample(w2["posteriors"], w2["proposal"], it2, n_chains=C2, initial_parameters=[t for t in th2],
                          subchain_length=10, seed=4, store_model_output=False, store_coarse_chain=False,
                          dtype="float32", return_engine=True)
    kern = e2.kernel()
    lo2, hi2 = res2["local_chains"]
    mine2 = np.stack([res2["chain_fine_%d" % c].parameters for c in range(lo2, hi2)])
    spec2 = lower_problem(w2["posteriors"], w2["proposal"], 10)
    eng = Engine(spec2, C2, dtype="float32", seed=4, store=[STORE_NONE, STORE_STATS], capacity_iterations=it2, device=local)
    eng.init(th2)
    eng.run(it2)
    one2 = np.transpose(eng.fetch(1, "theta"), (2, 0, 1))
    assert eng.kernel() == kern == "tc16"
    assert np.array_equal(mine2, one2[lo2:hi2]), "sharded DA differs from the single-engine run"

    # ---- R-hat reduction over NCCL -----------------------------------------------------------------
    mom = eng.get("moments")
    full = parallel_free_rhat(mom, it2 + 1)
    red = parallel.allreduce_chain_moments(mom[0][:, lo2:hi2], mom[1][:, lo2:hi2], it2 + 1)
    eng.close()
    e2.close()
    np.testing.assert_allclose(red["rhat"], full, rtol=1e-9)
    assert red["n_chains"] == C2
    # ---- ESS reduction over NCCL: sharded == all chains in one process -----------------------------
    from tinyda_b200.diagnostics import _ess_plain
    sub = one2[:64, :, :4].astype(np.float64)                  # 64 chains, 4 parameters
    a, b = parallel.shard_range(64, rank, world)
    ess = parallel.allreduce_ess(sub[a:b])
    np.testing.assert_allclose(ess, [_ess_plain(sub[:, :, k]) for k in range(4)], rtol=1e-9)
    dist.barrier()
    dist.destroy_process_group()


def parallel_free_rhat(mom, n):
    mean_c = mom[0] / n
    var_c = (mom[1] - n * mean_c ** 2) / (n - 1.0)
    W = var_c.mean(axis=1)
    B_over_n = mean_c.var(axis=1, ddof=1)
    return np.sqrt(((n - 1.0) / n * W + B_over_n) / W)


if __name__ == "__main__":
    outdir = sys.argv[1]
    rank = int(os.environ.get("RANK", "0"))
    try:
        main(outdir)
        msg = "ok"
    except Exception:
        msg = traceback.format_exc()
    with open(os.path.join(outdir, "rank%d.txt" % rank), "w") as f:
        f.write(msg)
    sys.exit(0 if msg == "ok" else 1)
