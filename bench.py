#!/usr/bin/env python
"""Benchmark of the batched-MCMC hot path (BASELINE.json metric: fine-level MH transitions/sec,
all chains).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--dtype float32]

Workload (config.workload): BASELINE.json configs[1] -- two-level Delayed Acceptance, pCN,
linear-Gaussian inverse problem with 64 params / 1024 obs (coarse = 128-obs subset),
subsampling_rate = 10, 65536 chains PER GPU (weak scaling: chains are independent, no
data-path collective).  A "step" is one engine run of ITERS fine-level iterations for every
chain (= ITERS*J coarse proposals + ITERS fine evaluations per chain), one kernel launch.

  value     transitions/s with chain state resident in HBM (CUDA events around the run calls)
  e2e       the same through the host-buffer path: initial states H2D from pinned memory,
            init + run, fine-level history (theta, log-prior, log-like, accept) D2H to pinned
  roofline  the dominant kernel against the measured tensor peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference: the CPU port of the reference's chain loop (oracle/),
            one chain per process on all host cores, like the reference's Ray path does.
"""
import argparse
import json
import os

# one BLAS thread per process: the CPU baseline runs one chain per process on every core
# (must be set before numpy loads OpenBLAS)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CHAINS_PER_GPU = 65536
ITERS_PER_STEP = 50
F_ALG = 458752.0       # algorithmic flop per fine transition
# tensor-core flop the fp16-split kernel EXECUTES per fine transition: per coarse step
# theta @ aG_c^T (3 products) + z @ bTG_c^T (2) + z @ T (2), per fine step theta @ [G_f^T | LP] (3)
F_EXEC_TC16 = 2.0 * (10 * 64 * (3 * 128 + 2 * 128 + 2 * 64) + 3 * 64 * (1024 + 64))
METRIC = "fine-level MH transitions/sec, all chains"
UNIT = "transitions/s"


def workload_spec():
    from tinyda_b200 import lower_problem
    from tinyda_b200.workloads import cfg2_da
    w = cfg2_da()
    spec = lower_problem(w["posteriors"], w["proposal"], w["kwargs"]["subchain_length"])
    return w, spec


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's DAChain loop, one chain per process
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    spec, theta0, seed, iters, faithful = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import tinyda_oracle as orc
    rng = np.random.default_rng(seed)
    d, J = spec["d"], spec["J"][0]
    z = rng.standard_normal(iters * J * d + 8)
    u = rng.random(iters * (J + 1) + 8)
    ch = orc.ChainOracle(spec, theta0, z, u, svd_per_proposal=faithful, keep_history=False)
    t0 = time.perf_counter()
    ch.run(iters)
    return time.perf_counter() - t0


def cpu_port_rate(spec, prior, iters, faithful, procs=None):
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    rng = np.random.default_rng(123)
    theta0 = np.atleast_2d(prior.rvs(procs, random_state=rng)).reshape(procs, -1)
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [(spec, theta0[i], 1000 + i, iters, faithful) for i in range(procs)])
    wall = time.perf_counter() - t0
    return procs * iters / wall, procs, wall


def calibrated_iters(spec, prior, faithful, target_s, lo=4, hi=400000):
    """Fine iterations per process so that one cpu_port_rate() call takes about target_s
    (two probes: the first one is dominated by the fork / import overhead of the pool)."""
    n = 4 if faithful else 40
    for _ in range(2):
        rate, procs, wall = cpu_port_rate(spec, prior, n, faithful)
        if wall > 0.4 * target_s:
            break
        n = int(min(hi, max(lo, n * min(40.0, target_s / max(wall, 1e-3)) * 0.7)))
    rate, procs, wall = cpu_port_rate(spec, prior, n, faithful)
    per_proc = n / max(wall, 1e-3)
    return int(min(hi, max(lo, per_proc * target_s)))


def cpu_baseline_block(spec, prior, target_s=15.0, target_fast_s=6.0):
    n_f = calibrated_iters(spec, prior, True, target_s)
    rate, procs, wall = cpu_port_rate(spec, prior, n_f, True)
    n_p = calibrated_iters(spec, prior, False, target_fast_s)
    rate_fast, _, wall_fast = cpu_port_rate(spec, prior, n_p, False)
    return {
        "value": rate, "unit": UNIT, "cores": procs, "kind": "port",
        "sample": "oracle port of tinyDA DAChain.sample, %d processes x 1 chain x %d fine iterations "
                  "(= %d transitions), SVD of the 64x64 proposal covariance on every draw as "
                  "np.random.multivariate_normal does inside the reference; %.1f s wall"
                  % (procs, n_f, procs * n_f, wall),
        "port_precomputed_factor": {
            "value": rate_fast, "sample": "same port with the SVD factor computed once, %d x %d iterations, %.1f s"
                                          % (procs, n_p, wall_fast)},
    }


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
        return pk, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def sharded_self_check(rank, world, local_rank):
    """Untimed, N > 1 only: sharding a job over the ranks must change nothing.  (1) DREAM with the shared
    archive -- the path's one real exchange step: each step's new archive rows all-gathered over NCCL
    (ray.py:366-384) -- and (2) the Delayed-Acceptance job sharded by tda.sample() are compared, bit for
    bit, with the same job on ONE engine.  The verdict is AND-reduced over the ranks."""
    import contextlib
    import io
    import torch
    import torch.distributed as dist
    import tinyda_b200 as tda
    from tinyda_b200 import lower_problem, parallel, workloads
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    ok = {}
    try:
        w = workloads.cfg5_dream()
        C, iters, M0, d = 128 * world, 24, 16, 32
        rng = np.random.default_rng(5)
        theta0 = np.atleast_2d(w["prior"].rvs(C, random_state=rng))
        archive0 = w["prior"].rvs(C * M0, random_state=rng).reshape(C, M0, d)
        with contextlib.redirect_stdout(io.StringIO()):
            res = tda.sample(w["posteriors"][0], w["proposal"], iters, n_chains=C, initial_parameters=theta0, seed=9,
                             initial_archive=archive0, store_model_output=False, dtype="float32", device=local_rank)
        lo, hi = res.local_chains
        mine = res.history.dense("theta")
        eng = Engine(lower_problem(w["posteriors"], w["proposal"]), C, dtype="float32", seed=9, store=STORE_STATS,
                     capacity_iterations=iters, device=local_rank, archive0=archive0)
        eng.init(theta0)
        eng.run(iters)
        one = np.transpose(eng.fetch(0, "theta"), (2, 0, 1))
        eng.close()
        ok["dream_shared_allgather"] = bool(np.array_equal(mine, one[lo:hi]) and np.abs(np.diff(mine, axis=1)).max() > 0)

        w2 = workloads.cfg2_da()
        C2, it2 = 256 * world, 12
        th2 = w2["prior"].rvs(C2, random_state=np.random.default_rng(2))
        with contextlib.redirect_stdout(io.StringIO()):
            res2 = tda.sample(w2["posteriors"], w2["proposal"], it2, n_chains=C2, initial_parameters=th2, subchain_length=10,
                              seed=4, store_model_output=False, store_coarse_chain=False, dtype="float32", device=local_rank,
                              chunk_iterations=5)
        lo2, hi2 = res2.local_chains
        mine2 = res2.history.dense("theta")
        eng = Engine(lower_problem(w2["posteriors"], w2["proposal"], 10), C2, dtype="float32", seed=4,
                     store=[STORE_NONE, STORE_STATS], capacity_iterations=it2, device=local_rank)
        eng.init(th2)
        eng.run(it2)
        one2 = np.transpose(eng.fetch(1, "theta"), (2, 0, 1))
        eng.close()
        ok["da_sharded_by_sample"] = bool(np.array_equal(mine2, one2[lo2:hi2]))
    except Exception as exc:                     # a failed check is reported, it does not kill the benchmark line
        ok["error"] = repr(exc)[:300]
    flag = torch.tensor([1 if (ok.get("dream_shared_allgather") and ok.get("da_sharded_by_sample")) else 0],
                        dtype=torch.int32, device=torch.device("cuda", local_rank))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok["sharded_equals_single"] = bool(int(flag.item()))
    ok["ranks"] = world
    return ok


def strong_scaling_arm(args, spec, w, rank, world, local_rank, dtype, stream, flush, barrier):
    """north_star's own multi-GPU configuration: 65536 chains IN TOTAL (the reference's n_chains is a total,
    sampler.py:21-34), i.e. 65536 / N chains per GPU -- strong scaling.  Device-timed like `value`."""
    import torch
    import torch.distributed as dist
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE
    total = N_CHAINS_PER_GPU
    Cs = total // world
    iters = args.iters if args.iters is not None else ITERS_PER_STEP
    theta0 = w["prior"].rvs(Cs, random_state=np.random.default_rng(2000 + rank)).astype(np.float64)
    eng = Engine(spec, Cs, dtype=dtype, rng="philox", seed=2024, store=[STORE_NONE, STORE_STATS], capacity_iterations=iters,
                 device=local_rank, chain_offset=rank * Cs, n_chains_global=total, stream=stream)
    if args.kernel != "auto":
        eng.select_kernel(args.kernel)
    eng.init(theta0)
    eng.run(300 if not args.quick else 10, record=False)
    for _ in range(args.warmup):
        eng.history_reset()
        eng.run(iters)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        eng.history_reset()
        eng.run(iters)
        ev[k][1].record()
    barrier()
    ms = float(sum(a.elapsed_time(b) for a, b in ev))
    t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local_rank))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    kern = eng.kernel()
    eng.close()
    return {"value": float(total) * iters * args.steps / (ms * 1e-3), "unit": UNIT, "scaling": "strong", "chains_total": total,
            "chains_per_gpu": Cs, "ms_per_step": ms / args.steps, "kernel": kern,
            "tiles_per_gpu": (Cs + 127) // 128,
            "note": "%d chains per GPU = %d tiles of 128 chains on 148 SMs; below 74 tile pairs every CTA advances a single tile"
                    % (Cs, (Cs + 127) // 128)}


def run_ours(args):
    import torch
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_NONE, launch_count

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    w, spec = workload_spec()
    C = args.chains if args.chains is not None else N_CHAINS_PER_GPU
    iters = args.iters if args.iters is not None else ITERS_PER_STEP
    dtype = args.dtype
    d = spec["d"]
    rng = np.random.default_rng(1000 + rank)
    theta0 = w["prior"].rvs(C, random_state=rng).astype(np.float64)

    stream = torch.cuda.current_stream().cuda_stream
    # device-resident arm: stats-only history at the fine level (theta, prior, like, accept)
    coarse_hist = args.history == "coarse"
    J0 = int(spec["J"][0])
    eng = Engine(spec, C, dtype=dtype, rng="philox", seed=2024, store=[STORE_STATS if coarse_hist else STORE_NONE, STORE_STATS],
                 capacity_iterations=iters, device=local_rank, chain_offset=rank * C,
                 n_chains_global=world * C, stream=stream)
    if args.kernel != "auto":
        eng.select_kernel(args.kernel)
    eng.init(theta0)
    eng.run(300 if not args.quick else 10, record=False)      # burn-in from the prior draws (nothing recorded)
    eng.sync()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > L2 (126 MB)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        eng.history_reset()
        eng.run(iters)

    for _ in range(args.warmup):
        one_step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    acc_t0 = eng.get("accept_counts").astype(np.float64)
    l0 = launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed steps (untimed)
        ev[k][0].record()
        one_step()
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = launch_count() - l0
    acc_t1 = eng.get("accept_counts").astype(np.float64)
    accept_timed = {"coarse": float((acc_t1[0] - acc_t0[0]).mean() / (iters * args.steps * J0)),
                    "fine": float((acc_t1[1] - acc_t0[1]).mean() / (iters * args.steps))}
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(ms_steps))

    # ---- end-to-end arm: the user's call, host buffers in, host buffers out -----------------
    # tda.sample(...) itself (the drop-in for tinyDA.sample): initial parameters in pinned host memory ->
    # H2D -> engine construction + initial Links + the run in blocks; after every block the finest level's
    # records are compacted on the device to the accepted ones (a rejected step repeats the previous Link,
    # chain.py:116, :434) and copied to pinned host memory while the next block runs.  The chains start from
    # the burnt-in states the device arm left (so the accept rates are the stationary ones).
    import contextlib
    import io
    import tinyda_b200 as tda
    from tinyda_b200.engine import pinned_empty
    # with every coarse Link recorded (10 x 264 B per fine iteration and chain, fetched densely) the host holds the
    # whole coarse chain: keep that call short
    e2e_iters = args.e2e_iters if args.e2e_iters is not None else ((1000 if not coarse_hist else 20) if not args.quick else 40)
    cur = eng.get("theta", 1)                                            # [C, d] float64, burnt-in states
    theta_host = pinned_empty((world * C, d), np.float64)
    theta_host[:] = 0.0
    theta_host[rank * C:(rank + 1) * C] = cur
    store_coarse = coarse_hist
    d2h_seen = [0]

    def e2e_step(k):
        ta = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            res = tda.sample(w["posteriors"], w["proposal"], e2e_iters, n_chains=world * C, initial_parameters=theta_host,
                             subchain_length=J0, dtype=dtype, seed=5000 + k, store_model_output=False,
                             store_coarse_chain=store_coarse, device=local_rank)
        h = res.history
        # the device->host read of the step's result: every block's accept flags, row offsets and accepted rows
        d2h_seen[0] = sum(ch.accept.nbytes + ch.offsets.nbytes + ch.theta.nbytes + ch.prior.nbytes + ch.like.nbytes for ch in h.chunks)
        assert h.n_records == e2e_iters + 1
        out_val = float(h.chunks[-1].like[-1])
        tb = time.perf_counter()
        del res, h
        if os.environ.get("TDA_PROFILE"):
            sys.stderr.write("[bench e2e rank %d] sample() %.1f ms, release %.1f ms\n" % (rank, (tb - ta) * 1e3, (time.perf_counter() - tb) * 1e3))
        return out_val

    e2e_steps = max(2, min(args.steps, 3))
    for k in range(3):                      # warm-up: device / pinned pools, allocator, lazy imports
        e2e_step(-1 - k)
    barrier()
    l_e2e0 = launch_count()
    t0 = time.perf_counter()
    e2e_call_ms = []
    for k in range(e2e_steps):
        tk = time.perf_counter()
        e2e_step(k)
        e2e_call_ms.append((time.perf_counter() - tk) * 1e3)
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_launches = launch_count() - l_e2e0

    # engine-level variant of the same thing (no engine construction in the timed region): init from the
    # pinned states + run in blocks + compaction + D2H, through the C ABI calls sample() makes
    from tinyda_b200.link import CompactHistory
    n_chunks = max(1, min(5, iters // 5))
    bounds = [iters * k // n_chunks for k in range(n_chunks + 1)]
    pin_cur = pinned_empty((C, d), np.float64)
    pin_cur[:] = cur

    def engine_e2e_step():
        hist = CompactHistory(C)
        eng.init(pin_cur)
        pending = None
        for k in range(n_chunks):
            if k:
                eng.history_reset()
            n = bounds[k + 1] - bounds[k]
            eng.run(n)
            eng.compact_begin(0, n + (k == 0), k == 0, ("theta", "stats"), slot=k & 1)
            if pending is not None:
                hist.append(eng.compact_collect(pending))
            pending = k & 1
        hist.append(eng.compact_collect(pending))
        eng.compact_sync()
        return hist

    for _ in range(3):
        hist_e = engine_e2e_step()
    barrier()
    t0 = time.perf_counter()
    eng_e2e_steps = max(2, min(args.steps, 5))
    for _ in range(eng_e2e_steps):
        hist_e = engine_e2e_step()
    barrier()
    eng_e2e_wall = time.perf_counter() - t0
    eng_d2h = sum(ch.accept.nbytes + ch.offsets.nbytes + ch.theta.nbytes + ch.prior.nbytes + ch.like.nbytes for ch in hist_e.chunks)
    clock_info = clocks.stop()

    times = torch.tensor([dev_ms, e2e_wall * 1e3, t_wall * 1e3, eng_e2e_wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms, eng_e2e_ms = [float(x) for x in times.cpu()]
    total_trans = float(world) * C * iters * args.steps
    value = total_trans / (dev_ms * 1e-3)
    e2e_value = float(world) * C * e2e_iters * e2e_steps / (e2e_ms * 1e-3)
    eng_e2e_value = float(world) * C * iters * eng_e2e_steps / (eng_e2e_ms * 1e-3)
    h2d = C * d * 8
    d2h = d2h_seen[0]
    checks = sharded_self_check(rank, world, local_rank) if world > 1 else None
    strong = strong_scaling_arm(args, spec, w, rank, world, local_rank, dtype, stream, flush, barrier) if world > 1 else None

    kernel_used = args.kernel if args.kernel != "auto" else ("tc16" if dtype == "float32" else "generic")
    ess = None
    if not args.no_ess:
        # second half of BASELINE.json's metric: min over parameters of the bulk ESS per second -- MEASURED on the
        # benchmarked configuration, on ALL chains: a recorded run of `keep` fine iterations from the burnt-in states
        # with the parameter history kept in HBM (keep x chains x 64 x 4 bytes), rank-normalised split bulk ESS
        # (Vehtari et al. 2021, what ArviZ computes) on the device (tda_ess_sums: radix sort, normal scores, all-lag
        # autocovariance sums), all-reduced over the ranks; ESS / (device time of that run).
        from tinyda_b200.diagnostics import ess_rhat_from_sums
        from tinyda_b200 import parallel
        keep = 600 if not args.quick else 40
        e2 = Engine(spec, C, dtype=dtype, rng="philox", seed=4048, store=[STORE_NONE, STORE_STATS], capacity_iterations=keep,
                    device=local_rank, chain_offset=rank * C, n_chains_global=world * C, stream=stream)
        if args.kernel != "auto":
            e2.select_kernel(args.kernel)
        # chains at stationarity: draws of the closed-form (conjugate) posterior of the fine level, so that the
        # estimate is the sampler's autocorrelation and not a leftover of the burn-in (R-hat is reported)
        from tinyda_b200.workloads import conjugate_posterior
        mu_post, S_post = conjugate_posterior(w["G"], w["y"], w["sigma2"], w["prior"])
        th_stat = np.random.default_rng(77 + rank).multivariate_normal(mu_post, S_post, size=C)
        e2.init(th_stat)
        e2.run(50 if not args.quick else 5, record=False)        # settle the coarse/fine Link pairs
        e2.history_reset()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        e2.run(keep)
        ev1.record()
        barrier()
        run_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(run_ms, op=dist.ReduceOp.MAX)
        t_e0 = time.perf_counter()
        sums, folded = e2.ess_sums(1, 0, keep)
        ess_wall = time.perf_counter() - t_e0
        e2.close()
        sums, folded = parallel.allreduce_ess_sums(sums, folded)
        per_param, rhat_pp = ess_rhat_from_sums(sums, folded)
        run_s = float(run_ms.item()) * 1e-3
        ess = {"min_ess": float(per_param.min()), "median_ess": float(np.median(per_param)), "max_rhat": float(rhat_pp.max()),
               "min_ess_per_transition": float(per_param.min() / (world * C * keep)),
               "run_seconds": run_s, "diagnostics_seconds_on_device": ess_wall,
               "sample": "%d chains x %d fine iterations started from draws of the closed-form posterior, all chains, "
                         "rank-normalised split bulk ESS per parameter computed on the device" % (world * C, keep)}
    if rank == 0:
        pk, src = measured_peaks()
        per_gpu_rate = C * iters / (np.mean(ms_steps) * 1e-3)
        achieved = per_gpu_rate * F_ALG / 1e12
        # every timed step is one isolated launch of a few milliseconds: the burst figure applies
        peak = float(pk.get("bf16_tflops", pk.get("bf16_tflops_sustained")))
        peak_sustained = float(pk.get("bf16_tflops_sustained", peak))
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == "float32" else "f64",
            "data": "synthetic",
            "config": {
                "workload": w["name"], "chains_per_gpu": C, "fine_iterations_per_step": iters,
                "transitions_per_step": world * C * iters, "rng": "philox4x32-10 in-kernel",
                "history": ("fine level theta+log-prior+log-like+accept (265 B/transition f32)" if not coarse_hist else
                            "fine level theta+log-prior+log-like+accept and every coarse Link theta+log-like+accept "
                            "(265 + %d x 261 = %d B/transition f32); coarse log-prior rebuilt from theta at fetch time"
                            % (J0, 265 + J0 * 261)),
                "l2": "256 MiB buffer written between timed steps (L2 flush); chain state is kept "
                      "L2/SMEM-resident by design", "kernel": eng_kernel_name(args, dtype),
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "fine_iterations_per_step": e2e_iters, "gpu_launches": int(e2e_launches),
                    "ms_per_call": [round(x, 2) for x in e2e_call_ms],
                    "what": "tinyda_b200.sample(posteriors, pCN, %d iterations, n_chains=%d, initial_parameters=<pinned host array>, "
                            "dtype=float32, store_model_output=False, store_coarse_chain=%s): engine construction, initial states "
                            "H2D, initial Links, the run in blocks, device-side compaction of each block to its accepted "
                            "records, D2H to pinned host memory overlapped with the next block, lazy result dict"
                            % (e2e_iters, world * C, store_coarse),
                    "d2h_bytes_per_transition": d2h / float(C * e2e_iters),
                    "engine_level": {"value": eng_e2e_value, "unit": UNIT, "steps": eng_e2e_steps, "fine_iterations_per_step": iters,
                                     "d2h_bytes_per_step": int(eng_d2h),
                                     "what": "the same C-ABI calls on a live engine (no construction in the timed region): "
                                             "pinned states H2D + tda_engine_init + %d x (tda_engine_run, tda_compact_begin, "
                                             "tda_compact_fetch)" % n_chunks}},
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "frac_of_sustained_peak": achieved / peak_sustained,
                # dram__bytes_read.sum + dram__bytes_write.sum of one da_tc16_kernel launch of this workload
                # (ncu --set full, profiles/r02_ncu_tc16_summary.txt): 0.164 + 1.132 GB against 0.868 GB of
                # algorithmic history bytes (265 B x 3,276,800 transitions); the rest is the chain-state hand-off
                # between work units and the running moments (red.global.add)
                "traffic": 1.296e9 if (kernel_used == "tc16" and C == N_CHAINS_PER_GPU and iters == ITERS_PER_STEP and not coarse_hist) else None,
                "executed_tflops": per_gpu_rate * F_EXEC_TC16 / 1e12 if kernel_used == "tc16" else None,
                "executed_frac": per_gpu_rate * F_EXEC_TC16 / 1e12 / peak if kernel_used == "tc16" else None,
                "note": "achieved = 458752 algorithmic flop/transition x per-GPU transitions/s "
                        "(CUDA events, mean over timed launches); peak = bf16 dense burst (each timed step is one isolated launch), " + src +
                        " (MEASURED_PEAKS.json). fp32-grade accuracy on 16-bit tensor cores needs a two-term "
                        "fp16 split (3 products per contraction), so the algorithmic figure caps near peak/3; "
                        "executed_* counts the fp16 tensor flop the kernel really issues (1.40 MFLOP/transition)",
            },
            "accept_rate_timed_region": accept_timed,
            "wall_ms_timed_region": wall_ms,
        }
        if checks is not None:
            out["checks"] = checks
        if strong is not None:
            out["strong"] = strong
        if coarse_hist:
            hb = 265 + J0 * 261
            out["link_writeout"] = {"bytes_per_transition": hb, "achieved_gbs": per_gpu_rate * hb / 1e9,
                                    "hbm_peak_gbs": float(pk["hbm_gbs"]), "frac": per_gpu_rate * hb / 1e9 / float(pk["hbm_gbs"])}
        if ess is not None:
            ess["min_ess_per_s"] = ess["min_ess"] / ess["run_seconds"]
            out["min_ess_per_s"] = ess["min_ess_per_s"]
            out["ess"] = ess
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline_block(spec, w["prior"], 1.0 if args.quick else 15.0,
                                                     0.5 if args.quick else 6.0)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# Secondary workloads (BASELINE.json configs[2..4]); not the headline line, same timing rules
# ------------------------------------------------------------------------------------------
def _other_workload(name, world, dream_sync=1):
    from tinyda_b200 import lower_problem, workloads
    w = workloads.WORKLOADS[name]()
    kw = w["kwargs"]
    if name == "cfg5" and dream_sync > 1:
        w["proposal"].sync_every = int(dream_sync)         # bounded staleness of the shared archive (extension)
        w["name"] += ", sync_every=%d" % dream_sync
    spec = lower_problem(w["posteriors"], w["proposal"], kw.get("subchain_length"), kw.get("adaptive_error_model"))
    d = spec["d"]
    s = 4
    if name == "cfg1":     # README linear regression: theta, prior, loglike + accept byte (stats-only history)
        cfgd = dict(chains=2, iters=12000, bytes_unit=(d + 2) * s + 1, flop_unit=2.0 * d * 100 + 4.0 * d * d,
                    bound="latency", store="stats")
    elif name == "cfg2rw":  # cfg2's shape with a random-walk proposal: same algorithmic flop + the coarse log-priors
        cfgd = dict(chains=65536, iters=50, bytes_unit=(d + 2) * s + 1, flop_unit=F_ALG + 2.0 * 10 * 64 * 64, bound="tensor",
                    store="stats")
    elif name == "cfg3":   # theta, prior, loglike, F + accept byte
        cfgd = dict(chains=1 << 20, iters=100, bytes_unit=(d + 3) * s + 1, flop_unit=120.0, bound="hbm",
                    store="full")
    elif name == "cfg4":
        ns = [lv["model"]["n_grid"] for lv in spec["levels"]]
        J = spec["J"]
        evals = [J[0] * J[1] * J[2], J[1] * J[2], J[2], 1]
        flop = sum(e * (2 * d + 9) * n for e, n in zip(evals, ns))
        cfgd = dict(chains=32768, iters=4, bytes_unit=(d + 2) * s + 1, flop_unit=float(flop), bound="fp32", store="stats")
    elif name == "cfg5":
        m = spec["levels"][0]["model"]["m"]
        # 500 lock-step steps per launch: the ranks' persistent kernels wait for each other from their first step on, so
        # the host-side launch skew between the processes (hundreds of microseconds) is part of every timed launch
        cfgd = dict(chains=8192 // world, iters=500, bytes_unit=(d + 2) * s + 1 + 3 * d * s, flop_unit=2.0 * d * m,
                    bound="latency", store="stats")
    else:
        raise SystemExit("unknown workload " + name)
    return w, spec, cfgd


def run_other(args):
    """configs[2..4] through the generic lock-step kernel: device-timed value, e2e with the
    finest-level history copied to pinned host memory, and the SURVEY 8(d) roofline figure."""
    import torch
    from tinyda_b200.engine import Engine, STORE_STATS, STORE_FULL, STORE_NONE, launch_count
    from tinyda_b200 import parallel
    from tinyda_b200.proposal import PROP_DREAM

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w, spec, cd = _other_workload(args.workload, world, args.dream_sync)
    C = args.chains if args.chains is not None else cd["chains"]
    iters = args.iters if args.iters is not None else cd["iters"]
    L, d = spec["n_levels"], spec["d"]
    shared = int(spec["proposal"]["kind"]) == PROP_DREAM
    rng = np.random.default_rng(1000 + rank)
    theta0 = np.atleast_2d(w["prior"].rvs(C, random_state=rng)).reshape(C, -1).astype(np.float64)
    archive0 = None
    total_runs = args.warmup + args.steps + 2 + max(2, min(args.steps, 5))
    if shared:
        M0 = int(spec["proposal"]["M0"])
        arng = np.random.default_rng(77)
        archive0 = np.atleast_2d(w["prior"].rvs(world * C * M0, random_state=arng)).reshape(world * C, M0, d)
    store = [STORE_NONE] * (L - 1) + [STORE_FULL if cd["store"] == "full" else STORE_STATS]
    stream = torch.cuda.current_stream().cuda_stream
    eng = Engine(spec, C, dtype=args.dtype, rng="philox", seed=2024, store=store,
                 capacity_iterations=iters, archive_iterations=(iters * total_runs + 64) if shared else None, device=local_rank,
                 chain_offset=rank * C, n_chains_global=world * C if shared else C, archive0=archive0, stream=stream)
    eng.init(theta0)
    exchange = None
    if shared:
        exchange = "single device: persistent launch, grid barrier per step"
        if world > 1:
            ok = parallel.connect_dream_peers(eng, rank, world)
            exchange = ("persistent launch; each step's rows stored into every replica over NVLink peer memory, arrival counters (one system fence per CTA, remote reductions)"
                        if ok else "one launch + one NCCL all-gather per step")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        eng.history_reset()
        if shared and world > 1:
            parallel.run_dream_shared(eng, iters, rank, world)
        else:
            eng.run(iters)

    for _ in range(args.warmup):
        one_step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    acc0 = eng.get("accept_counts").astype(np.float64)
    l0 = launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        one_step()
        ev[k][1].record()
    barrier()
    launches = launch_count() - l0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(ms_steps))
    acc1 = eng.get("accept_counts").astype(np.float64)
    accept_rate = [float((b - a).mean() / (iters * args.steps * st)) for a, b, st in zip(acc0, acc1, eng.steps)]

    # e2e: initial states from pinned host memory, run, finest-level history to pinned host memory
    tdt = torch.float32 if args.dtype == "float32" else torch.float64
    top = L - 1
    m_top = int(spec["levels"][top]["model"]["m"])
    pin_theta0 = torch.from_numpy(theta0).pin_memory()
    h_theta = torch.empty((iters, d, C), dtype=tdt).pin_memory()
    h_prior = torch.empty((iters, C), dtype=tdt).pin_memory()
    h_like = torch.empty((iters, C), dtype=tdt).pin_memory()
    h_acc = torch.empty((iters, C), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((iters, m_top, C), dtype=tdt).pin_memory() if cd["store"] == "full" else None

    def e2e_step():
        eng.history_reset()
        if not shared:                                   # DREAM keeps its archive: states continue
            eng.init(pin_theta0.numpy())
        one_rec0 = int(eng.n_records()[top])
        if shared and world > 1:
            parallel.run_dream_shared(eng, iters, rank, world)
        else:
            eng.run(iters)
        eng.fetch(top, "theta", one_rec0, iters, out=h_theta.numpy(), sync=False)
        eng.fetch(top, "prior", one_rec0, iters, out=h_prior.numpy(), sync=False)
        eng.fetch(top, "like", one_rec0, iters, out=h_like.numpy(), sync=False)
        eng.fetch(top, "accept", one_rec0, iters, out=h_acc.numpy(), sync=False)
        if h_out is not None:
            eng.fetch(top, "output", one_rec0, iters, out=h_out.numpy(), sync=False)
        eng.sync()

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t0
    clock_info = clocks.stop()
    times = torch.tensor([dev_ms, e2e_wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = [float(x) for x in times.cpu()]
    if rank == 0:
        pk, src = measured_peaks()
        value = float(world) * C * iters * args.steps / (dev_ms * 1e-3)
        per_gpu = C * iters / (np.mean(ms_steps) * 1e-3)
        d2h = sum(t.numel() * t.element_size() for t in (h_theta, h_prior, h_like, h_acc) + ((h_out,) if h_out is not None else ()))
        if cd["bound"] == "tensor":
            roof = {"bound": "tensor", "achieved": per_gpu * cd["flop_unit"] / 1e12, "peak": float(pk["bf16_tflops"]), "unit": "TFLOP/s",
                    "note": "%.0f algorithmic flop/transition x per-GPU transitions/s; peak = bf16 dense burst, %s; the kernel runs 3xTF32 "
                            "(three tf32 products per contraction) for fp32-grade accuracy" % (cd["flop_unit"], src)}
        elif cd["bound"] == "hbm":
            roof = {"bound": "hbm", "achieved": per_gpu * cd["bytes_unit"] / 1e9, "peak": float(pk["hbm_gbs"]), "unit": "GB/s",
                    "note": "%d B/transition (SURVEY 8d) x per-GPU transitions/s; peak = copy bandwidth, %s" % (cd["bytes_unit"], src)}
        else:
            # measured with tools/ubench/fma_peak.cu on this pool's B200 (profiles/r02_measured_fma_peaks.json)
            fpath = os.path.join(ROOT, "profiles", "r02_measured_fma_peaks.json")
            fpk = json.load(open(fpath)) if os.path.exists(fpath) else {}
            if args.dtype == "float64":
                peak, what = float(fpk.get("fp64_dfma_tflops", 37.0)), "fp64 DFMA"
            else:
                peak, what = float(fpk.get("fp32_ffma_tflops", 71.0)), "fp32 FFMA"
            roof = {"bound": cd["bound"], "achieved": per_gpu * cd["flop_unit"] / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "note": "%.0f algorithmic flop/unit (SURVEY 8d) x per-GPU units/s; peak = %s measured with tools/ubench/fma_peak.cu "
                            "(profiles/r02_measured_fma_peaks.json)" % (cd["flop_unit"], what)}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["traffic"] = None
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if shared else "weak",
            "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
            "config": {"workload": w["name"], "chains_per_gpu": C, "finest_iterations_per_step": iters,
                       "kernel": "%s (%s)" % (eng.kernel(), args.dtype), "l2": "256 MiB buffer written between timed steps"},
            "e2e": {"value": float(world) * C * iters * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(theta0.nbytes) if not shared else 0, "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roof,
            "accept_rate_timed_region": accept_rate,
            **({"archive_exchange": exchange, "us_per_lockstep_step": dev_ms / args.steps / iters * 1e3} if shared else {}),
        }))
    if world > 1:
        dist.destroy_process_group()


def eng_kernel_name(args, dtype):
    k = args.kernel if args.kernel != "auto" else ("auto -> tc16" if dtype == "float32" else "auto -> generic")
    return "%s (%s)" % (k, dtype)


def run_reference(args):
    """The reference's own CPU implementation of the path: tinyDA is pure Python and does not
    travel to the GPU box, so this times the oracle port of its DAChain loop (one chain per
    process on all host cores, as tinyDA/ray.py does)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, spec = workload_spec()
    # every step is a bounded sample: the whole --steps K --warmup W run is sized for ~150 s
    per_step_s = 1.0 if args.quick else max(3.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    iters = calibrated_iters(spec, w["prior"], True, per_step_s)
    steps_done = []
    t_all0 = time.perf_counter()
    for _ in range(args.warmup):
        cpu_port_rate(spec, w["prior"], max(2, iters // 10), True)
    for _ in range(args.steps):
        rate, procs, wall = cpu_port_rate(spec, w["prior"], iters, True)
        steps_done.append((rate, procs, wall))
        if time.perf_counter() - t_all0 > 240:
            break
    n = len(steps_done)
    procs = steps_done[0][1]
    total_wall = sum(s[2] for s in steps_done)
    value = procs * iters * n / total_wall
    sample = ("oracle port of tinyDA DAChain.sample (per-draw SVD like np.random.multivariate_normal), "
              "%d processes x 1 chain x %d fine iterations per step" % (procs, iters))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": n, "warmup": args.warmup, "ms_per_step": total_wall / n * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "chains": procs, "fine_iterations_per_step": iters},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "tc", "tc16", "tcr"])
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (default: the workload's)")
    ap.add_argument("--iters", type=int, default=None, help="finest-level iterations per step (default: the workload's)")
    ap.add_argument("--history", default="fine", choices=["fine", "coarse"],
                    help="cfg2: 'fine' = fine-level Links only (the headline line); 'coarse' = the reference's "
                         "store_coarse_chain=True: every coarse Link (theta, log-like, accept) is recorded too")
    ap.add_argument("--e2e-iters", type=int, default=None, help="fine iterations per tda.sample() call of the e2e arm (default 1000)")
    ap.add_argument("--dream-sync", type=int, default=1, help="cfg5: DREAM(sync_every=K), the chains' view of the shared archive "
                    "is refreshed every K steps (default 1 = the lock-step rule)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ess", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg2rw", "cfg3", "cfg4", "cfg5"],
                    help="cfg2 = the headline line (BASELINE.json configs[1]); the others are secondary measurements")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not args.quick:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "cfg2":
        run_other(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
