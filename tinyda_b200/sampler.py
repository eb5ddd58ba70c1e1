"""``sample()``: the drop-in for ``tinyDA.sample`` (tinyDA/sampler.py:21-292).

Same signature, validation, warnings and result dict.  Where the reference builds one Python
chain object per chain and runs them one after another or as Ray actors
(sampler.py:295-508, ray.py:12-210), this lowers the problem once and advances ALL chains in
lock-step on the GPU (``engine.Engine`` -> C ABI -> CUDA kernels).  Extra keyword-only
arguments select engine options (dtype, RNG mode, storage); the positional / reference
keywords keep their meaning.  Under ``torch.distributed`` (one process per GPU) ``n_chains``
is the global count and every rank samples its contiguous shard.
"""
import os
import sys
import time
import warnings

import numpy as np
import scipy.stats as stats

from .posterior import Posterior
from .proposal import CrankNicolson, DREAMZ, DREAM, PROP_DREAM, PROP_DREAMZ, PROP_AM
from .lowering import lower_problem
from .link import LinkSequence, CompactHistory, SampleResult
from . import parallel


def _dense_family(parts):
    """Coarser levels are fetched densely, block by block: parts = [(theta, prior, like, acc, out, qoi), ...]
    in the device layout ([nrec, width, C] / [nrec, C]).  Returns the per-chain LinkSequence factory."""
    cat = lambda xs: xs[0] if len(xs) == 1 else np.concatenate(xs, axis=0)
    theta, prior, like, acc = (cat([p[k] for p in parts]) for k in range(4))
    out = None if parts[0][4] is None else cat([p[4] for p in parts])
    qoi = None if parts[0][5] is None else cat([p[5] for p in parts])
    acc = acc.astype(bool)
    return lambda c: LinkSequence(theta[:, :, c], prior[:, c], like[:, c], None if out is None else out[:, :, c], acc[:, c],
                                  None if qoi is None else qoi[:, :, c])


def _fresh_seed():
    """seed=None: a fresh 63-bit seed from the legacy global generator -- un-seeded calls differ from one
    another like the reference's do, and ``np.random.seed`` makes a run repeatable like it does there.
    Under torch.distributed rank 0's draw is broadcast (one job, one seed)."""
    seed = int(np.random.randint(0, np.iinfo(np.int64).max, dtype=np.int64))
    return parallel.broadcast_int(seed)


def _auto_chunk(iterations, n_local, spec, store, steps, itemsize, budget_bytes):
    """Fine iterations per block so that the dense device history of a block stays within the budget."""
    per_iter = 0
    for l, lv in enumerate(spec["levels"]):
        if not store[l]:
            continue
        w = spec["d"] + 2 + (int(lv["model"]["m"]) if store[l] & 4 else 0)
        per_iter += steps[l] * (w * itemsize + 1)
    pad = (n_local + 255) // 256 * 256
    return int(max(1, min(iterations, budget_bytes // max(1, per_iter * pad))))


def sample(
    posteriors,
    proposal,
    iterations,
    n_chains=1,
    initial_parameters=None,
    subchain_length=1,
    randomize_subchain_length=False,
    adaptive_error_model=None,
    store_coarse_chain=True,
    force_sequential=False,
    force_progress_bar=False,
    subsampling_rate=None,
    *,
    dtype="float64",
    rng="philox",
    seed=None,
    streams=None,
    store_model_output=True,
    initial_archive=None,
    device=None,
    return_engine=False,
    chunk_iterations=None,
    history_budget_bytes=2 << 30,
):
    """Returns MCMC samples given a Posterior (or a list of them, coarsest first) and a proposal,
    exactly like ``tinyDA.sample``: one posterior -> Metropolis-Hastings, two -> Delayed
    Acceptance, more -> MLDA.  ``force_sequential`` / ``force_progress_bar`` are accepted for
    compatibility and ignored (there is nothing sequential to force).

    Engine options (keyword-only): ``dtype`` 'float64' | 'float32'; ``rng`` 'philox' (per-chain
    counter-based streams generated in-kernel, seeded by ``seed``) or 'injected' with
    ``streams=(normals[n_chains, nz], uniforms[n_chains, nu])``; ``store_model_output`` keeps
    F(theta) of every stored link (Link.model_output) -- switch it off for large runs.
    ``seed=None`` draws a fresh seed from ``np.random`` (so ``np.random.seed`` makes the call repeatable,
    as it does for the reference).

    The run is cut into blocks of ``chunk_iterations`` fine iterations (default: as many as keep a block's
    device history within ``history_budget_bytes``).  After each block the finest level's records are
    compacted on the device to the accepted ones -- a rejected step repeats the previous Link
    (chain.py:116, :434) -- and copied to pinned host memory while the next block runs; the result dict
    expands a chain to a ``LinkSequence`` when its key is first read.
    """
    _prof = [] if os.environ.get("TDA_PROFILE") else None
    _t0 = time.perf_counter()

    def _lap(name):
        if _prof is not None:
            _prof.append((name, time.perf_counter()))

    if subsampling_rate is not None:                                   # sampler.py:113-115
        warnings.warn(" subsampling_rate has been deprecated in favour of subchain_length.")
        subchain_length = subsampling_rate

    if not isinstance(posteriors, list):
        posteriors = [posteriors]
    n_levels = len(posteriors)
    for p in posteriors:
        if not isinstance(p, Posterior):
            raise TypeError("posteriors must be tinyda_b200.Posterior instances")

    # sampler.py:138-143
    if isinstance(proposal, CrankNicolson) and not isinstance(
        posteriors[0].prior, stats._multivariate.multivariate_normal_frozen
    ):
        raise TypeError("Prior must be of type scipy.stats.multivariate_normal for pCN proposal")
    elif hasattr(proposal, "kernel"):                                  # sampler.py:146-152
        if isinstance(proposal.kernel, CrankNicolson) and not isinstance(
            posteriors[0].prior, stats._multivariate.multivariate_normal_frozen
        ):
            raise TypeError("Prior must be of type scipy.stats.multivariate_normal for pCN kernel")

    # sampler.py:184-193
    if n_levels > 2 and adaptive_error_model == "state-dependent":
        warnings.warn(
            " A state-dedependent adaptive error model for MLDA has not been implemented yet, "
            "defaulting to state-independent AEM..."
        )
        adaptive_error_model = "state-independent"
    if adaptive_error_model == "state-dependent" and subchain_length > 1:
        warnings.warn(
            " Using a state-dependent error model for subchain lengths larger than 1 is not "
            "guaranteed to be ergodic. \n"
        )
    if adaptive_error_model not in (None, "state-independent", "state-dependent"):
        raise ValueError("Adaptive error model can only be state-dependent, state-independent or None.")
    if n_levels == 2 and randomize_subchain_length:                   # chain.py:310-314
        if subchain_length == 1:
            raise ValueError("Randomize subchain length requires a subchain_length > 1.")
        if not store_coarse_chain:
            raise ValueError("Randomize subchain length requires storing the coarse chain.")

    if seed is None:
        seed = _fresh_seed()
    # sharding: one process per GPU, contiguous chain ranges (ray.py:68-74 -> chain sharding)
    rank, world = parallel.rank_world()
    lo, hi = parallel.shard_range(n_chains, rank, world)
    n_local = hi - lo

    # initial parameters, sampler.py:196-209
    d_prior = np.atleast_1d(posteriors[0].prior.rvs()).size
    if initial_parameters is not None:
        if type(initial_parameters) == list:
            assert (
                len(initial_parameters) == n_chains
            ), "If list of initial parameters is provided, it must have length n_chains"
        elif type(initial_parameters) == np.ndarray and initial_parameters.ndim == 2 and initial_parameters.shape == (n_chains, d_prior):
            pass                                     # extension: one row per chain (no Python list of 65536 arrays)
        elif type(initial_parameters) == np.ndarray:
            assert (
                d_prior == initial_parameters.size
            ), "If an array of initial parameters is provided, it must have the same dimension as the prior"
            initial_parameters = [initial_parameters] * n_chains
        else:
            raise TypeError("Initial paramaters must be list, numpy array or None")
        if type(initial_parameters) == np.ndarray:
            theta0 = np.ascontiguousarray(initial_parameters[lo:hi], dtype=np.float64)
        else:
            theta0 = np.array([np.atleast_1d(t) for t in initial_parameters[lo:hi]], dtype=np.float64)
    else:
        host_rng = np.random.default_rng([int(seed), 7])
        theta0 = np.atleast_2d(posteriors[0].prior.rvs(n_chains, random_state=host_rng)).reshape(n_chains, -1)[lo:hi]

    _lap("arguments + initial parameters")
    spec = lower_problem(posteriors, proposal, subchain_length if n_levels > 1 else None,
                         adaptive_error_model if n_levels > 1 else None,
                         randomize_subchain_length if n_levels == 2 else False)
    kind = int(spec["proposal"]["kind"])
    _lap("lowering")

    archive0 = None
    if kind in (PROP_DREAMZ, PROP_DREAM):
        M0 = int(spec["proposal"]["M0"])
        if initial_archive is not None:
            archive0 = np.asarray(initial_archive, dtype=np.float64)
        else:                                                           # proposal.py:788, per chain
            # every rank builds the same rows from the (broadcast) seed: one shared archive, ray.py:366-384
            host_rng = np.random.default_rng([int(seed), 11])
            base = proposal.kernel if hasattr(proposal, "kernel") else proposal
            archive0 = np.stack([base.initial_archive(posteriors[0].prior, host_rng) for _ in range(n_chains)])
        if kind == PROP_DREAMZ:
            archive0 = archive0[lo:hi]

    from .engine import Engine, STORE_FULL, STORE_STATS, STORE_NONE, steps_per_iteration   # loads the CUDA library
    full = STORE_FULL if store_model_output else STORE_STATS
    store = [full] * n_levels
    if not store_coarse_chain:
        for l in range(n_levels - 1):
            store[l] = STORE_NONE
    if streams is not None:
        streams = (np.asarray(streams[0])[lo:hi], np.asarray(streams[1])[lo:hi])
    if device is None:
        device = parallel.local_device()
    shared = kind == PROP_DREAM
    steps = steps_per_iteration(spec)
    itemsize = np.dtype(dtype).itemsize
    if chunk_iterations is None:
        chunk_iterations = _auto_chunk(iterations, n_local, spec, store, steps, itemsize, int(history_budget_bytes))
    chunk_iterations = int(max(1, min(max(iterations, 1), chunk_iterations)))
    from ._lib import EngineError
    try:
        eng = Engine(spec, n_local, dtype=dtype, rng=rng, seed=seed, store=store,
                     capacity_iterations=chunk_iterations, archive_iterations=iterations, streams=streams, device=device,
                     chain_offset=lo, n_chains_global=n_chains if shared else n_local,
                     archive0=archive0, am_device_refactor=True)
    except EngineError as exc:
        if "cudaMalloc" in str(exc):
            # the reference keeps every Link (with its model output) of every level in host lists;
            # here a block of that history lives in HBM until it is fetched
            raise EngineError(str(exc) + " -- a block of %d iterations of the Link history of %d chains does not fit in "
                              "device memory: pass a smaller chunk_iterations, store_model_output=False and/or "
                              "store_coarse_chain=False" % (chunk_iterations, n_local)) from None
        raise
    if shared and world > 1:
        parallel.connect_dream_peers(eng, rank, world)      # in-kernel archive exchange over NVLink peer memory
    _lap("engine construction")
    print("Sampling {} chains in lock-step on GPU {}".format(n_chains, device))
    eng.init(theta0)
    _lap("init (H2D + initial Links, enqueue)")

    top = n_levels - 1
    n_qoi = [int(lv["model"].get("n_qoi", 0)) for lv in spec["levels"]]
    fields = ["theta", "stats"] + (["output"] if store_model_output else []) + (["qoi"] if n_qoi[top] else [])
    hist = CompactHistory(n_local)
    dense_parts = {l: [] for l in range(top) if store[l]}
    done, slot, pending, block = 0, 0, None, 0
    while True:
        n = min(chunk_iterations, iterations - done)
        if block:
            eng.history_reset()
        if n > 0:
            if shared and world > 1:
                parallel.run_dream_shared(eng, n, rank, world)
            else:
                eng.run(n)
        nrec = n + (1 if block == 0 else 0)          # block 0 starts with the initial Link
        if nrec > 0:
            for l in dense_parts:
                nl = n * steps[l]
                if nl:
                    dense_parts[l].append((eng.fetch(l, "theta", 0, nl, sync=False), eng.fetch(l, "prior", 0, nl, sync=False),
                                           eng.fetch(l, "like", 0, nl, sync=False), eng.fetch(l, "accept", 0, nl, sync=False),
                                           eng.fetch(l, "output", 0, nl, sync=False) if store_model_output else None,
                                           eng.fetch(l, "qoi", 0, nl, sync=False) if n_qoi[l] else None))
            eng.compact_begin(0, nrec, block == 0, fields, slot)
            if pending is not None:
                hist.append(eng.compact_collect(pending))
            pending, slot = slot, slot ^ 1
        done += n
        block += 1
        if done >= iterations:
            break
    _lap("blocks enqueued (run + compaction + collect of the previous block)")
    if pending is not None:
        hist.append(eng.compact_collect(pending))
    eng.compact_sync()
    eng.sync()
    _lap("last block: collect + sync")

    # result dict, sampler.py:305-309, :406-439, :510-547
    if n_levels == 1:
        result = SampleResult({"sampler": "MH", "n_chains": n_chains, "iterations": iterations + 1})
        keys = ["chain_{}"]
    elif n_levels == 2:
        result = SampleResult({"sampler": "DA", "n_chains": n_chains, "iterations": iterations + 1,
                               "subchain_length": subchain_length})
        keys = ["chain_coarse_{}", "chain_fine_{}"]
    else:
        result = SampleResult({"sampler": "MLDA", "n_chains": n_chains, "iterations": iterations + 1,
                               "levels": n_levels, "subchain_lengths": list(spec["J"])})
        keys = ["chain_l%d_{}" % l for l in range(n_levels)]
    for l in reversed(range(top)):
        if not store_coarse_chain:
            factory = lambda c: None                  # sampler.py:429-431, :541-543
        elif dense_parts[l]:
            factory = _dense_family(dense_parts[l])
        else:
            factory = lambda c: []                    # no iterations: no coarse Link exists yet
        result.add_chains(keys[l], lo, hi, factory)
    result.add_chains(keys[top], lo, hi, hist.chain)
    result.history = hist                     # the finest level of the local chains, compacted (link.CompactHistory)
    result.local_chains = (lo, hi)
    if shared:
        result.dream_exchange = ("peer-memory" if getattr(eng, "peers_connected", False) else
                                 "nccl all-gather per step" if world > 1 else "single device")
    if world > 1:
        result["local_chains"] = (lo, hi)
    if return_engine:
        return result, eng
    eng.close()
    _lap("result dict + engine close")
    if _prof is not None:
        prev = _t0
        for name, t in _prof:
            sys.stderr.write("[tda.sample] %-70s %8.2f ms\n" % (name, (t - prev) * 1e3))
            prev = t
        sys.stderr.write("[tda.sample] %-70s %8.2f ms (%d blocks, %d compact rows)\n"
                         % ("total", (prev - _t0) * 1e3, len(hist.chunks), hist.n_rows()))
    return result
