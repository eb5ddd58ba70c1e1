"""``sample()``: the drop-in for ``tinyDA.sample`` (tinyDA/sampler.py:21-292).

Same signature, validation, warnings and result dict.  Where the reference builds one Python
chain object per chain and runs them one after another or as Ray actors
(sampler.py:295-508, ray.py:12-210), this lowers the problem once and advances ALL chains in
lock-step on the GPU (``engine.Engine`` -> C ABI -> CUDA kernels).  Extra keyword-only
arguments select engine options (dtype, RNG mode, storage); the positional / reference
keywords keep their meaning.  Under ``torch.distributed`` (one process per GPU) ``n_chains``
is the global count and every rank samples its contiguous shard.
"""
import warnings

import numpy as np
import scipy.stats as stats

from .posterior import Posterior
from .proposal import CrankNicolson, DREAMZ, DREAM, PROP_DREAM, PROP_DREAMZ, PROP_AM
from .lowering import lower_problem
from .link import LinkSequence
from . import parallel


def _result_sequences(eng, level, n_chains_local, store_output):
    theta = eng.fetch(level, "theta")            # [nrec, d, C]
    prior = eng.fetch(level, "prior")            # [nrec, C]
    like = eng.fetch(level, "like")
    acc = eng.fetch(level, "accept").astype(bool)
    out = eng.fetch(level, "output") if store_output else None
    seqs = []
    for c in range(n_chains_local):
        seqs.append(LinkSequence(theta[:, :, c], prior[:, c], like[:, c],
                                 None if out is None else out[:, :, c], acc[:, c]))
    return seqs


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
):
    """Returns MCMC samples given a Posterior (or a list of them, coarsest first) and a proposal,
    exactly like ``tinyDA.sample``: one posterior -> Metropolis-Hastings, two -> Delayed
    Acceptance, more -> MLDA.  ``force_sequential`` / ``force_progress_bar`` are accepted for
    compatibility and ignored (there is nothing sequential to force).

    Engine options (keyword-only): ``dtype`` 'float64' | 'float32'; ``rng`` 'philox' (per-chain
    counter-based streams generated in-kernel, seeded by ``seed``) or 'injected' with
    ``streams=(normals[n_chains, nz], uniforms[n_chains, nu])``; ``store_model_output`` keeps
    F(theta) of every stored link (Link.model_output) -- switch it off for large runs.
    """
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
        elif type(initial_parameters) == np.ndarray:
            assert (
                d_prior == initial_parameters.size
            ), "If an array of initial parameters is provided, it must have the same dimension as the prior"
            initial_parameters = [initial_parameters] * n_chains
        else:
            raise TypeError("Initial paramaters must be list, numpy array or None")
        theta0 = np.array([np.atleast_1d(t) for t in initial_parameters[lo:hi]], dtype=np.float64)
    else:
        host_rng = np.random.default_rng(None if seed is None else [int(seed), 7])
        theta0 = np.atleast_2d(posteriors[0].prior.rvs(n_chains, random_state=host_rng)).reshape(n_chains, -1)[lo:hi]

    spec = lower_problem(posteriors, proposal, subchain_length if n_levels > 1 else None,
                         adaptive_error_model if n_levels > 1 else None,
                         randomize_subchain_length if n_levels == 2 else False)
    kind = int(spec["proposal"]["kind"])

    archive0 = None
    if kind in (PROP_DREAMZ, PROP_DREAM):
        M0 = int(spec["proposal"]["M0"])
        if initial_archive is not None:
            archive0 = np.asarray(initial_archive, dtype=np.float64)
        else:                                                           # proposal.py:788, per chain
            host_rng = np.random.default_rng(None if seed is None else [int(seed), 11])
            base = proposal.kernel if hasattr(proposal, "kernel") else proposal
            archive0 = np.stack([base.initial_archive(posteriors[0].prior, host_rng) for _ in range(n_chains)])
        if kind == PROP_DREAMZ:
            archive0 = archive0[lo:hi]

    from .engine import Engine, STORE_FULL, STORE_STATS, STORE_NONE   # loads the CUDA library
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
    from ._lib import EngineError
    try:
        eng = Engine(spec, n_local, dtype=dtype, rng=rng, seed=0 if seed is None else seed, store=store,
                     capacity_iterations=iterations, streams=streams, device=device,
                     chain_offset=lo, n_chains_global=n_chains if shared else n_local,
                     archive0=archive0, am_device_refactor=True)
    except EngineError as exc:
        if "cudaMalloc" in str(exc):
            # the reference keeps every Link (with its model output) of every level in host lists;
            # here that history lives in HBM until it is fetched
            raise EngineError(str(exc) + " -- the Link history of %d chains x %d iterations does not fit in device "
                              "memory: pass store_model_output=False and/or store_coarse_chain=False, or "
                              "sample in several calls" % (n_local, iterations)) from None
        raise
    if kind == PROP_DREAMZ:
        # per-chain archives live in the same [slot][chain][d] array, indexed by the local chain
        pass
    print("Sampling {} chains in lock-step on GPU {}".format(n_chains, device))
    eng.init(theta0)
    if shared and world > 1:
        parallel.run_dream_shared(eng, iterations, rank, world)
    else:
        eng.run(iterations)
    eng.sync()

    # result dict, sampler.py:305-309, :406-439, :510-547
    if n_levels == 1:
        info = {"sampler": "MH", "n_chains": n_chains, "iterations": iterations + 1}
        seqs = _result_sequences(eng, 0, n_local, store_model_output)
        chains = {"chain_{}".format(lo + i): s for i, s in enumerate(seqs)}
        result = {**info, **chains}
    elif n_levels == 2:
        info = {"sampler": "DA", "n_chains": n_chains, "iterations": iterations + 1,
                "subchain_length": subchain_length}
        if store_coarse_chain:
            seqs = _result_sequences(eng, 0, n_local, store_model_output)
            coarse = {"chain_coarse_{}".format(lo + i): s for i, s in enumerate(seqs)}
        else:
            coarse = {"chain_coarse_{}".format(lo + i): None for i in range(n_local)}
        seqs = _result_sequences(eng, 1, n_local, store_model_output)
        fine = {"chain_fine_{}".format(lo + i): s for i, s in enumerate(seqs)}
        result = {**info, **coarse, **fine}
    else:
        info = {"sampler": "MLDA", "n_chains": n_chains, "iterations": iterations + 1,
                "levels": n_levels, "subchain_lengths": list(spec["J"])}
        result = dict(info)
        for l in reversed(range(n_levels)):
            if l == n_levels - 1 or store_coarse_chain:
                seqs = _result_sequences(eng, l, n_local, store_model_output)
                result.update({"chain_l{}_{}".format(l, lo + i): s for i, s in enumerate(seqs)})
            else:
                result.update({"chain_l{}_{}".format(l, lo + i): None for i in range(n_local)})
    if world > 1:
        result["local_chains"] = (lo, hi)
    if return_engine:
        return result, eng
    eng.close()
    return result
