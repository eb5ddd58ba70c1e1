"""Lowering of the user's tinyDA-style objects into the plain problem spec the engine uploads.

The spec is a nested dict of NumPy arrays and Python scalars (no objects): it is what
``engine.Engine`` turns into the POD ``tda_config`` + constant buffers of the C ABI
(include/tinyda_b200.h), what the golden fixtures store, and what the test oracle consumes.
"""
import numpy as np

from .posterior import Posterior, lower_prior
from .proposal import (PROP_RWMH, PROP_PCN, PROP_AM, PROP_MALA, PROP_DREAMZ, PROP_DREAM,
                       GaussianRandomWalk, MultipleTry, IndependenceSampler, PROP_INDEP)
from .distributions import LIK_ADAPTIVE
from .models import MODEL_LINEAR, MODEL_ROSENBROCK

MAX_LEVELS = 4
MAX_D = 64


def lower_problem(posteriors, proposal, subchain_lengths=None, adaptive_error_model=None,
                  randomize_subchain_length=False):
    if isinstance(posteriors, Posterior):
        posteriors = [posteriors]
    L = len(posteriors)
    if L < 1 or L > MAX_LEVELS:
        raise ValueError("the engine supports 1..%d levels" % MAX_LEVELS)
    if not isinstance(proposal, (GaussianRandomWalk, MultipleTry, IndependenceSampler)):
        raise TypeError("proposal %s cannot be lowered to the device engine"
                        % type(proposal).__name__)
    prior = lower_prior(posteriors[0].prior)
    d = prior["mean"].shape[0]
    if d > MAX_D:
        raise ValueError("the engine supports up to %d parameters" % MAX_D)
    levels = [p.lower() for p in posteriors]
    for lv in levels:
        if lv["model"]["d"] != d:
            raise ValueError("model parameter dimension does not match the prior")
    if L > 1:
        if isinstance(subchain_lengths, (int, np.integer)):
            J = [int(subchain_lengths)] * (L - 1)
        else:
            J = [int(j) for j in subchain_lengths]
        if len(J) != L - 1:
            raise ValueError("subchain_length must have len(posteriors)-1 entries")
    else:
        J = []
    aem = 0
    if adaptive_error_model is not None and L > 1:
        if adaptive_error_model not in ("state-independent", "state-dependent"):
            raise ValueError("Adaptive error model can only be state-dependent, state-independent or None.")
        aem = 1
        if adaptive_error_model == "state-dependent":
            if L != 2:       # sampler.py:184-188 falls back before this point
                raise ValueError("the state-dependent error model is a two-level method")
            aem = 2
        ms = {lv["model"]["m"] for lv in levels}
        if len(ms) != 1:
            raise ValueError("the adaptive error model needs equal output sizes on all levels")
        for lv in levels[:-1]:
            if lv["lik"]["kind"] != LIK_ADAPTIVE:
                raise TypeError("coarse likelihoods must be AdaptiveGaussianLogLike when an "
                                "adaptive error model is used")
    prop = proposal.lower(prior)
    if aem == 2 and prop["kind"] not in (PROP_RWMH, PROP_AM, PROP_PCN):
        # chain.py:456-460 needs is_symmetric or a working get_q(fine link, fine link): MALA's get_q
        # reads a gradient the fine links do not carry and DREAM(Z) inherits a get_q that returns None
        raise TypeError("the state-dependent error model needs a symmetric proposal or CrankNicolson")
    if prop.get("mtm_k", 0) and aem == 2:
        raise NotImplementedError("MultipleTry with the state-dependent error model is not lowered")
    randomize = 0
    if randomize_subchain_length:
        if L != 2:
            raise ValueError("randomize_subchain_length is a two-level (Delayed Acceptance) option")
        if J[0] == 1:                                                # chain.py:311-312
            raise ValueError("Randomize subchain length requires a subchain_length > 1.")
        if prop["kind"] in (PROP_DREAMZ, PROP_DREAM):
            raise NotImplementedError("randomize_subchain_length with a DREAM(Z) base proposal is not lowered "
                                      "(the index draw is read ahead of a fixed number of uniforms per step)")
        randomize = 1
    if prop["kind"] == PROP_MALA:
        if L != 1:
            raise NotImplementedError("MALA is lowered for single-level sampling only")
        if levels[0]["model"]["kind"] not in (MODEL_LINEAR, MODEL_ROSENBROCK):
            raise TypeError("MALA needs a model with an analytic gradient")
    if prop["kind"] == PROP_INDEP and L != 1:
        raise NotImplementedError("IndependenceSampler is lowered for single-level sampling only")
    if prop["kind"] == PROP_DREAM and L != 1:
        # the shared archive needs a grid-wide boundary after every base-level step
        raise NotImplementedError("DREAM with the shared archive is lowered for single-level sampling only "
                                  "(use DREAMZ as the base proposal of DA / MLDA)")
    return dict(n_levels=L, d=d, J=J, aem=aem, randomize=randomize, prior=prior, levels=levels,
                proposal=prop)


# ---- (de)serialisation for golden fixtures ------------------------------------------------
def spec_to_flat(spec, prefix="spec"):
    """Flatten a spec into {key: ndarray} for np.savez."""
    out = {}

    def rec(obj, key):
        if isinstance(obj, dict):
            out[key + "/__dict__"] = np.array(sorted(obj.keys()))
            for k, v in obj.items():
                rec(v, key + "/" + k)
        elif isinstance(obj, (list, tuple)) and (len(obj) == 0 or isinstance(obj[0], dict)):
            out[key + "/__list__"] = np.array(len(obj))
            for i, v in enumerate(obj):
                rec(v, key + "/%d" % i)
        else:
            out[key] = np.asarray(obj)

    rec(spec, prefix)
    return out


def spec_from_flat(flat, prefix="spec"):
    def rec(key):
        if key + "/__dict__" in flat:
            return {str(k): rec(key + "/" + str(k)) for k in flat[key + "/__dict__"]}
        if key + "/__list__" in flat:
            return [rec(key + "/%d" % i) for i in range(int(flat[key + "/__list__"]))]
        v = flat[key]
        if v.ndim == 0:
            v = v.item()
        return v

    spec = rec(prefix)
    if not isinstance(spec.get("J", []), list):
        spec["J"] = [int(j) for j in np.atleast_1d(spec["J"])]
    return spec
