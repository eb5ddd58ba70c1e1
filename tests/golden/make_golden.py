"""Generates the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference
(tinyDA, imported from /root/reference) under injected random streams.

Run in the development container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [case ...]

Every fixture holds: the lowered problem spec (what the engine / oracle consume), the
initial states, the pre-drawn normal and uniform streams, and the reference's per-level
histories (parameters, log-prior, log-likelihood, accept flags, optionally model outputs)
exactly as its chain objects recorded them:
  level L-1 (finest)  chain.chain / chain.accepted            (iterations+1 records)
  level l < L-1       compress(chain, is_local|is_coarse)      (local records only)
Chains are driven one OS-process-equivalent at a time like tinyDA/ray.py:192-210 does (one
chain object per chain, posteriors deep-copied per chain like Ray's pickling does).
"""
import sys
import os
import copy
from itertools import compress

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refharness as rh            # noqa: E402
import problems                    # noqa: E402
import tinyda_b200 as ours         # noqa: E402
from tinyda_b200.lowering import spec_to_flat   # noqa: E402


def _hist(links, accepted, store_F):
    out = dict(theta=np.array([l.parameters for l in links], dtype=np.float64),
               prior=np.array([l.prior for l in links], dtype=np.float64),
               like=np.array([l.likelihood for l in links], dtype=np.float64),
               acc=np.array(list(accepted), dtype=bool))
    if store_F:
        out["F"] = np.array([l.model_output for l in links], dtype=np.float64)
    return out


def run_reference_chain(tda, posts, prop, kw, theta0, iterations, store_F):
    """One reference chain object, like tinyDA/ray.py:192-210 / sampler.py:295-368."""
    posts = copy.deepcopy(posts)
    prop = copy.deepcopy(prop)
    L = len(posts)
    if L == 1:
        ch = tda.Chain(posts[0], prop, theta0)
        ch.sample(iterations, progressbar=False)
        return [_hist(ch.chain, ch.accepted, store_F)], ch
    if L == 2:
        ch = tda.DAChain(posts[0], posts[1], prop, kw["subchain_length"],
                         kw.get("randomize_subchain_length", False), theta0,
                         kw.get("adaptive_error_model"), True)
        ch.sample(iterations, progressbar=False)
        coarse = _hist(list(compress(ch.chain_coarse, ch.is_coarse)),
                       list(compress(ch.accepted_coarse, ch.is_coarse)), store_F)
        fine = _hist(ch.chain_fine, ch.accepted_fine, store_F)
        return [coarse, fine], ch
    ch = tda.MLDAChain(posts, prop, kw["subchain_length"], theta0,
                       kw.get("adaptive_error_model"), True)
    ch.sample(iterations, progressbar=False)
    hists = [None] * L
    hists[L - 1] = _hist(ch.chain, ch.accepted, store_F)
    cur = ch.proposal
    for l in range(L - 2, -1, -1):
        hists[l] = _hist(list(compress(cur.chain, cur.is_local)),
                         list(compress(cur.accepted, cur.is_local)), store_F)
        cur = cur.proposal
    return hists, ch


def run_reference_dream_shared(tda, posts, prop, theta0, iterations, z, u, store_F):
    """DREAM with the shared archive, stepped round-robin one iteration at a time
    (Chain.sample is resumable, chain.py:78-129) with an ArchiveManager shim that makes the
    rows pushed during a round visible only after the round: the deterministic lock-step
    rule documented in DESIGN.md."""
    C = theta0.shape[0]

    class Archive:
        def __init__(self):
            self.visible = [None] * C
            self.pending = []

        def update_archive(self, sample, chain_id):
            self.pending.append((chain_id, np.array(sample)))

        def flush(self):
            for cid, s in self.pending:
                self.visible[cid] = s[None, :] if self.visible[cid] is None else np.vstack((self.visible[cid], s))
            self.pending = []

        def get_archive(self):
            return np.concatenate(self.visible)

    arch = Archive()

    class Handle:           # what ray's actor handle looks like to SharedArchiveProposal
        class _M:
            def __init__(self, fn):
                self.fn = fn

            def remote(self, *a):
                return self.fn(*a)

        def __getattr__(self, name):
            return Handle._M(getattr(arch, name))

        def __deepcopy__(self, memo):
            return self

    chains, streams, archive0 = [], [], []
    for c in range(C):
        p = copy.deepcopy(prop)
        p.link_archive(Handle())
        p.set_id(c)
        S = rh.Streams(z[c], u[c])
        np.random.seed(1000 + c)           # feeds prior.rvs(M0) for the initial archive only
        with rh.injected(S):
            ch = tda.Chain(copy.deepcopy(posts[0]), p, theta0[c])
        archive0.append(np.array(p.Z))
        chains.append(ch)
        streams.append(S)
    arch.flush()
    for _ in range(iterations):
        for c in range(C):
            with rh.injected(streams[c]):
                chains[c].sample(1, progressbar=False)
        arch.flush()
    hists = [[_hist(ch.chain, ch.accepted, store_F)] for ch in chains]
    return hists, np.array(archive0), streams


def make(name):
    tda = rh.import_reference()
    defn = problems.CASES[name]()
    posts_ref, prop_ref, kw = defn["build"](tda)
    posts_our, prop_our, kw2 = defn["build"](ours)
    spec = ours.lower_problem(posts_our, prop_our, kw2.get("subchain_length"),
                              kw2.get("adaptive_error_model"), kw2.get("randomize_subchain_length", False))
    C, iters, seed = defn["n_chains"], defn["iterations"], defn["seed"]
    store_F = defn.get("store_F", True)
    rng = np.random.default_rng(seed)
    nz, nu = problems.stream_sizes(spec, iters)
    z = rng.standard_normal((C, nz))
    u = rng.random((C, nu))
    theta0 = np.atleast_2d(defn["prior"].rvs(C, random_state=rng)).reshape(C, -1)

    out = dict(theta0=theta0, z=z, u=u, iterations=np.array(iters))
    L = spec["n_levels"]
    if defn.get("shared"):
        hists, archive0, streams = run_reference_dream_shared(
            tda, posts_ref, prop_ref, theta0, iters, z, u, store_F)
        out["archive0"] = archive0
        consumed = np.array([[S.nz, S.nu] for S in streams])
    else:
        hists, consumed, archive0 = [], [], []
        for c in range(C):
            S = rh.Streams(z[c], u[c])
            np.random.seed(1000 + c)       # only feeds scipy's prior.rvs() calls in setup_proposal
            with rh.injected(S):
                h, ch = run_reference_chain(tda, posts_ref, prop_ref, kw, theta0[c], iters, store_F)
            hists.append(h)
            consumed.append([S.nz, S.nu])
            if defn.get("archive"):
                base = ch.proposal                      # MLDA nests the base proposal (proposal.py:1395-1440)
                while not hasattr(base, "Z"):
                    base = base.proposal
                archive0.append(np.array(base.Z[:spec["proposal"]["M0"]]))
        consumed = np.array(consumed)
        if defn.get("archive"):
            out["archive0"] = np.array(archive0)
    out["consumed"] = consumed
    for l in range(L):
        for k in hists[0][l]:
            out["ref/l%d/%s" % (l, k)] = np.stack([hists[c][l][k] for c in range(C)])
    out.update(spec_to_flat(spec))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-22s levels=%d chains=%d iters=%d consumed(z,u)=%s -> %s (%.1f KB)"
          % (name, L, C, iters, consumed[0].tolist(), os.path.basename(path), os.path.getsize(path) / 1024))


def make_long(name="da_pcn_cfg2", C=8, iters=200, seed=143):
    """BASELINE cfg2 at its real shape, 8 chains x 200 fine iterations (2000 coarse steps per chain): the
    fixture the float32 kernels' measured state error and decision-flip rate are quoted on
    (tests/golden/long/<name>.npz; tests/test_gpu_fp32_error.py, tests/test_oracle_golden.py).
    To keep it small the normals sit on the fp16 grid at scale 4096 (stored as float16 of 4096 z: exactly
    what the tensor-core kernel's operand image holds) and the coarse level keeps log-likelihood and accept
    flags only."""
    tda = rh.import_reference()
    defn = problems.CASES[name]()
    posts_ref, prop_ref, kw = defn["build"](tda)
    posts_our, prop_our, kw2 = defn["build"](ours)
    spec = ours.lower_problem(posts_our, prop_our, kw2.get("subchain_length"))
    rng = np.random.default_rng(seed)
    nz, nu = problems.stream_sizes(spec, iters)
    z16 = (rng.standard_normal((C, nz)) * 4096.0).astype(np.float16)
    z = z16.astype(np.float64) / 4096.0
    u = rng.random((C, nu))
    theta0 = np.atleast_2d(defn["prior"].rvs(C, random_state=rng)).reshape(C, -1)
    out = dict(theta0=theta0, z16=z16, u=u, iterations=np.array(iters))
    hists, consumed = [], []
    for c in range(C):
        S = rh.Streams(z[c], u[c])
        np.random.seed(1000 + c)
        with rh.injected(S):
            h, ch = run_reference_chain(tda, posts_ref, prop_ref, kw, theta0[c], iters, False)
        hists.append(h)
        consumed.append([S.nz, S.nu])
        print("chain", c, "fine accept rate", h[1]["acc"].mean(), "coarse", h[0]["acc"].mean())
    out["consumed"] = np.array(consumed)
    for l, keys in ((0, ("like", "acc")), (1, ("theta", "prior", "like", "acc"))):
        for k in keys:
            out["ref/l%d/%s" % (l, k)] = np.stack([hists[c][l][k] for c in range(C)])
    out.update(spec_to_flat(spec))
    os.makedirs(os.path.join(HERE, "long"), exist_ok=True)
    path = os.path.join(HERE, "long", name + ".npz")
    np.savez_compressed(path, **out)
    print("%s: %d chains x %d iterations -> %s (%.1f KB)" % (name, C, iters, path, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    if sys.argv[1:2] == ["--long"]:
        make_long()
        sys.exit(0)
    names = sys.argv[1:] or list(problems.CASES)
    for n in names:
        make(n)
