"""Multi-GPU plumbing: one process per GPU, chains sharded in contiguous ranges.

The reference's only parallelism is one Ray actor per chain (tinyDA/ray.py:12-210) and its only
inter-chain communication is the DREAM ArchiveManager actor (ray.py:366-384).  Here
MH / DA / MLDA / MALA / AM chains never communicate; DREAM's shared archive is replicated in
every GPU's HBM and each step's new rows are all-gathered over NCCL (NVLink / NVSwitch); the
final R-hat / ESS reduction all-reduces per-chain moment sums.  Everything in this module is
plain torch.distributed and works on the gloo backend with CPU tensors too (that is how the
host-side logic is tested without GPUs).
"""
import os

import numpy as np


def _dist():
    # a process group can only exist if the caller has already imported torch: do not pay for the
    # import (about a second) in single-process use
    import sys
    if "torch" not in sys.modules:
        return None
    try:
        import torch.distributed as dist
    except Exception:
        return None
    return dist if (dist.is_available() and dist.is_initialized()) else None


def rank_world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d is not None else (0, 1)


def local_device():
    return int(os.environ.get("LOCAL_RANK", "0")) if _dist() is not None else 0


def broadcast_int(value, src=0):
    """Rank `src`'s integer on every rank (identity without a process group)."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return int(value)
    import torch
    t = torch.tensor([int(value)], dtype=torch.int64)
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.broadcast(t, src)
    return int(t.cpu()[0])


def shard_range(n_chains, rank, world):
    """Contiguous chain-id range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n_chains), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def archive_row_to_chain_slot(r, n_slots):
    """Flat row index of the shared archive -> (chain, slot): chain-major order of
    np.concatenate(shared_archive) (ray.py:381) when every chain holds n_slots rows."""
    return r // n_slots, r % n_slots


class _CudaView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


def archive_tensor(eng):
    """Zero-copy torch view [capacity, n_chains_global, d] of the engine's DREAM archive."""
    import torch
    ptr, nbytes = eng.device_buffer("dream_archive")
    item = eng.dtype.itemsize
    cap = nbytes // (item * eng.Cg * eng.d)
    typestr = "<f4" if item == 4 else "<f8"
    return torch.as_tensor(_CudaView(ptr, (cap, eng.Cg, eng.d), typestr), device="cuda:%d" % eng.device)


def allgather_rows(rows, lo, hi, world):
    """In-place all-gather of one archive slot: `rows` is [n_chains_global, d] on every rank,
    rank r has filled rows[lo:hi]; afterwards every rank holds all rows.  NCCL runs it in place
    (send buffer is a slice of the receive buffer) when the shards are equal."""
    d = _dist()
    if d is None or world == 1:
        return rows
    n = rows.shape[0]
    if n % world == 0:
        d.all_gather_into_tensor(rows, rows[lo:hi].clone() if rows.device.type == "cpu" else rows[lo:hi])
    else:
        # unequal shards: pad every rank's block to the largest shard
        import torch
        sizes = [shard_range(n, r, world) for r in range(world)]
        big = max(b - a for a, b in sizes)
        mine = torch.zeros((big,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        mine[: hi - lo] = rows[lo:hi]
        outs = [torch.empty_like(mine) for _ in range(world)]
        d.all_gather(outs, mine)
        for (a, b), o in zip(sizes, outs):
            rows[a:b] = o[: b - a]
    return rows


def connect_dream_peers(eng, rank, world):
    """Maps every rank's archive replica and step flags into the others (cudaIpc handles exchanged with one
    all-gather) so that the engine's persistent kernel exchanges each step's rows through NVLink peer memory
    itself.  Returns False -- and the per-step NCCL all-gather of run_dream_shared is used -- when the ranks are
    not NCCL ranks of one node or the mapping fails on any of them."""
    import torch
    d = _dist()
    if d is None or world == 1 or d.get_backend() != "nccl":
        return False
    ok = 1
    try:
        mine = eng.peer_export()
    except Exception:
        mine, ok = np.zeros(192, dtype=np.uint8), 0
    dev = "cuda:%d" % eng.device
    t = torch.from_numpy(mine.copy()).to(dev)
    allb = torch.empty(world * t.numel(), dtype=torch.uint8, device=dev)
    d.all_gather_into_tensor(allb, t)
    if ok:
        try:
            eng.peer_import(world, rank, allb.cpu().numpy())
        except Exception:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    d.all_reduce(flag, op=d.ReduceOp.MIN)
    good = bool(int(flag.item()))
    if not good:
        eng.peers_connected = False
    return good


def run_dream_shared(eng, iterations, rank, world):
    """DREAM with the shared archive over `world` GPUs.  With the replicas mapped into each other
    (connect_dream_peers) this is one persistent launch: the kernel stores each step's rows into every replica
    and closes the step with a flag handshake.  Otherwise: one lock-step step per launch, then the NCCL
    all-gather of that step's new rows (chain-major archive slot)."""
    M0 = int(eng.spec["proposal"]["M0"])
    if getattr(eng, "peers_connected", False):
        eng.run(iterations)
        return M0
    arch = archive_tensor(eng)
    lo, hi = shard_range(eng.Cg, rank, world)
    for t in range(iterations):
        slot = eng.dream_slots()
        eng.run(1)
        allgather_rows(arch[slot], lo, hi, world)
    return M0


def allreduce_chain_moments(sum1, sum2, n_draws):
    """Classic (non-split) R-hat and the pooled mean / variance from per-chain running sums.
    sum1, sum2: [d, n_chains_local] sums of x and x^2 over n_draws draws.  One all-reduce of
    3*d+1 numbers.  Returns dict(mean, var_within, var_between, rhat, n_chains)."""
    import torch
    sum1 = np.asarray(sum1, dtype=np.float64)
    sum2 = np.asarray(sum2, dtype=np.float64)
    n = float(n_draws)
    mean_c = sum1 / n
    var_c = (sum2 - n * mean_c ** 2) / (n - 1.0)
    d = sum1.shape[0]
    pack = np.concatenate([mean_c.sum(axis=1), (mean_c ** 2).sum(axis=1), var_c.sum(axis=1), [sum1.shape[1]]])
    t = torch.from_numpy(pack)
    dist = _dist()
    if dist is not None:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
        t = t.cpu()
    pack = t.numpy()
    m = pack[-1]
    s_mean, s_mean2, s_var = pack[:d], pack[d:2 * d], pack[2 * d:3 * d]
    grand = s_mean / m
    B_over_n = (s_mean2 - m * grand ** 2) / (m - 1.0)        # variance of the chain means
    W = s_var / m
    var_plus = (n - 1.0) / n * W + B_over_n
    return dict(mean=grand, var_within=W, var_between=B_over_n * n, rhat=np.sqrt(var_plus / W), n_chains=int(m))


def allreduce_ess(x_local):
    """Multi-chain effective sample size per parameter over ALL ranks' chains.  x_local:
    [n_chains_local, n_draws, d] draws held by this rank.  Each rank reduces its chains to
    per-parameter sums (autocovariances by FFT, chain means, squared chain means) and one
    all-reduce of (n_draws + 3) * d numbers finishes the job; the result equals the single-process
    estimate on the concatenated chains (``diagnostics._ess_plain``, the non-rank-normalised ESS of
    Stan / ArviZ).  Returns ess[d]."""
    import torch
    from .diagnostics import _autocov, _ess_from_sums
    x = np.asarray(x_local, dtype=np.float64)
    m, n, d = x.shape
    pack = np.zeros((d, n + 3))
    for k in range(d):
        xk = x[:, :, k]
        means = xk.mean(axis=1)
        pack[k, :n] = _autocov(xk).sum(axis=0)
        pack[k, n] = means.sum()
        pack[k, n + 1] = (means ** 2).sum()
        pack[k, n + 2] = m
    t = torch.from_numpy(pack)
    dist = _dist()
    if dist is not None:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
        t = t.cpu()
    pack = t.numpy()
    return np.array([_ess_from_sums(int(round(pack[k, n + 2])), n, pack[k, :n], pack[k, n], pack[k, n + 1])
                     for k in range(d)])


def allreduce_sums(*arrays):
    """Sums float64 arrays (e.g. the device-side ESS / R-hat sums of Engine.ess_sums) over all ranks; identity
    without a process group.  Over NCCL the reduction runs on the device."""
    import torch
    dist = _dist()
    out = []
    for a in arrays:
        a = np.ascontiguousarray(a, dtype=np.float64)
        if dist is None or dist.get_world_size() == 1:
            out.append(a)
            continue
        t = torch.from_numpy(a.copy())
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
        out.append(t.cpu().numpy())
    return out if len(out) > 1 else out[0]


def allreduce_ess_sums(sums, folded):
    """All-reduces the per-parameter sums of Engine.ess_sums over the ranks (every rank ranks its own chains:
    with thousands of chains per GPU the local empirical distribution is the pooled one to O(1/sqrt(N))); the
    draw count per split chain is a property of the run, not a sum."""
    dist = _dist()
    world = 1 if dist is None else dist.get_world_size()
    sums = np.array(sums, dtype=np.float64)
    folded = np.array(folded, dtype=np.float64)
    if world == 1:
        return sums, folded
    n_half = sums[:, -1].copy()
    s, f = allreduce_sums(sums, folded)
    s[:, -1] = n_half
    return s, f
