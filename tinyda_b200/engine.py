"""Python handle on one device engine (one per GPU / per process).

Turns a lowered problem spec (``lowering.lower_problem``) into the POD ``tda_config`` and
constant uploads of the C ABI, and exposes run / fetch.  All arithmetic happens in the CUDA
library behind ``_lib``; importing this module fails loudly if that library is missing.
"""
import ctypes as C
import os
import sys
import time
import warnings
import weakref

import numpy as np

from . import _lib as L
from ._lib import lib, check
from .proposal import PROP_AM, PROP_MALA, PROP_DREAMZ, PROP_DREAM
from .distributions import LIK_ISO, LIK_DIAG, LIK_DENSE, LIK_ADAPTIVE
from .models import MODEL_LINEAR, MODEL_POISSON1D

STORE_FULL = L.TDA_STORE_THETA | L.TDA_STORE_STATS | L.TDA_STORE_OUTPUT | L.TDA_STORE_ACCEPT
STORE_STATS = L.TDA_STORE_THETA | L.TDA_STORE_STATS | L.TDA_STORE_ACCEPT
STORE_NONE = 0


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# Page-locking is slow (of the order of 0.1-0.3 ms per MiB), so released pinned blocks are kept in a
# small pool keyed by size class and handed out again (a long-lived process that calls sample()
# repeatedly pays for its result buffers once).
_PIN_POOL = {}
_PIN_POOL_STATE = {"bytes": 0, "max_bytes": 16 << 30}


def _pin_class(nbytes):
    # powers of two: the row count of a compacted block varies from call to call, and a call that falls into
    # a class the pool has not seen yet pays for page-locking (hundreds of milliseconds per GiB)
    nbytes = max(int(nbytes), 1 << 16)
    return 1 << (nbytes - 1).bit_length()


def _pin_release(ptr, cls):
    if _PIN_POOL_STATE["bytes"] + cls <= _PIN_POOL_STATE["max_bytes"]:
        _PIN_POOL.setdefault(cls, []).append(ptr)
        _PIN_POOL_STATE["bytes"] += cls
    else:
        lib.tda_host_free(C.c_void_p(ptr))


def pinned_pool_clear():
    for cls, ptrs in _PIN_POOL.items():
        for ptr in ptrs:
            lib.tda_host_free(C.c_void_p(ptr))
    _PIN_POOL.clear()
    _PIN_POOL_STATE["bytes"] = 0


def pinned_empty(shape, dtype):
    """Page-locked host array (cudaHostAlloc through the C ABI): device->host copies into it are
    asynchronous.  Returned to the pool when the array (and every view of it) is garbage collected."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    cls = _pin_class(count * dtype.itemsize)
    free = _PIN_POOL.get(cls)
    if free:
        addr = free.pop()
        _PIN_POOL_STATE["bytes"] -= cls
    else:
        ptr = C.c_void_p()
        check(lib.tda_host_alloc(cls, C.byref(ptr)))
        addr = ptr.value
    buf = (C.c_ubyte * cls).from_address(addr)
    weakref.finalize(buf, _pin_release, addr, cls)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


class CompactChunk:
    """Host copy of one compacted block of finest-level records (tda_compact_*): the accept flag of
    every record and the fields of the accepted records, chain-major (rows offsets[c]..offsets[c+1]
    belong to chain c, in time order)."""

    __slots__ = ("nrec", "accept", "offsets", "theta", "prior", "like", "output", "qoi")

    def __init__(self, nrec):
        self.nrec = nrec
        self.accept = self.offsets = self.theta = self.prior = self.like = self.output = self.qoi = None


def steps_per_iteration(spec):
    """Local steps each level takes per finest-level iteration."""
    Ln = spec["n_levels"]
    steps = [1] * Ln
    for l in range(Ln - 2, -1, -1):
        steps[l] = steps[l + 1] * int(spec["J"][l])
    return steps


class Engine:
    def __init__(self, spec, n_chains, dtype="float64", rng="philox", seed=0, store=None,
                 capacity_iterations=0, streams=None, device=0, chain_offset=0, n_chains_global=None,
                 archive0=None, am_device_refactor=True, stream=None, archive_iterations=None):
        self.spec = spec
        self.C = int(n_chains)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype("float32"), np.dtype("float64")):
            raise ValueError("dtype must be float32 or float64")
        self.Ln = Ln = int(spec["n_levels"])
        self.d = d = int(spec["d"])
        self.device = int(device)
        self.stream = stream
        prop = spec["proposal"]
        kind = int(prop["kind"])
        self.steps = steps_per_iteration(spec)
        if store is None:
            store = [STORE_STATS] * Ln
        elif isinstance(store, int):
            store = [store] * Ln
        self.store = list(store)

        cfg = L.Config()
        cfg.abi_version = L.TDA_ABI_VERSION
        cfg.dtype = L.TDA_F64 if self.dtype == np.float64 else L.TDA_F32
        cfg.n_levels = Ln
        cfg.d = d
        for i, j in enumerate(spec["J"]):
            cfg.subchain[i] = int(j)
        cfg.aem = int(spec.get("aem", 0))
        cfg.randomize_subchain = int(spec.get("randomize", 0))
        cfg.mtm_k = int(prop.get("mtm_k", 0))
        cfg.mtm_include_current = int(prop.get("mtm_include_current", 0))
        cfg.rng_mode = L.TDA_RNG_INJECTED if rng == "injected" else L.TDA_RNG_PHILOX
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.n_chains = self.C
        cfg.chain_offset = int(chain_offset)
        cfg.n_chains_global = int(n_chains_global if n_chains_global is not None else self.C)
        self.Cg = int(cfg.n_chains_global)
        cfg.prop_kind = kind
        cfg.adaptive = int(bool(prop.get("adaptive", False)))
        cfg.period = int(prop.get("period", 100))
        cfg.scaling = float(prop.get("scaling", 1.0))
        cfg.gamma = float(prop.get("gamma", 1.01))
        cfg.alpha_star = float(prop.get("alpha_star", 0.24))
        if kind == PROP_AM:
            cfg.am_sd = float(prop["am_sd"])
            cfg.am_eps = float(prop["am_eps"])
            cfg.am_t0 = int(prop["am_t0"])
            cfg.am_device_refactor = int(bool(am_device_refactor))
        if kind in (PROP_DREAMZ, PROP_DREAM):
            cfg.dream_M0 = int(prop["M0"])
            cfg.dream_delta = int(prop["delta"])
            cfg.dream_nCR = int(prop["nCR"])
            cfg.dream_b = float(prop["b"])
            cfg.dream_b_star = float(prop["b_star"])
            cfg.dream_sync_every = int(prop.get("sync_every", 1))
            # one archive row per base-level step (proposal.py:794): steps[0] of them per finest iteration
            cfg.dream_capacity = int(prop["M0"]) + int(capacity_iterations if archive_iterations is None else archive_iterations) * self.steps[0] + 1
        if rng == "injected":
            z, u = streams
            z = np.ascontiguousarray(z, dtype=np.float64)
            u = np.ascontiguousarray(u, dtype=np.float64)
            cfg.stream_z_len = z.shape[1]
            cfg.stream_u_len = u.shape[1]
        cfg.prior_logconst = float(spec["prior"]["logconst"])
        self.capacity = []
        for l, lv in enumerate(spec["levels"]):
            lc = cfg.level[l]
            lc.model_kind = int(lv["model"]["kind"])
            lc.m = int(lv["model"]["m"])
            lc.n_grid = int(lv["model"]["n_grid"])
            lc.lik_kind = int(lv["lik"]["kind"])
            lc.lik_var = float(lv["lik"]["var"]) if lc.lik_kind == LIK_ISO else 0.0
            for i in range(4):
                lc.model_scalars[i] = float(np.asarray(lv["model"]["scalars"])[i])
            lc.store = int(self.store[l])
            lc.n_qoi = int(lv["model"].get("n_qoi", 0))
            cap = int(capacity_iterations) * self.steps[l] + (1 if l == Ln - 1 else 0)
            lc.hist_capacity = cap if lc.store else 0
            self.capacity.append(int(lc.hist_capacity))
        self.cfg = cfg
        self._h = C.c_void_p()
        _t0 = time.perf_counter()
        check(lib.tda_engine_create(C.byref(cfg), self.device, C.byref(self._h)))
        _t1 = time.perf_counter()

        # constants
        pr = spec["prior"]
        self._up(L.TDA_UP_PRIOR_MEAN, 0, pr["mean"])
        self._up(L.TDA_UP_PRIOR_LP, 0, pr["LP"])
        if kind == PROP_MALA:
            self._up(L.TDA_UP_PRIOR_PREC, 0, np.linalg.inv(pr["cov"]))     # utils.py:272-278
        if "T" in prop:
            self._up(L.TDA_UP_PROP_T, 0, prop["T"])
        if "S" in prop:
            self._up(L.TDA_UP_PROP_S, 0, prop["S"])
        if "S2" not in prop and "ow_lambda" in prop:
            self._up(L.TDA_UP_PROP_LAMBDA, 0, prop["ow_lambda"])
        if "S2" in prop:
            self._up(L.TDA_UP_PROP_S2, 0, prop["S2"])
            self._up(L.TDA_UP_PROP_LAMBDA, 0, prop["ow_lambda"])
        for l, lv in enumerate(spec["levels"]):
            mk = int(lv["model"]["kind"])
            if mk in (MODEL_LINEAR, MODEL_POISSON1D):
                self._up(L.TDA_UP_MODEL_A, l, lv["model"]["A"])
            if mk == MODEL_LINEAR:
                self._up(L.TDA_UP_MODEL_B, l, lv["model"]["b"])
            if int(lv["model"].get("n_qoi", 0)):
                self._up(L.TDA_UP_QOI_W, l, lv["model"]["qoi_Q"])
                self._up(L.TDA_UP_QOI_B, l, lv["model"]["qoi_q0"])
            lk = int(lv["lik"]["kind"])
            self._up(L.TDA_UP_LIK_DATA, l, lv["lik"]["data"])
            if lk == LIK_DIAG:
                self._up(L.TDA_UP_LIK_VAR, l, lv["lik"]["var"])
            if lk == LIK_DENSE:
                self._up(L.TDA_UP_LIK_PREC, l, np.linalg.inv(lv["lik"]["cov"]))   # distributions.py:280
            if lk == LIK_ADAPTIVE:
                # r^T inv(cov) r = |Li r|^2 with Li = inv(chol(cov)); the device re-factorises
                # cov + cov_bias per chain whenever set_bias would re-invert (distributions.py:399-402)
                self._up(L.TDA_UP_LIK_PREC, l, np.linalg.inv(np.linalg.cholesky(lv["lik"]["cov"])))
            if lk == LIK_ADAPTIVE:
                self._up(L.TDA_UP_LIK_COV, l, lv["lik"]["cov"])
        if rng == "injected":
            self._up(L.TDA_UP_STREAM_Z, 0, z)
            self._up(L.TDA_UP_STREAM_U, 0, u)
        if kind in (PROP_DREAMZ, PROP_DREAM):
            if archive0 is None:
                raise ValueError("DREAM(Z) needs the initial archive [n_chains_global, M0, d]")
            self._up(L.TDA_UP_DREAM_ARCHIVE0, 0, archive0)
        self.iterations_done = 0
        self.peers_connected = False
        if os.environ.get("TDA_PROFILE"):
            sys.stderr.write("[Engine] tda_engine_create %.2f ms, uploads %.2f ms\n" % ((_t1 - _t0) * 1e3, (time.perf_counter() - _t1) * 1e3))

    # ---- plumbing --------------------------------------------------------------------------
    def _up(self, what, level, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        check(lib.tda_upload(self._h, what, level, _dptr(a), a.size))

    def _stream_ptr(self, stream=None):
        st = self.stream if stream is None else stream
        return C.c_void_p(st) if st else C.c_void_p(0)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.tda_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- run ---------------------------------------------------------------------------------
    def init(self, theta0):
        theta0 = np.ascontiguousarray(np.atleast_2d(theta0), dtype=np.float64)
        if theta0.shape != (self.C, self.d):
            raise ValueError("initial parameters must have shape (n_chains, d)")
        self._up(L.TDA_UP_INIT_THETA, 0, theta0)
        check(lib.tda_engine_init(self._h, self._stream_ptr()))
        self.iterations_done = 0

    def run(self, iterations, stream=None, record=True):
        """Advances every chain by `iterations` finest-level iterations; asynchronous on the CUDA
        stream (raw handle) given here or at construction.  record=False: burn-in, nothing is stored
        and the history position does not move."""
        fn = lib.tda_engine_run if record else lib.tda_engine_burn
        check(fn(self._h, int(iterations), self._stream_ptr(stream)))
        self.iterations_done += int(iterations)

    def sync(self, stream=None):
        check(lib.tda_engine_sync(self._h, self._stream_ptr(stream)))
        flags = self.error_flags()
        if flags and flags != getattr(self, "_warned_flags", 0):
            self._warned_flags = flags
            warnings.warn("tinyda_b200: a non-positive Cholesky pivot of a per-chain covariance (error model / "
                          "Adaptive Metropolis) was clamped on the device; the affected chains continue with a "
                          "regularised factor", RuntimeWarning)

    def error_flags(self):
        out = np.zeros(1, dtype=np.int64)
        check(lib.tda_get(self._h, L.TDA_G_ERROR_FLAGS, 0, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return int(out[0])

    # ---- compacted history of the finest level -------------------------------------------------
    def compact_begin(self, rec0, nrec, first_is_full, fields=("theta", "stats"), slot=0, stream=None):
        """Enqueues the device-side compaction of finest-level records [rec0, rec0+nrec) to the accepted
        ones (a rejected step repeats the previous Link, chain.py:116, :434)."""
        bits = {"theta": L.TDA_STORE_THETA, "stats": L.TDA_STORE_STATS, "output": L.TDA_STORE_OUTPUT, "qoi": L.TDA_STORE_QOI}
        mask = 0
        for f in fields:
            mask |= bits[f]
        check(lib.tda_compact_begin(self._h, self.Ln - 1, int(rec0), int(nrec), int(bool(first_is_full)), mask, int(slot),
                                    self._stream_ptr(stream)))
        self._compact = getattr(self, "_compact", {})
        self._compact[int(slot)] = (int(nrec), tuple(fields))

    def compact_collect(self, slot=0, pinned=True):
        """Waits for the row count of `slot`, enqueues the device->host copies on the engine's copy stream
        and returns the CompactChunk they fill (valid after compact_sync())."""
        nrec, fields = self._compact[int(slot)]
        n = C.c_int64(0)
        check(lib.tda_compact_rows(self._h, int(slot), C.byref(n)))
        n = int(n.value)
        alloc = pinned_empty if pinned else (lambda shape, dt: np.empty(shape, dtype=dt))
        ch = CompactChunk(nrec)
        m = int(self.spec["levels"][self.Ln - 1]["model"]["m"])
        nq = int(self.spec["levels"][self.Ln - 1]["model"].get("n_qoi", 0))

        def get(field, shape, dt):
            out = alloc(shape, dt)
            nb = C.c_size_t(0)
            check(lib.tda_compact_fetch(self._h, int(slot), field, out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(nb)))
            return out

        ch.accept = get(L.TDA_F_ACCEPT, (nrec, self.C), np.uint8)
        ch.offsets = get(L.TDA_CF_OFFSETS, (self.C + 1,), np.int64)
        if "theta" in fields:
            ch.theta = get(L.TDA_F_THETA, (n, self.d), self.dtype)
        if "stats" in fields:
            ch.prior = get(L.TDA_F_PRIOR, (n,), self.dtype)
            ch.like = get(L.TDA_F_LIKE, (n,), self.dtype)
        if "output" in fields:
            ch.output = get(L.TDA_F_OUTPUT, (n, m), self.dtype)
        if "qoi" in fields:
            ch.qoi = get(L.TDA_F_QOI, (n, nq), self.dtype)
        return ch

    def compact_sync(self):
        check(lib.tda_compact_sync(self._h))

    # ---- shared-archive DREAM over several GPUs: peer-memory exchange inside the kernel --------------------
    def peer_export(self):
        n = C.c_size_t(0)
        check(lib.tda_peer_export(self._h, None, 0, C.byref(n)))
        blob = np.zeros(n.value, dtype=np.uint8)
        check(lib.tda_peer_export(self._h, blob.ctypes.data_as(C.c_void_p), blob.nbytes, C.byref(n)))
        return blob

    def peer_import(self, n_ranks, my_rank, blobs):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8)
        check(lib.tda_peer_import(self._h, int(n_ranks), int(my_rank), blobs.ctypes.data_as(C.c_void_p), blobs.nbytes))
        self.peers_connected = True

    def ess_sums(self, level=None, rec0=0, nrec=None, n_lag=0, stream=None):
        """Rank-normalised split-chain sums of the level's recorded parameters, computed on the device
        (tda_ess_sums): (sums [d, n_lag + 4], folded [d, 4]); feed them to diagnostics.ess_rhat_from_sums,
        after parallel.allreduce_sums when the chains are sharded over ranks."""
        level = self.Ln - 1 if level is None else level
        if nrec is None:
            nrec = int(self.n_records()[level]) - rec0
        nh = nrec // 2
        if n_lag <= 0 or n_lag > nh:
            n_lag = nh
        sums = np.zeros((self.d, n_lag + 4))
        folded = np.zeros((self.d, 4))
        check(lib.tda_ess_sums(self._h, int(level), int(rec0), int(nrec), int(n_lag), _dptr(sums), _dptr(folded), self._stream_ptr(stream)))
        return sums, folded

    def select_kernel(self, which):
        """"auto" | "generic" | "tc" (tcgen05, 3xTF32) | "tc16" (tcgen05, fp16 split + RNG warps) |
        "tcr" (tcgen05, whitened state + output recursion) | "reg" (register-resident single-level kernel, d <= 8) |
        "dreamw" (warp-per-chain DREAM(Z) / DREAM, d <= 32) | "mldaw" (warp-per-chain MH / DA / MLDA on the 1-D Poisson
        model, state-independent error model).""" 
        check(lib.tda_select_kernel(self._h, {"auto": 0, "generic": 1, "tc": 2, "tc16": 3, "reg": 4, "tcr": 5, "dreamw": 6, "mldaw": 7}[which]))

    def kernel(self):
        """Name of the kernel `run` launches for this configuration."""
        out = np.zeros(1, dtype=np.int64)
        check(lib.tda_get(self._h, L.TDA_G_KERNEL, 0, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return {1: "generic", 2: "tc", 3: "tc16", 4: "reg", 5: "tcr", 6: "dreamw", 7: "mldaw"}[int(out[0])]

    def set_z_round(self, on=True):
        """Philox normals on the fp16 grid (the "z16" stream the tc16 kernel consumes) also for the
        generic / tc kernels -- lets the kernels be compared on identical streams."""
        v = np.array([1.0 if on else 0.0])
        check(lib.tda_set(self._h, L.TDA_G_ZROUND, 0, v.ctypes.data_as(C.c_void_p), v.nbytes))

    def save_state(self):
        """Checkpoint: the whole sampler state (no history) as a bytes-like NumPy array."""
        n = C.c_size_t(0)
        check(lib.tda_state_size(self._h, C.byref(n)))
        blob = np.empty(n.value, dtype=np.uint8)
        check(lib.tda_state_save(self._h, blob.ctypes.data_as(C.c_void_p), blob.nbytes))
        self._saved_iterations = self.iterations_done
        return blob

    def load_state(self, blob, iterations_done=0):
        """Restores a checkpoint made by an engine with the same configuration; `run` continues the
        chains exactly where the saved engine stood."""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        check(lib.tda_state_load(self._h, blob.ctypes.data_as(C.c_void_p), blob.nbytes))
        self.iterations_done = int(iterations_done)

    def history_reset(self):
        check(lib.tda_history_reset(self._h))

    def n_records(self):
        out = np.zeros(self.Ln, dtype=np.int64)
        check(lib.tda_get(self._h, L.TDA_G_NRECORDS, 0, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    # ---- fetch ---------------------------------------------------------------------------------
    def fetch(self, level, field, rec0=0, nrec=None, out=None, stream=None, sync=True):
        """History of one level in the device layout: theta [nrec, d, C], prior/like [nrec, C],
        output [nrec, m, C], accept [nrec, C] (uint8).  With sync=False the copy is only enqueued
        on `stream` (pinned `out` required for it to be asynchronous)."""
        if nrec is None:
            nrec = int(self.n_records()[level]) - rec0
        m = int(self.spec["levels"][level]["model"]["m"])
        fid, shape, dt = {
            "theta": (L.TDA_F_THETA, (nrec, self.d, self.C), self.dtype),
            "prior": (L.TDA_F_PRIOR, (nrec, self.C), self.dtype),
            "like": (L.TDA_F_LIKE, (nrec, self.C), self.dtype),
            "output": (L.TDA_F_OUTPUT, (nrec, m, self.C), self.dtype),
            "accept": (L.TDA_F_ACCEPT, (nrec, self.C), np.dtype(np.uint8)),
            "qoi": (L.TDA_F_QOI, (nrec, int(self.spec["levels"][level]["model"].get("n_qoi", 0)), self.C), self.dtype),
        }[field]
        if out is None:
            out = np.empty(shape, dtype=dt)
        nb = C.c_size_t(0)
        check(lib.tda_fetch(self._h, level, fid, int(rec0), int(nrec), out.ctypes.data_as(C.c_void_p),
                            out.nbytes, C.byref(nb), self._stream_ptr(stream)))
        if sync:
            self.sync(stream)
        return out

    def get(self, what, level=0):
        Cn, d = self.C, self.d
        if what == "scaling":
            out = np.zeros(Cn)
            code = L.TDA_G_SCALING
        elif what == "accept_counts":
            out = np.zeros((self.Ln, Cn), dtype=np.int64)
            code = L.TDA_G_ACCEPT_COUNTS
        elif what == "cursors":
            out = np.zeros((2, Cn), dtype=np.int64)
            code = L.TDA_G_CURSORS
        elif what == "am_sigma":
            out = np.zeros((Cn, d, d))
            code = L.TDA_G_AM_SIGMA
        elif what == "am_mu":
            out = np.zeros((Cn, d))
            code = L.TDA_G_AM_MU
        elif what == "theta":
            out = np.zeros((Cn, d))
            code = L.TDA_G_THETA
        elif what == "moments":
            out = np.zeros((2, d, Cn))
            code = L.TDA_G_MOMENTS
        else:
            raise KeyError(what)
        check(lib.tda_get(self._h, code, level, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def set_scaling(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        check(lib.tda_set(self._h, L.TDA_G_SCALING, 0, s.ctypes.data_as(C.c_void_p), s.nbytes))

    def upload_am_factors(self, T):
        self._up(L.TDA_UP_AM_FACTORS, 0, T)

    def device_buffer(self, which, level=0):
        ptr, nb = C.c_void_p(), C.c_size_t()
        code = {"dream_archive": L.TDA_BUF_DREAM_ARCHIVE, "hist_theta": L.TDA_BUF_HIST_THETA}[which]
        check(lib.tda_device_buffer(self._h, code, level, C.byref(ptr), C.byref(nb)))
        return ptr.value, nb.value

    def dream_slots(self):
        s = C.c_int64()
        check(lib.tda_dream_slots(self._h, C.byref(s)))
        return s.value

    def fill_streams(self, nz, nu):
        z = np.zeros((self.C, nz))
        u = np.zeros((self.C, nu))
        check(lib.tda_fill_streams(self._h, _dptr(z), nz, _dptr(u), nu))
        return z, u


def launch_count():
    return int(lib.tda_launch_count())
