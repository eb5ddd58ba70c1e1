"""ctypes binding of the C ABI (include/tinyda_b200.h).

Loads the in-tree shared library ``tinyda_b200/libtinyda_b200.so`` (built by
``tinyda_b200/csrc/Makefile`` / ``__graft_entry__.build()``).  If the library is missing the
import of this module raises: there is no CPU fallback anywhere in the package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtinyda_b200.so")

TDA_ABI_VERSION = 3
TDA_MAX_LEVELS = 4
TDA_MAX_D = 64
TDA_F32, TDA_F64 = 0, 1
TDA_RNG_PHILOX, TDA_RNG_INJECTED = 0, 1
TDA_STORE_THETA, TDA_STORE_STATS, TDA_STORE_OUTPUT, TDA_STORE_ACCEPT, TDA_STORE_QOI = 1, 2, 4, 8, 16
(TDA_UP_PRIOR_MEAN, TDA_UP_PRIOR_LP, TDA_UP_PRIOR_PREC, TDA_UP_PROP_T, TDA_UP_MODEL_A, TDA_UP_MODEL_B,
 TDA_UP_LIK_DATA, TDA_UP_LIK_VAR, TDA_UP_LIK_PREC, TDA_UP_LIK_COV, TDA_UP_INIT_THETA, TDA_UP_STREAM_Z,
 TDA_UP_STREAM_U, TDA_UP_DREAM_ARCHIVE0, TDA_UP_AM_FACTORS, TDA_UP_PROP_S, TDA_UP_PROP_S2,
 TDA_UP_PROP_LAMBDA, TDA_UP_QOI_W, TDA_UP_QOI_B) = range(1, 21)
TDA_F_THETA, TDA_F_PRIOR, TDA_F_LIKE, TDA_F_OUTPUT, TDA_F_ACCEPT, TDA_F_QOI, TDA_CF_OFFSETS = 1, 2, 3, 4, 5, 6, 7
(TDA_G_SCALING, TDA_G_ACCEPT_COUNTS, TDA_G_CURSORS, TDA_G_AM_SIGMA, TDA_G_AM_MU, TDA_G_THETA,
 TDA_G_NRECORDS, TDA_G_MOMENTS, TDA_G_ZROUND, TDA_G_TC16_TIMELINE, TDA_G_KERNEL, TDA_G_ERROR_FLAGS) = range(1, 13)
TDA_BUF_DREAM_ARCHIVE, TDA_BUF_HIST_THETA = 1, 2

EXPORTS = [
    "tda_abi_version", "tda_last_error", "tda_engine_create", "tda_engine_destroy", "tda_upload",
    "tda_engine_init", "tda_engine_run", "tda_engine_sync", "tda_fetch", "tda_get", "tda_set",
    "tda_device_buffer", "tda_dream_slots", "tda_fill_streams", "tda_history_reset",
    "tda_select_kernel", "tda_launch_count", "tda_tc_gemm_selftest", "tda_tc16_gemm_selftest",
    "tda_state_size", "tda_state_save", "tda_state_load",
    "tda_engine_burn", "tda_compact_begin", "tda_compact_rows", "tda_compact_fetch", "tda_compact_sync",
    "tda_host_alloc", "tda_host_free", "tda_pool_trim", "tda_ess_sums", "tda_peer_export", "tda_peer_import",
]


class LevelConfig(C.Structure):
    _fields_ = [
        ("model_kind", C.c_int32), ("m", C.c_int32), ("n_grid", C.c_int32), ("lik_kind", C.c_int32),
        ("lik_var", C.c_double), ("model_scalars", C.c_double * 4),
        ("store", C.c_int32), ("n_qoi", C.c_int32), ("hist_capacity", C.c_int64),
    ]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("dtype", C.c_int32), ("n_levels", C.c_int32), ("d", C.c_int32),
        ("subchain", C.c_int32 * TDA_MAX_LEVELS), ("aem", C.c_int32), ("rng_mode", C.c_int32),
        ("randomize_subchain", C.c_int32), ("mtm_k", C.c_int32),
        ("mtm_include_current", C.c_int32), ("dream_sync_every", C.c_int32),
        ("seed", C.c_uint64), ("n_chains", C.c_int64), ("chain_offset", C.c_int64),
        ("n_chains_global", C.c_int64),
        ("prop_kind", C.c_int32), ("adaptive", C.c_int32), ("period", C.c_int32), ("am_t0", C.c_int32),
        ("scaling", C.c_double), ("gamma", C.c_double), ("alpha_star", C.c_double),
        ("am_sd", C.c_double), ("am_eps", C.c_double),
        ("am_device_refactor", C.c_int32), ("dream_M0", C.c_int32), ("dream_delta", C.c_int32),
        ("dream_nCR", C.c_int32), ("dream_b", C.c_double), ("dream_b_star", C.c_double),
        ("dream_capacity", C.c_int64), ("stream_z_len", C.c_int64), ("stream_u_len", C.c_int64),
        ("prior_logconst", C.c_double), ("level", LevelConfig * TDA_MAX_LEVELS),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "tinyda_b200: %s is missing. Build it with `make -C tinyda_b200/csrc` (or "
            "`python -c 'import __graft_entry__ as g; g.build()'`). The engine is CUDA-only; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    lib.tda_abi_version.restype = i32
    lib.tda_last_error.restype = C.c_char_p
    lib.tda_launch_count.restype = i64
    lib.tda_engine_create.argtypes = [C.POINTER(Config), i32, C.POINTER(vp)]
    lib.tda_engine_destroy.argtypes = [vp]
    lib.tda_upload.argtypes = [vp, i32, i32, C.POINTER(C.c_double), sz]
    lib.tda_engine_init.argtypes = [vp, vp]
    lib.tda_engine_run.argtypes = [vp, i64, vp]
    lib.tda_engine_sync.argtypes = [vp, vp]
    lib.tda_engine_burn.argtypes = [vp, i64, vp]
    lib.tda_compact_begin.argtypes = [vp, i32, i64, i64, i32, i32, i32, vp]
    lib.tda_compact_rows.argtypes = [vp, i32, C.POINTER(i64)]
    lib.tda_compact_fetch.argtypes = [vp, i32, i32, vp, sz, C.POINTER(sz)]
    lib.tda_compact_sync.argtypes = [vp]
    lib.tda_peer_export.argtypes = [vp, vp, sz, C.POINTER(sz)]
    lib.tda_peer_import.argtypes = [vp, i32, i32, vp, sz]
    lib.tda_ess_sums.argtypes = [vp, i32, i64, i64, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), vp]
    lib.tda_host_alloc.argtypes = [sz, C.POINTER(vp)]
    lib.tda_host_free.argtypes = [vp]
    lib.tda_fetch.argtypes = [vp, i32, i32, i64, i64, vp, sz, C.POINTER(sz), vp]
    lib.tda_get.argtypes = [vp, i32, i32, vp, sz]
    lib.tda_set.argtypes = [vp, i32, i32, vp, sz]
    lib.tda_device_buffer.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(sz)]
    lib.tda_dream_slots.argtypes = [vp, C.POINTER(i64)]
    lib.tda_fill_streams.argtypes = [vp, C.POINTER(C.c_double), i64, C.POINTER(C.c_double), i64]
    lib.tda_history_reset.argtypes = [vp]
    lib.tda_select_kernel.argtypes = [vp, i32]
    lib.tda_state_size.argtypes = [vp, C.POINTER(sz)]
    lib.tda_state_save.argtypes = [vp, vp, sz]
    lib.tda_state_load.argtypes = [vp, vp, sz]
    lib.tda_tc_gemm_selftest.argtypes = [vp, vp, i32, vp, i32, i32]
    lib.tda_tc16_gemm_selftest.argtypes = [vp, vp, i32, vp, i32]
    for name in EXPORTS:
        if getattr(lib, name).restype is C.c_int and name not in ("tda_abi_version",):
            getattr(lib, name).restype = i32
    if lib.tda_abi_version() != TDA_ABI_VERSION:
        raise ImportError("tinyda_b200: ABI version mismatch between _lib.py and the shared library")
    return lib


lib = _load()


class EngineError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise EngineError("tinyda_b200 engine error %d: %s" % (rc, lib.tda_last_error().decode()))
