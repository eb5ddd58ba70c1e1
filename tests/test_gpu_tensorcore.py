"""tcgen05 tensor-core path: self-test GEMM (descriptor / TMEM conventions, 3xTF32 accuracy)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(A, B, a_in_tmem, split):
    from tinyda_b200._lib import lib, check
    N = B.shape[1]
    D = np.zeros((128, N), dtype=np.float32)
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    check(lib.tda_tc_gemm_selftest(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), N,
                                   D.ctypes.data_as(C.c_void_p), a_in_tmem, split))
    return D


@pytest.mark.parametrize("a_in_tmem", [1, 0])
@pytest.mark.parametrize("N", [64, 128, 8, 256])
def test_tf32x3_gemm_matches_fp64(a_in_tmem, N):
    rng = np.random.default_rng(N + a_in_tmem)
    A = rng.standard_normal((128, 64)).astype(np.float32)
    B = (rng.standard_normal((64, N)) / 8).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    D = _gemm(A, B, a_in_tmem, 1)
    err = np.abs(D - ref).max()
    scale = np.abs(ref).max()
    assert err < 2e-6 * scale + 1e-6, (err, scale)
    # a single TF32 pass on B is ~1e-3 relative: the split matters
    D1 = _gemm(A, B, a_in_tmem, 0)
    err1 = np.abs(D1 - ref).max()
    assert err1 < 5e-3 * scale and err1 > 5 * err
