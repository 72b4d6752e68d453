"""The generator behind pb200_dlarnv (csrc/mv_utils.cu) restated with Python integers and checked against LAPACK's
dlarnv itself (the oracle twin of pb200_dlarnv calls dlarnv_): dlaruv is x_k = a^k x_0 mod 2^48 with
a = 33952834046453, value k = 2 x_k / 2^48 - 1, and the seed left behind is x_n in base 4096.  The GPU test
(test_kernels_gpu.py::test_dlarnv_on_device_is_lapack_dlarnv) compares the kernel with the same oracle bit for bit;
this one pins the arithmetic the kernel implements on a machine without a GPU."""
import ctypes as C

import numpy as np
import pytest

import harness as H

A = 33952834046453
MASK = (1 << 48) - 1


def lcg_stream(seed, count):
    x = (seed[0] << 36) | (seed[1] << 24) | (seed[2] << 12) | seed[3]
    out = np.empty(count)
    for k in range(count):
        x = (x * A) & MASK
        out[k] = 2.0 * (x / float(1 << 48)) - 1.0  # exact: 48 significant bits
    return out, ((x >> 36) & 4095, (x >> 24) & 4095, (x >> 12) & 4095, x & 4095)


def jump(seed, k):
    """the kernel's jump-ahead: x_k = a^k x_0 mod 2^48 by square-and-multiply"""
    x = (seed[0] << 36) | (seed[1] << 24) | (seed[2] << 12) | seed[3]
    return (x * pow(A, k, 1 << 48)) & MASK


@pytest.mark.parametrize("n,ncols,seed", [(1, 1, (0, 0, 0, 1)), (64, 2, (4095, 4095, 4095, 4095)), (129, 3, (7, 0, 11, 13)),
                                          (5000, 4, (0, 1, 2, 3)), (777, 8, (1234, 567, 89, 1011))])
def test_lcg_restatement_equals_lapack_dlarnv(n, ncols, seed):
    lib = H.lib_oracle_kernels()
    ld = n + 3
    X = np.zeros((ncols, ld))
    iseed = (C.c_longlong * 4)(*seed)
    lib.pb200_dlarnv.restype = C.c_int
    lib.pb200_dlarnv.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int64, C.c_int, C.c_void_p, C.c_int64]
    assert lib.pb200_dlarnv(None, iseed, n, ncols, X.ctypes.data, ld) == 0
    want, seed_after = lcg_stream(seed, n * ncols)
    assert np.array_equal(X[:, :n].ravel(), want)
    assert tuple(iseed) == seed_after
    # jump-ahead to the middle of the stream and to its end
    k = (n * ncols) // 2 + 1
    xk = jump(seed, k)
    assert 2.0 * (xk / float(1 << 48)) - 1.0 == want[k - 1]
    xn = jump(seed, n * ncols)
    assert ((xn >> 36) & 4095, (xn >> 24) & 4095, (xn >> 12) & 4095, xn & 4095) == seed_after
