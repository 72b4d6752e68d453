"""dprimme with the reference's HOST contract (include/primme_eigs.h:390): host evecs, host
matrixMatvec callback written by the user (here: Python), as tests/driver.c and examples/ use it.
The basis lives in HBM, every block visits the host for the callback, and the eigenvectors come
back into the caller's host array with the caller's leading dimension."""
import ctypes as C

import numpy as np
import pytest

import harness as H
import lunda_cases as L
from primme_b200 import api, matrices as M

pytestmark = pytest.mark.gpu

MV = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int),
                 C.c_void_p, C.POINTER(C.c_int))


@pytest.mark.parametrize("locking,ldpad", [(1, 0), (0, 0), (1, 5)])
def test_dprimme_host_callbacks(locking, ldpad):
    lib = H.lib_product()
    csr = L.CSR
    n = len(csr[0]) - 1
    k = 5
    calls = []

    def matvec(x, ldx, y, ldy, bs, primme, ierr):
        b = bs[0]
        X = np.ctypeslib.as_array(C.cast(x, C.POINTER(C.c_double)), shape=(b, ldx[0]))[:, :n]
        Y = np.ctypeslib.as_array(C.cast(y, C.POINTER(C.c_double)), shape=(b, ldy[0]))
        Y[:, :n] = M.csr_matvec(*csr, X.T).T
        calls.append(b)
        ierr[0] = 0

    cb = MV(matvec)
    p = api.new_params(lib, n, numEvals=k, target=api.primme_largest, eps=1e-12, maxBasisSize=40, maxBlockSize=2,
                       locking=locking, aNorm=L.FNORM)
    assert lib.primme_set_method(api.PRIMME_GD_Olsen_plusK, C.byref(p)) == 0
    p.matrixMatvec = C.cast(cb, C.c_void_p).value
    ld = n + ldpad
    if ldpad:
        p.ldevecs = ld
    evals, rn = np.zeros(k), np.zeros(k)
    evecs = np.full((k, ld), 7.0)
    rc = lib.dprimme(evals.ctypes.data, evecs.ctypes.data, rn.ctypes.data, C.byref(p))
    assert rc == 0 and p.initSize == k and sum(calls) == p.stats.numMatvecs
    X = evecs[:, :n].T
    assert np.abs(X.T @ X - np.eye(k)).max() < 1e-10
    R = M.csr_matvec(*csr, X) - X * evals
    assert np.all(np.linalg.norm(R, axis=0) <= 1e-12 * L.FNORM * 10)
    if ldpad:
        assert np.all(evecs[:, n:] == 7.0)  # the padding rows of the caller's array are untouched
