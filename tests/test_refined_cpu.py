"""Refined and harmonic extraction (primme_proj_refined: the QR factorisation of (A - tau I) V carried next to V and W,
coefficient vectors from the SVD of R; reference src/eigs/update_W.c:69-113, solve_projection.c:541-628,
842-985, restart.c:1837-2160) -- the host logic over the CPU restatement of the kernels against the
UNMODIFIED reference on the same matrices and parameters.

Interior eigenproblems amplify rounding differences (the reference orthogonalises the new columns of Q in
one block, the product in 8-column chunks; the traces below show the Ritz values agreeing to 1e-14 while the
residual norms near the threshold already differ), so over hundreds of iterations the counts drift; the
first outer iterations -- several restarts of the factorisation -- must agree to rounding."""
import ctypes as C

import numpy as np
import pytest

import harness as H
from primme_b200 import api, matrices as M

R = api.primme_proj_refined
CASES = {
    # name: (numEvals, parameters)
    "closest_abs": (4, dict(target=api.primme_closest_abs, targetShifts=[1.3], eps=1e-8)),
    "closest_abs_block3": (5, dict(target=api.primme_closest_abs, targetShifts=[1.3], eps=1e-8, maxBlockSize=3)),
    "closest_geq_locking": (4, dict(target=api.primme_closest_geq, targetShifts=[1.3], eps=1e-8, locking=1)),
    "closest_leq_locking_block2": (4, dict(target=api.primme_closest_leq, targetShifts=[1.3], eps=1e-8, locking=1, maxBlockSize=2)),
    "closest_geq_three_shifts": (3, dict(target=api.primme_closest_geq, targetShifts=[0.5, 1.0, 1.3], eps=1e-8, locking=1)),
    "jdqmr_closest_abs_jacobi": (3, dict(target=api.primme_closest_abs, targetShifts=[1.3], eps=1e-8, method=api.PRIMME_JDQMR, jacobi=True)),
    "jdqmr_closest_geq_three_shifts": (3, dict(target=api.primme_closest_geq, targetShifts=[0.5, 1.0, 1.3], eps=1e-8,
                                               method=api.PRIMME_JDQMR, locking=1)),
    "largest_abs": (3, dict(target=api.primme_largest_abs, targetShifts=[1.3], eps=1e-8)),
}
MATRIX = lambda: M.laplacian_nd((7, 11, 13))     # n = 1001, no repeated eigenvalues


@pytest.mark.parametrize("case", sorted(CASES))
def test_refined_hostcheck_matches_reference(case):
    k, kw = CASES[case]
    csr = MATRIX()
    ref = H.solve("reference", csr, k, projection=R, **kw)
    got = H.solve("hostcheck", csr, k, projection=R, **kw)
    assert ref["ret"] == 0 and got["ret"] == 0 and got["initSize"] == k
    assert np.abs(np.sort(got["evals"]) - np.sort(ref["evals"])).max() <= 1e-8 * 12 * 10
    X = got["evecs"]
    res = np.linalg.norm(M.csr_matvec(*csr, X) - X * got["evals"], axis=0)
    aNorm = max(abs(ref["stats"]["estimateLargestSVal"]), 1.0)
    assert res.max() < 1e-8 * aNorm * 1.1
    assert np.abs(X.T @ X - np.eye(k)).max() < 1e-8
    for key in ("numOuterIterations", "numMatvecs"):
        assert abs(got["stats"][key] - ref["stats"][key]) <= 0.3 * ref["stats"][key], (got["stats"][key], ref["stats"][key])


HARMONIC = {
    "closest_abs": (3, dict(target=api.primme_closest_abs, targetShifts=[1.3], eps=1e-8)),
    "closest_geq_locking": (3, dict(target=api.primme_closest_geq, targetShifts=[1.3], eps=1e-8, locking=1)),
    "closest_leq_block2": (3, dict(target=api.primme_closest_leq, targetShifts=[1.3], eps=1e-8, locking=1, maxBlockSize=2)),
}


@pytest.mark.parametrize("case", sorted(HARMONIC))
def test_harmonic_hostcheck_matches_reference(case):
    """primme_proj_harmonic (solve_projection.c:395-470, restart.c:2256-2326): harmonic Ritz pairs from
    (Q'V inv(R), Q'Q); Q, R and Q'V are rebuilt at every restart"""
    k, kw = HARMONIC[case]
    csr = MATRIX()
    ref = H.solve("reference", csr, k, projection=api.primme_proj_harmonic, **kw)
    got = H.solve("hostcheck", csr, k, projection=api.primme_proj_harmonic, **kw)
    assert ref["ret"] == 0 and got["ret"] == 0 and got["initSize"] == k
    assert np.abs(np.sort(got["evals"]) - np.sort(ref["evals"])).max() <= 1e-6
    X = got["evecs"]
    res = np.linalg.norm(M.csr_matvec(*csr, X) - X * got["evals"], axis=0)
    assert res.max() < 1e-8 * 12 * 1.1
    for key in ("numOuterIterations", "numMatvecs"):
        assert abs(got["stats"][key] - ref["stats"][key]) <= 0.3 * ref["stats"][key], (got["stats"][key], ref["stats"][key])


def test_harmonic_iteration_trace_equals_reference():
    k, kw = HARMONIC["closest_abs"]
    csr = MATRIX()
    a = trace("reference", csr, k, projection=api.primme_proj_harmonic, **kw)
    b = trace("hostcheck", csr, k, projection=api.primme_proj_harmonic, **kw)
    assert len(a) > 40 and len(b) > 40
    for x, y in zip(a[:40], b[:40]):
        assert x[0] == y[0] and x[1] == y[1]
        assert abs(x[2] - y[2]) <= 1e-9 * max(1.0, abs(x[2]))
        assert abs(x[3] - y[3]) <= 1e-5 * x[3]


MON = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p,
                  C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int), C.c_void_p,
                  C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int))


def trace(which, csr, k, **kw):
    """(basis size, block index, Ritz value, residual norm) of every outer iteration, from monitorFun"""
    events = []

    def mon(be, bs, bf, ib, blk, bn, nc, le, nl, lf, ln, ii, ls, msg, t, event, p, err):
        if event[0] == 0 and blk[0] > 0:
            e = np.ctypeslib.as_array(C.cast(be, C.POINTER(C.c_double)), shape=(bs[0],))
            r = np.ctypeslib.as_array(C.cast(bn, C.POINTER(C.c_double)), shape=(bs[0],))
            events.append((bs[0], ib[0], float(e[ib[0]]), float(r[ib[0]])))
        err[0] = 0

    cb = MON(mon)
    H.solve(which, csr, k, monitorFun=C.cast(cb, C.c_void_p).value, **kw)
    return events


@pytest.mark.parametrize("case", ["closest_abs", "closest_geq_three_shifts", "closest_abs_block3"])
def test_refined_iteration_trace_equals_reference(case):
    """the first 40 outer iterations span several restarts of the factorisation: same basis sizes, same
    block, Ritz values equal to 1e-9, residual norms to 1e-5"""
    k, kw = CASES[case]
    csr = MATRIX()
    a = trace("reference", csr, k, projection=R, **kw)
    b = trace("hostcheck", csr, k, projection=R, **kw)
    assert len(a) > 40 and len(b) > 40
    sizes = [e[0] for e in a[:40]]
    assert any(sizes[i + 1] < sizes[i] for i in range(39))          # restarts happened inside the window
    for x, y in zip(a[:40], b[:40]):
        assert x[0] == y[0] and x[1] == y[1]
        assert abs(x[2] - y[2]) <= 1e-9 * max(1.0, abs(x[2]))
        assert abs(x[3] - y[3]) <= 1e-5 * x[3]
