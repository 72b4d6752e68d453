"""zprimme (complex Hermitian) host control code on the CPU: the typed host sources compiled with -DPB_COMPLEX
over the complex oracle kernels (oracle/kernels_ref_z.c; test infrastructure) against the UNMODIFIED reference's
zprimme (oracle/_ref/libprimme_ref.so) on the same matrices, callbacks and parameters: eigenvalues to 1e-10
relative, residual norms below eps*|A|, and IDENTICAL outer-iteration / restart / matvec counts (the projected
problems go through the same zhegvx/zheevx calls with the same arguments, reference blaslapack.c:1058-1235)."""
import numpy as np
import pytest

import harness as H
from primme_b200 import api, matrices as M

CASES = {
    # name: (n, kwargs of H.zsolve)
    "gd_olsen_b2_smallest": (3000, dict(numEvals=6, method=api.PRIMME_GD_Olsen_plusK, maxBlockSize=2, maxBasisSize=30, eps=1e-10)),
    "gd_olsen_b1_cgs_largest_locking": (2500, dict(numEvals=5, target=api.primme_largest, method=api.PRIMME_GD_Olsen_plusK,
                                                     maxBlockSize=1, maxBasisSize=24, locking=1, eps=1e-10)),
    "gd_olsen_b4_largest": (4000, dict(numEvals=8, target=api.primme_largest, method=api.PRIMME_GD_Olsen_plusK,
                                         maxBlockSize=4, maxBasisSize=40, eps=1e-9)),
    # config C3's shape (interior pairs near sigma = 0.5, JDQMR_ETol, Jacobi, locking, block 1 => CGS ortho)
    "c3_jdqmr_etol_interior_jacobi": (4000, dict(numEvals=8, target=api.primme_closest_abs, targetShifts=[0.5],
                                                  method=api.PRIMME_JDQMR_ETol, jacobi=True, eps=1e-10)),
    "jdqmr_smallest_b2": (3000, dict(numEvals=4, method=api.PRIMME_JDQMR, maxBlockSize=2, eps=1e-9)),
    "gd_plusk_closest_geq_locking": (1500, dict(numEvals=2, target=api.primme_closest_geq, targetShifts=[0.3],
                                                  method=api.PRIMME_GD_plusK, locking=1, eps=1e-8)),
}
# interior targets amplify last-bit differences between the reference's BLAS calls and the oracle's plain loops
# (same behaviour as the real solver): counts within 10 %, everything else identical
CLOSE = {"c3_jdqmr_etol_interior_jacobi", "gd_plusk_closest_geq_locking"}


def _dense_evals(csr, n):
    A = np.zeros((n, n), dtype=complex)
    ip, ix, da = csr
    rows = np.repeat(np.arange(n), np.diff(ip))
    A[rows, ix] = da
    return np.linalg.eigvalsh(A), A


@pytest.mark.parametrize("name", sorted(CASES))
def test_zprimme_hostcheck_matches_reference(name):
    if not H.have_reference():
        pytest.skip("reference build not available")
    n, kw = CASES[name]
    csr = M.hermitian_c3(n, **M.C3_MATRIX)
    ref = H.zsolve("reference", csr, **kw)
    got = H.zsolve("hostcheck", csr, **kw)
    assert ref["ret"] == 0 and got["ret"] == 0
    assert np.allclose(got["evals"], ref["evals"], rtol=1e-10, atol=1e-12)
    for key in ("numOuterIterations", "numRestarts", "numMatvecs"):
        a, b = got["stats"][key], ref["stats"][key]
        if name in CLOSE:
            assert abs(a - b) <= 0.10 * b + 3, (name, key, a, b)
        else:
            assert a == b, (name, key, a, b)
    # independent check of what came back: residuals and orthonormality in complex arithmetic
    ev, A = _dense_evals(csr, n)
    X = got["evecs"]
    R = A @ X - X * got["evals"]
    eps = kw.get("eps", 1e-10)
    assert np.linalg.norm(R, axis=0).max() < 10 * eps * max(1.0, np.abs(ev).max())
    assert np.abs(X.conj().T @ X - np.eye(X.shape[1])).max() < 1e-8


def test_complex_refined_extraction_is_refused():
    csr = M.hermitian_c3(500, **M.C3_MATRIX)

    def tweak(p):
        p.projectionParams.projection = api.primme_proj_refined

    r = H.zsolve("hostcheck", csr, 2, target=api.primme_closest_abs, targetShifts=[0.5], tweak=tweak)
    assert r["ret"] == api.PRIMME_FUNCTION_UNAVAILABLE


@pytest.mark.parametrize("kw,exact", [
    (dict(numEvals=5, method=api.PRIMME_JDQMR, jacobi=True, eps=1e-9, locking=1, projectors=(1, 1, 1, 0, 1, 0)), True),
    (dict(numEvals=5, method=api.PRIMME_JDQMR, jacobi=True, eps=1e-9, locking=0, projectors=(0, 1, 1, 0, 1, 0)), True),
    (dict(numEvals=5, method=api.PRIMME_JDQR, jacobi=True, eps=1e-9), False),
    (dict(numEvals=4, target=api.primme_largest, method=api.PRIMME_JDQR, jacobi=True, maxBlockSize=2, eps=1e-9), False),
], ids=["skewQ_locking", "skewQ_soft", "jdqr_jacobi_smallest", "jdqr_jacobi_largest_b2"])
def test_zprimme_skewQ_with_preconditioner_matches_fixed_reference(kw, exact):
    """complex twin of tests/test_jdqmr_cpu.py::test_skewQ_with_preconditioner_*: the skew-Q projector with the Jacobi
    preconditioner (K^{-1}Q, zhetrf / zhetrs of Q^H K^{-1} Q) against the reference with the two one-line fixes of
    oracle/Makefile (the unmodified build crashes in this configuration).  With the skew-Q projector alone the counts
    are identical; the PRIMME_JDQR preset adds the skew-X projector, whose oblique projection amplifies the last-bit
    differences between the reference's zgemm-based dots and the plain loops of the oracle kernels on this matrix
    (the same pair of solvers agrees exactly at n = 800, 1500, 3000): counts within 15 % there."""
    n = 2000
    csr = M.hermitian_c3(n, **M.C3_MATRIX)
    ref = H.zsolve("reference_skewq", csr, **kw)
    got = H.zsolve("hostcheck", csr, **kw)
    assert ref["ret"] == 0 and got["ret"] == 0
    assert np.allclose(got["evals"], ref["evals"], rtol=1e-10, atol=1e-12)
    for key in ("numOuterIterations", "numRestarts", "numMatvecs"):
        a, b = got["stats"][key], ref["stats"][key]
        if exact:
            assert a == b, (key, a, b)
        else:
            assert abs(a - b) <= 0.15 * b + 3, (key, a, b)
    ev, A = _dense_evals(csr, n)
    X = got["evecs"]
    R = A @ X - X * got["evals"]
    assert np.linalg.norm(R, axis=0).max() < 10 * kw["eps"] * max(1.0, np.abs(ev).max())
