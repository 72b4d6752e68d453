"""SciPy-flavoured front ends primme_b200.api.eigsh_csr / svds_csr with the option names of the reference's
Python eigsh / svds (Python/primme.pyx:284-600,1074-1400): `which`, `sigma`, `v0`, `ncv`, `maxiter`, `lock`,
`method`, against dense LAPACK.  On CPU they drive the host-check build (product host code over the CPU
restatement of the kernels: test infrastructure); on the GPU the product."""
import numpy as np
import pytest

import harness as H
import svds_harness as S
from primme_b200 import api, matrices as M

CSR = M.laplacian_nd((7, 11, 13))
N = len(CSR[0]) - 1
SPEC = np.linalg.eigvalsh(M.csr_matvec(*CSR, np.eye(N)))


def check_eigsh(lib):
    for which, sigma, want in (("SA", None, SPEC[:4]), ("LA", None, SPEC[::-1][:4]),
                               ("SM", 1.3, SPEC[np.argsort(np.abs(SPEC - 1.3))][:4]),
                               (1.3, None, SPEC[np.argsort(np.abs(SPEC - 1.3))][:4]),
                               ("LM", 6.0, SPEC[np.argsort(-np.abs(SPEC - 6.0))][:4]),
                               ("CGT", 1.3, SPEC[SPEC >= 1.3][:4]), ("CLT", 1.3, SPEC[SPEC <= 1.3][::-1][:4])):
        w, X, st = api.eigsh_csr(*CSR, k=4, which=which, sigma=sigma, tol=1e-9, maxBlockSize=2, lib=lib, return_stats=True)
        if which in ("CGT", "CLT"):
            # one-sided targets may skip a value next to the shift (the reference returns the same set):
            # eigenvalues of A on the right side of sigma
            assert all(np.abs(SPEC - e).min() < 1e-7 for e in w)
            assert np.all(w >= sigma - 1e-7) if which == "CGT" else np.all(w <= sigma + 1e-7)
        else:
            assert np.allclose(np.sort(w), np.sort(want), atol=1e-7), (which, sigma, w, want)
        assert np.linalg.norm(M.csr_matvec(*CSR, X) - X * w, axis=0).max() < 1e-9 * 12 * 1.1
        assert st["numMatvecs"] > 0
    # projection, ncv, lock, initial guesses, values only
    w = api.eigsh_csr(*CSR, k=3, which="SM", sigma=1.3, tol=1e-8, projection="refined", ncv=30, lib=lib, return_eigenvectors=False)
    assert np.allclose(np.sort(w), np.sort(SPEC[np.argsort(np.abs(SPEC - 1.3))][:3]), atol=1e-6)
    w0, X0 = api.eigsh_csr(*CSR, k=2, which="SA", tol=1e-10, lib=lib)
    w1, X1, st = api.eigsh_csr(*CSR, k=2, which="SA", tol=1e-10, v0=X0, lock=True, lib=lib, return_stats=True)
    assert np.allclose(w1, w0) and st["numOuterIterations"] <= 10       # started from the solution
    with pytest.raises(ValueError):
        api.eigsh_csr(*CSR, k=2, which=1.0, sigma=2.0, lib=lib)
    # iteration budget: unconverged pairs raise unless asked not to
    with pytest.raises(RuntimeError):
        api.eigsh_csr(*CSR, k=4, which="SM", sigma=1.3, tol=1e-12, maxiter=5, lib=lib)
    w = api.eigsh_csr(*CSR, k=4, which="SM", sigma=1.3, tol=1e-12, maxiter=5, lib=lib, raise_for_unconverged=False,
                      return_eigenvectors=False)
    assert len(w) < 4


def check_svds(lib):
    m, n = 400, 90
    csr = S.random_rect(m, n, 6, 8)
    A = S.dense(csr, (m, n))
    sv = np.linalg.svd(A, compute_uv=False)
    for which, method, want, tol in (("LM", "hybrid", sv[:3], 1e-12), ("LM", "normalequations", sv[:3], 1e-9),
                                     ("LM", "augmented", sv[:3], 1e-10), ("SM", "hybrid", sv[::-1][:3], 1e-11),
                                     (float(sv[40]) + 1e-3, "hybrid", sv[np.argsort(np.abs(sv - sv[40] - 1e-3))][:3], 1e-10)):
        U, s, Vt, st = api.svds_csr(*csr, (m, n), k=3, which=which, method=method, tol=tol, lib=lib, return_stats=True,
                                    methodStage1=api.PRIMME_GD_Olsen_plusK, methodStage2=api.PRIMME_GD_Olsen_plusK)
        assert np.allclose(np.sort(s), np.sort(want), atol=1e-7 * sv[0]), (which, method, s, want)
        assert np.abs(Vt @ Vt.T - np.eye(3)).max() < 1e-7
        R = np.sqrt(np.linalg.norm(A @ Vt.T - U * s, axis=0) ** 2 + np.linalg.norm(A.T @ U - Vt.T * s, axis=0) ** 2)
        assert R.max() < max(tol, 1e-8 if method == "normalequations" else tol) * st["aNorm"] * 1.1 + 1e-12
    s = api.svds_csr(*csr, (m, n), k=2, tol=1e-8, lib=lib, return_singular_vectors=False)
    assert np.allclose(np.sort(s)[::-1], sv[:2], rtol=1e-7)


def test_eigsh_csr_host_logic():
    check_eigsh(H.lib_hostcheck())


def test_svds_csr_host_logic():
    check_svds(H.lib_hostcheck())


@pytest.mark.gpu
def test_eigsh_csr_gpu():
    check_eigsh(H.lib_product())


@pytest.mark.gpu
def test_svds_csr_gpu():
    check_svds(H.lib_product())
