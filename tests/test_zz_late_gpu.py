"""GPU cases written after the round's GPU budget had ended (collected last, in a file of their own):
* the skew-Q projector with a preconditioner (PRIMME_JDQR + applyPreconditioner): the host logic is verified on the
  CPU host-check build (tests/test_jdqmr_cpu.py: counts identical to the reference with its two crashing lines
  fixed), the kernels it calls are the ortho-sweep entry points the other JDQR cases already run on the GPU
  (tests/test_jdqmr_gpu.py);
* config C5's shape at n = 10^6 on one GPU, checked independently of the solver (bench.py runs the same solve at
  n = 10^7 on 1-8 GPUs and reports its residuals)."""
import numpy as np
import pytest

import harness as H
import test_jdqmr_cpu as T
from primme_b200 import matrices as M

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["aniso_jdqr_jacobi", "aniso_jdqmr_all_projectors_soft"])
def test_skewQ_with_preconditioner_product_matches_fixed_reference(name):
    """the skew-Q projector with a preconditioner on the GPU (K^{-1}Q next to the locked vectors, the overlaps and the
    update through the ortho-sweep kernels, M factorised on the host) against the reference with the two one-line
    fixes of oracle/Makefile (tests/test_jdqmr_cpu.py has the count-identical CPU cases and the crash of the
    unmodified build)"""
    mat, k, kw = T.SKEWQ[name]
    csr = mat()
    ref = H.solve("reference_skewq", csr, k, **kw)
    got = H.solve("product", csr, k, **kw)
    assert ref["ret"] == 0 and got["ret"] == 0 and got["initSize"] == k and got["launches"] > 0
    scale = max(1.0, np.abs(ref["evals"]).max())
    assert np.abs(got["evals"] - ref["evals"]).max() <= 1e-9 * scale
    X = got["evecs"]
    assert np.abs(X.T @ X - np.eye(k)).max() < 1e-8
    R = M.csr_matvec(*csr, X) - X * got["evals"]
    anorm = np.abs(np.asarray(csr[2])).sum() / (len(csr[0]) - 1) * 4
    assert np.linalg.norm(R, axis=0).max() <= 10 * kw["eps"] * anorm
    for key in ("numOuterIterations", "numMatvecs"):
        assert abs(got["stats"][key] - ref["stats"][key]) <= max(3, 0.10 * ref["stats"][key]), (got["stats"], ref["stats"])


def test_power_law_one_million_rows_c5_shape():
    """config C5's shape on one GPU at n = 10^6 (same generator as the n = 10^7 matrix of bench.py's `c5` block:
    power-law degrees, nnz ~ 1.5e7): 20 largest pairs, GD_Olsen_plusK, block 8, basis 64, eps 1e-8 -- the wide panel
    kernels (b = 8, m <= 64), the column-split restart kernel and the row-major SpMM with its long-row passes.  Checked
    independently of the solver: residual norms and orthonormality recomputed with numpy, the largest eigenvalues
    against scipy's Lanczos on the same matrix."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from primme_b200 import api
    n, k = 1000000, 20
    csr = M.power_law_rows(n, mean_degree=15.0, seed=7)
    got = H.solve("product", csr, k, target=api.primme_largest, method=api.PRIMME_GD_Olsen_plusK, maxBlockSize=8,
                  maxBasisSize=64, eps=1e-8)
    assert got["ret"] == 0 and got["initSize"] == k and got["launches"] > 0
    ip, ix, da = csr
    A = sp.csr_matrix((da, ix, ip), shape=(n, n))
    X = got["evecs"]
    assert np.abs(X.T @ X - np.eye(k)).max() < 1e-8
    R = A @ X - X * got["evals"]
    anorm = np.abs(got["evals"]).max()
    assert np.linalg.norm(R, axis=0).max() <= 10 * 1e-8 * anorm
    want = np.sort(spl.eigsh(A, k=6, which="LA", tol=1e-10, return_eigenvectors=False))[::-1]
    assert np.abs(np.sort(got["evals"])[::-1][:6] - want).max() <= 1e-7 * anorm
