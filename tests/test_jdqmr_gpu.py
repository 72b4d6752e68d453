"""JDQMR family on the GPU: the product (sm_100a kernels) against the unmodified reference on the
same problems -- eigenvalues to 1e-10, residuals below tolerance, iteration / matvec counts within
10 % (the kernels sum in a different order than the CPU BLAS, the decisions are the same)."""
import numpy as np
import pytest

import harness as H
import test_jdqmr_cpu as T
from primme_b200 import matrices as M

pytestmark = pytest.mark.gpu

CASES = ["lap3d_jdqmr", "lap3d_jdqmr_etol_jacobi", "lap3d_jdqmr_locking", "lap3d_etol_largest", "lap3d_jdqmr_block2",
         "lap3d_min_time", "lap3d_closest_abs", "aniso_jdqmr_jacobi", "aniso_etol_locking", "lap2d_jdqmr_noprec",
         # right projectors: PRIMME_JDQR without a preconditioner, skew-X projector with the Jacobi preconditioner
         "aniso_jdqr", "aniso_jdqr_block3_locking", "aniso_jdqmr_skewX_jacobi"]


@pytest.mark.parametrize("name", CASES)
def test_jdqmr_product_matches_reference(name):
    mat, k, kw = T.EXACT[name]
    csr = mat()
    ref = H.solve("reference", csr, k, **kw)
    got = H.solve("product", csr, k, **kw)
    assert got["ret"] == 0 and got["initSize"] == k and got["launches"] > 0
    scale = max(1.0, np.abs(ref["evals"]).max())
    assert np.abs(got["evals"] - ref["evals"]).max() <= 1e-9 * scale
    X = got["evecs"]
    assert np.abs(X.T @ X - np.eye(k)).max() < 1e-8
    R = M.csr_matvec(*csr, X) - X * got["evals"]
    anorm = np.abs(np.asarray(csr[2])).sum() / (len(csr[0]) - 1) * 4
    assert np.linalg.norm(R, axis=0).max() <= 10 * kw["eps"] * anorm
    for key in ("numOuterIterations", "numMatvecs"):
        assert abs(got["stats"][key] - ref["stats"][key]) <= max(3, 0.10 * ref["stats"][key]), (got["stats"], ref["stats"])

