"""The skew-Q projector with a preconditioner (PRIMME_JDQR + applyPreconditioner) on the GPU.  Written after the
round's GPU budget had ended: the host logic is verified on the CPU host-check build (tests/test_jdqmr_cpu.py: counts
identical to the reference with its two crashing lines fixed), the kernels it calls are the ortho-sweep entry points
the other JDQR cases already run on the GPU (tests/test_jdqmr_gpu.py).  Kept in a file of its own, collected last."""
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
