"""The reference's own regression configurations on LUNDA.mtx (reference tests/tests/test_001 ..
test_005) and its check_solution acceptance test (tests/COMMON/ioandtest.c:71-155), restated."""
import os

import numpy as np

from primme_b200 import api as G, matrices as M

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = np.load(os.path.join(HERE, "golden", "lunda_fixture.npz"))
CSR = (FIX["indptr"], FIX["indices"], FIX["data"])
FNORM = float(FIX["fnorm"])  # the driver sets aNorm to the Frobenius norm (tests/COMMON/csr.c:241)

# name -> (numEvals, solve kwargs); from tests/tests/test_00N
LUNDA = {
    "001": (5, dict(eps=1e-12, maxBasisSize=140, minRestartSize=1, maxBlockSize=1, maxMatvecs=140,
                    target=G.primme_largest, locking=1, method=G.PRIMME_GD_Olsen_plusK)),
    "002": (30, dict(eps=1e-12, maxBasisSize=3, minRestartSize=1, maxBlockSize=1, maxOuterIterations=7800,
                     target=G.primme_largest, locking=1, maxPrevRetain=1, method=G.PRIMME_GD_Olsen_plusK)),
    "003": (50, dict(eps=1e-12, maxOuterIterations=7500, target=G.primme_largest, method=G.PRIMME_GD_Olsen_plusK)),
    "004": (50, dict(eps=1e-12, maxOuterIterations=7500, target=G.primme_closest_abs, targetShifts=[0.0],
                     method=G.PRIMME_GD_Olsen_plusK)),
    "005": (50, dict(eps=1e-12, maxOuterIterations=7500, target=G.primme_closest_abs, targetShifts=[0.0],
                     jacobi=True, method=G.PRIMME_GD_Olsen_plusK)),
}


def check_solution(name, r, eps, aNorm):
    """reference check_solution: orthonormality 1e-7, Rayleigh quotient, residual honesty,
    projected residual, and every returned vector inside the span of the stored golden vectors"""
    X = r["evecs"]
    evals, rnorms = r["evals"], r["rnorms"]
    k = r["initSize"]
    Xg = FIX["sol_" + name]  # golden vectors, one per row
    delta = aNorm
    for i in range(1, k):
        delta = min(delta, abs(evals[i] - evals[i - 1]))
    AX = M.csr_matvec(*CSR, X[:, :k])
    for i in range(k):
        h = X[:, : i + 1].T @ X[:, i]
        assert np.sqrt((h[:i] ** 2).sum()) <= 1e-7
        assert abs(np.sqrt(h[i]) - 1) <= 1e-7
        rq = X[:, i] @ AX[:, i]
        assert abs(evals[i] - rq) <= max(rnorms[i], aNorm * eps)
        res = AX[:, i] - evals[i] * X[:, i]
        rn0 = np.linalg.norm(res)
        assert abs(rnorms[i] - rn0) <= max(2 * rn0, 10 * max(aNorm, abs(evals[i])) * 2.2e-16)
        res = res - X[:, :k] @ (X[:, :k].T @ res)
        assert np.linalg.norm(res) <= eps * aNorm * 2
        prod = ((Xg @ X[:, i]) ** 2).sum()
        bound = aNorm * eps / delta
        s2 = np.sqrt(2.0)
        assert not ((s2 * prod + 1.0) / (s2 * bound + 1.0) < (s2 * prod - 1.0) / (1.0 - s2 * bound)), (name, i, prod)


def run(which, name):
    import harness as H
    k, kw = LUNDA[name]
    r = H.solve(which, CSR, k, aNorm=FNORM, **kw)
    assert r["ret"] == 0, (name, r["ret"])
    assert r["initSize"] == k
    check_solution(name, r, kw["eps"], FNORM)
    return r
