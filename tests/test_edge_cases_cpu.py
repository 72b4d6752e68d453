"""Edge cases of the hot path's caller against the UNMODIFIED reference (same return codes, numbers of
returned pairs, eigenvalues and -- where rounding cannot reorder decisions -- the same counts): tiny
problems, the whole spectrum, a basis larger than the matrix (input error), empty matrices / zero rows,
a fully degenerate spectrum, exhausted iteration budgets, default tolerance."""
import numpy as np
import pytest

import harness as H
from primme_b200 import api, matrices as M


def diag_csr(d):
    n = len(d)
    return np.arange(n + 1, dtype=np.int64), np.arange(n, dtype=np.int32), np.array(d, dtype=float)


ZERO = (np.zeros(11, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0))
DZ = np.arange(1, 41, dtype=float)
DZ[5] = 0.0
CASES = {
    # name: (csr, numEvals, parameters, counts must be identical)
    "n1": (diag_csr([3.0]), 1, {}, True),
    "n2_both": (diag_csr([3.0, 1.0]), 2, {}, True),
    "n2_largest": (diag_csr([3.0, 1.0]), 1, dict(target=api.primme_largest), True),
    "whole_space": (M.laplacian_1d(5), 5, {}, True),
    "whole_space_locking": (M.laplacian_1d(5), 5, dict(locking=1), True),
    "basis_larger_than_n": (M.laplacian_1d(6), 3, dict(maxBlockSize=4, maxBasisSize=40), True),   # -26 in both
    "all_pairs_block2": (M.laplacian_1d(30), 30, dict(maxBlockSize=2), False),
    "all_but_one_locking_block3": (M.laplacian_1d(30), 29, dict(maxBlockSize=3, locking=1), False),
    "zero_matrix": (ZERO, 2, {}, True),
    "zero_row": (diag_csr(DZ), 3, {}, True),
    "identity": (diag_csr(np.ones(20)), 3, {}, True),
    "matvec_budget": (M.laplacian_1d(500), 4, dict(maxMatvecs=50), True),
    "iteration_budget_locking": (M.laplacian_1d(500), 4, dict(maxOuterIterations=20, locking=1), True),
    "default_tolerance": (M.laplacian_1d(60), 2, dict(eps=0.0), True),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_edge_case_matches_reference(case):
    csr, k, kw, exact_counts = CASES[case]
    ref = H.solve("reference", csr, k, **kw)
    got = H.solve("hostcheck", csr, k, **kw)
    assert got["ret"] == ref["ret"] and got["initSize"] == ref["initSize"]
    assert np.allclose(np.sort(got["evals"]), np.sort(ref["evals"]), rtol=1e-9, atol=1e-12)
    if ref["ret"] == 0 and exact_counts:
        assert np.allclose(got["rnorms"], ref["rnorms"], rtol=0.5, atol=1e-12)
    keys = ("numOuterIterations", "numRestarts", "numMatvecs")
    if exact_counts:
        assert {s: got["stats"][s] for s in keys} == {s: ref["stats"][s] for s in keys}
    else:
        for s in keys:
            assert abs(got["stats"][s] - ref["stats"][s]) <= max(2, 0.1 * ref["stats"][s])
