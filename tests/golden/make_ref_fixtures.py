"""Convert the reference's own golden data for the LUNDA test configurations into one small
fixture: tests/LUNDA.mtx (147 x 147 symmetric, MatrixMarket) and the stored eigenvector files
tests/tests/sol_00{1..5}_double (format: reference tests/COMMON/ioandtest.c:159-263 -- three header
doubles [sizeof scalar, n, #columns] then the columns).  The solver parameters of each case are
those of tests/tests/test_00{1..5}.  Run in the build container (needs /root/reference):
    python tests/golden/make_ref_fixtures.py
Output: tests/golden/lunda_fixture.npz (committed; the GPU box has no /root/reference)."""
import os

import numpy as np

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_mtx_symmetric(path):
    with open(path) as f:
        header = f.readline()
        assert "symmetric" in header
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        n, m, nz = (int(x) for x in line.split())
        rows, cols, vals = [], [], []
        for _ in range(nz):
            i, j, v = f.readline().split()
            i, j, v = int(i) - 1, int(j) - 1, float(v)
            rows.append(i); cols.append(j); vals.append(v)
            if i != j:
                rows.append(j); cols.append(i); vals.append(v)
    rows, cols, vals = np.array(rows), np.array(cols), np.array(vals)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr), cols.astype(np.int32), vals


def read_sol(path, n):
    raw = np.fromfile(path, dtype=np.float64)
    assert int(raw[0]) == 8 and int(raw[1]) == n
    cols = int(raw[2])
    return raw[3:3 + n * cols].reshape(cols, n).copy()  # one eigenvector per row


ip, ix, da = read_mtx_symmetric(os.path.join(REF, "LUNDA.mtx"))
n = len(ip) - 1
out = dict(indptr=ip, indices=ix, data=da, fnorm=np.sqrt((da ** 2).sum()))
for t in ("001", "002", "003", "004", "005"):
    out["sol_" + t] = read_sol(os.path.join(REF, "tests", f"sol_{t}_double"), n)
np.savez_compressed(os.path.join(HERE, "lunda_fixture.npz"), **out)
print("n", n, "nnz", len(ix), {k: v.shape for k, v in out.items() if k.startswith("sol")})
