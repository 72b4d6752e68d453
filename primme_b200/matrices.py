"""Deterministic synthetic CSR matrices of the benchmark configurations.

All generators return (indptr int64, indices int32, data float64) with sorted column indices per
row, 0-based, symmetric where stated.  The same bytes feed the reference CPU path and the GPU.
"""
import numpy as np


def laplacian_1d(n):
    """tridiag(-1, 2, -1) (reference examples/ex_eigs_dseq.c:160-178)"""
    return laplacian_nd((n,))


def laplacian_nd(shape):
    """d-dimensional 2d+1 point Laplacian, natural ordering (first index fastest), Dirichlet.
    3-D with N=100: the matrix of config C2 (nnz = 7N^3 - 6N^2)."""
    shape = tuple(int(s) for s in shape)
    d = len(shape)
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.int64)
    coords = []
    rem = idx.copy()
    for s in shape:
        coords.append(rem % s)
        rem //= s
    strides = np.cumprod((1,) + shape[:-1]).astype(np.int64)
    rows = [idx]
    cols = [idx]
    vals = [np.full(n, 2.0 * d)]
    for ax in range(d):
        m = coords[ax] > 0
        rows.append(idx[m]); cols.append(idx[m] - strides[ax]); vals.append(np.full(m.sum(), -1.0))
        m = coords[ax] < shape[ax] - 1
        rows.append(idx[m]); cols.append(idx[m] + strides[ax]); vals.append(np.full(m.sum(), -1.0))
    return _assemble(n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))


def _assemble(n, rows, cols, vals):
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, cols.astype(np.int32), vals.astype(np.float64)


def power_law_symmetric(n, mean_degree=15.0, exponent=2.1, seed=7):
    """Chung-Lu style symmetric graph with expected degrees ~ k^-exponent, values U(0,1),
    mirrored (config C5 family).  Duplicate edges are merged (values summed)."""
    rng = np.random.default_rng(seed)
    # expected degree sequence
    u = rng.random(n)
    kmin = 1.0
    w = kmin * (1.0 - u) ** (-1.0 / (exponent - 1.0))
    w = np.minimum(w, np.sqrt(n * mean_degree))
    w *= mean_degree / w.mean()
    m = int(n * mean_degree / 2)
    p = w / w.sum()
    cdf = np.cumsum(p)
    a = np.searchsorted(cdf, rng.random(m)).astype(np.int64)
    b = np.searchsorted(cdf, rng.random(m)).astype(np.int64)
    a = np.minimum(a, n - 1); b = np.minimum(b, n - 1)
    keep = a != b
    a, b = a[keep], b[keep]
    v = rng.random(a.size)
    rows = np.concatenate([a, b]); cols = np.concatenate([b, a]); vals = np.concatenate([v, v])
    # merge duplicates
    key = rows * n + cols
    order = np.argsort(key, kind="stable")
    key, vals = key[order], vals[order]
    uniq, start = np.unique(key, return_index=True)
    vals = np.add.reduceat(vals, start)
    rows, cols = uniq // n, uniq % n
    return _assemble(n, rows, cols, vals)


def random_rectangular(m, n, per_row=20, seed=2024):
    """m x n rectangular CSR, per_row nonzeros per row at LCG-random columns, values U(-1,1)
    (config C4 family)."""
    rng = np.random.default_rng(seed)
    cols = rng.integers(0, n, size=(m, per_row), dtype=np.int64)
    cols.sort(axis=1)
    vals = rng.uniform(-1.0, 1.0, size=(m, per_row))
    rows = np.repeat(np.arange(m, dtype=np.int64), per_row)
    key = rows * n + cols.ravel()
    uniq, start = np.unique(key, return_index=True)
    v = np.add.reduceat(vals.ravel()[np.argsort(key, kind="stable")], start)
    return _assemble_rect(m, uniq // n, uniq % n, v)


def _assemble_rect(m, rows, cols, vals):
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(m + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr), cols.astype(np.int32), vals.astype(np.float64)


def csr_matvec(indptr, indices, data, x):
    """plain numpy y = A x (1-D or 2-D column block), for tests"""
    x = np.asarray(x)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    if x.ndim == 1:
        return np.bincount(rows, weights=data * x[indices], minlength=len(indptr) - 1)
    return np.stack([np.bincount(rows, weights=data * x[indices, j], minlength=len(indptr) - 1)
                     for j in range(x.shape[1])], axis=1)


def _plaw_params(n, mean_degree, exponent):
    """weight profile w(x) = min(c x^-alpha, wmax), x = (rank + 0.5) / n, scaled to the mean degree"""
    alpha = 1.0 / (exponent - 1.0)
    wmax = np.sqrt(n * mean_degree)

    def mean_w(c):
        x0 = min(1.0, (c / wmax) ** (1.0 / alpha))
        return wmax * x0 + c * (1.0 - x0 ** (1.0 - alpha)) / (1.0 - alpha), x0

    lo, hi = 1e-6, mean_degree
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if mean_w(mid)[0] < mean_degree:
            lo = mid
        else:
            hi = mid
    c = 0.5 * (lo + hi)
    return alpha, wmax, c, mean_w(c)[1]


def _stable_argsort(key):
    """stable argsort of int64 keys; on a GPU box the 10^8-key sort of the C5 matrix runs on the device
    (same permutation: stable sorts are unique)"""
    if key.size > 1 << 22:
        try:
            import torch
            if torch.cuda.is_available():
                k = torch.from_numpy(key).cuda()
                order = torch.sort(k, stable=True).indices.cpu().numpy()
                del k
                torch.cuda.empty_cache()
                return order
        except Exception:
            pass
    return np.argsort(key, kind="stable")


def power_law_rows(n, lo=0, hi=None, mean_degree=15.0, exponent=2.1, seed=7, chunk=1 << 22):
    """Rows [lo, hi) of the symmetric Chung-Lu power-law graph of config C5 (expected degrees
    ~ k^-exponent truncated at sqrt(n * mean_degree), mean `mean_degree`, values U(0,1) mirrored,
    duplicate edges merged, largest value kept).  The edge stream is generated in fixed chunks from
    counter-seeded generators, so every row shard of the SAME matrix can be produced independently
    (one process per GPU builds only its rows); node ids are a fixed random permutation of the weight
    ranks, so contiguous row shards are balanced.  Returns (indptr[hi-lo+1], indices (global), data)."""
    hi = n if hi is None else hi
    alpha, wmax, c, x0 = _plaw_params(n, mean_degree, exponent)
    perm = np.random.default_rng([seed, 0xC5]).permutation(n).astype(np.int64)
    m = int(n * mean_degree / 2)
    head = wmax * x0
    one_a = 1.0 - alpha
    x0p = x0 ** one_a
    total = head + c * (1.0 - x0p) / one_a

    def node(u):
        u = u * total
        x = np.where(u < head, u / wmax, np.power(x0p + one_a * np.maximum(u - head, 0.0) / c, 1.0 / one_a))
        k = np.minimum((x * n).astype(np.int64), n - 1)
        return perm[k]

    R, Cc, Vv = [], [], []
    for j, start in enumerate(range(0, m, chunk)):
        cnt = min(chunk, m - start)
        rng = np.random.default_rng([seed, 1, j])
        u = rng.random((3, cnt))
        a, b, v = node(u[0]), node(u[1]), u[2]
        keep = a != b
        for r_, c_ in ((a, b), (b, a)):
            sel = keep & (r_ >= lo) & (r_ < hi)
            R.append((r_[sel] - lo).astype(np.int32)); Cc.append(c_[sel].astype(np.int32)); Vv.append(v[sel])
    rows = np.concatenate(R); cols = np.concatenate(Cc); vals = np.concatenate(Vv)
    del R, Cc, Vv
    key = rows.astype(np.int64) * n + cols
    del rows, cols
    order = _stable_argsort(key)
    key = key[order]; vals = vals[order]
    del order
    first = np.ones(key.size, dtype=bool)
    first[1:] = key[1:] != key[:-1]
    start_idx = np.flatnonzero(first)
    # duplicates keep the LARGEST value: independent of the order of the stream, so A is exactly symmetric
    vals = np.maximum.reduceat(vals, start_idx) if key.size else vals
    key = key[first]
    rloc = (key // n).astype(np.int64)
    indices = (key % n).astype(np.int32)
    indptr = np.zeros(hi - lo + 1, dtype=np.int64)
    indptr[1:] = np.bincount(rloc, minlength=hi - lo)
    indptr = np.cumsum(indptr)
    return indptr, indices, vals.astype(np.float64)


def hermitian_c3(n, seed=12345, offdiag_pairs=4, amp=0.1, hole=1):
    """Complex Hermitian matrix of config C3: tridiagonal plus `offdiag_pairs` random
    off-diagonal pairs per row, off-diagonal values (re, im) ~ U(-amp, amp)^2 mirrored conjugate, real
    diagonal d_i = (i + 0.5) / n in (0, 1) so that sigma = 0.5 is interior.  Returns (indptr, indices,
    complex128 data), sorted columns, duplicates merged (largest magnitude kept)."""
    rng = np.random.default_rng(seed)
    i = np.arange(n - 1, dtype=np.int64)
    a = [i]
    b = [i + 1]
    for _ in range(offdiag_pairs):
        r = np.arange(n, dtype=np.int64)
        c = rng.integers(0, n, size=n, dtype=np.int64)
        keep = c > r + 1
        a.append(r[keep]); b.append(c[keep])
    a = np.concatenate(a); b = np.concatenate(b)
    v = rng.uniform(-amp, amp, size=a.size) + 1j * rng.uniform(-amp, amp, size=a.size)
    key = a * n + b
    order = np.argsort(key, kind="stable")
    key, v = key[order], v[order]
    first = np.ones(key.size, dtype=bool)
    first[1:] = key[1:] != key[:-1]
    key, v = key[first], v[first]          # first occurrence of a duplicated pair
    a, b = key // n, key % n
    d = (np.arange(n) + 0.5) / n
    if hole > 1:
        # thin the spectrum around sigma = 0.5: density ~ |2d - 1|^(hole - 1), still d in (0, 1)
        d = 0.5 + np.sign(d - 0.5) * np.abs(2 * d - 1) ** (1.0 / hole) / 2
    rows = np.concatenate([a, b, np.arange(n)])
    cols = np.concatenate([b, a, np.arange(n)])
    vals = np.concatenate([v, np.conj(v), d.astype(np.complex128)])
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(n + 1, dtype=np.int64)
    indptr[1:] = np.bincount(rows, minlength=n)
    return np.cumsum(indptr), cols.astype(np.int32), vals.astype(np.complex128)


# Config C3 as benchmarked and tested.  SURVEY 8(d) sketches amp = 0.1 on the plain diagonal (i + 0.5)/n;
# with that matrix the UNMODIFIED reference does not converge either (zprimme returns -3 after 2*10^5
# matvecs already at n = 600: the Jacobi preconditioner (D - 0.5)^-1 is singular where the spectrum is
# densest and the couplings exceed the eigenvalue gaps ~1/n by orders of magnitude).  The spectrum is
# therefore thinned around sigma (hole = 3: density ~ (2d - 1)^2) and the couplings scaled to 0.01, which
# keeps the structure (tridiagonal + 4 random pairs per row, complex couplings, interior target, Jacobi)
# and lets both solvers converge in a few hundred matvecs.
C3_MATRIX = dict(amp=0.01, hole=3)
