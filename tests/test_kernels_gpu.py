"""GPU parity of every C-ABI kernel (include/primme_b200.h) against the CPU restatement
(oracle/kernels_ref.c) on the same seeded inputs.  fp64: the results differ only by summation
order; tolerances are stated per test (relative to the data scale)."""
import ctypes as C

import numpy as np
import pytest

import harness as H
from primme_b200 import api, matrices as M

pytestmark = pytest.mark.gpu


class Dev:
    """device-side mirror of column-major host arrays for one library"""

    def __init__(self, lib):
        self.lib = lib
        self.ctx = C.c_void_p()
        assert lib.pb200_ctx_create(C.byref(self.ctx), -1) == 0
        self.ptrs = []

    def up(self, a):  # a: numpy (cols, rows) C-order == column-major rows x cols
        a = np.ascontiguousarray(a, dtype=np.float64)
        p = C.c_void_p()
        assert self.lib.pb200_malloc(self.ctx, max(a.nbytes, 8), C.byref(p)) == 0
        ld = a.shape[-1]
        cols = a.shape[0] if a.ndim == 2 else 1
        assert self.lib.pb200_copy_h2d(self.ctx, a.ctypes.data, ld, p, ld, ld, cols, 8) == 0
        self.ptrs.append(p)
        return p

    def down(self, p, cols, rows):
        out = np.zeros((cols, rows))
        assert self.lib.pb200_copy_d2h(self.ctx, p, rows, out.ctypes.data, rows, rows, cols, 8) == 0
        return out

    def close(self):
        for p in self.ptrs:
            self.lib.pb200_free(self.ctx, p)
        self.lib.pb200_ctx_destroy(self.ctx)


@pytest.fixture(scope="module")
def libs():
    return H.lib_product(), H.lib_oracle_kernels()


def off(p, nbytes):
    return C.c_void_p(p.value + nbytes)


@pytest.mark.parametrize("n,q,mv,b,update,useY,xx", [
    # even n with pad 4 => even leading dimension, 16-byte aligned columns: TMA-staged kernel
    (3000 + 4j, 0, 28, 4, True, True, True),
    (3000 + 4j, 0, 28, 4, False, False, True),
    (5000 + 4j, 0, 40, 4, False, False, False),
    (2500 + 2j, 4, 60, 8, True, True, True),
    (6002 + 4j, 2, 30, 1, True, False, True),
    (10000 + 0j, 0, 12, 2, True, True, True),
    (70000 + 0j, 0, 36, 4, True, True, True),
    (1000, 0, 12, 4, False, False, True),
    (1000, 0, 12, 4, True, True, True),
    # specialised instances (no locked vectors, 3..5 tiles of V)
    (40000 + 4j, 0, 20, 4, True, True, True),
    (40000 + 4j, 0, 17, 3, True, False, True),
    (33000 + 2j, 0, 24, 8, True, True, True),
    (33000 + 2j, 0, 33, 8, False, False, False),
    (33000 + 2j, 0, 40, 5, False, False, True),
    (4097, 3, 20, 4, True, True, True),
    (4097, 3, 20, 3, True, False, True),
    (5000, 0, 40, 4, False, False, False),   # projection-like: X = W block, no xx
    (3001, 5, 0, 1, True, False, True),      # CGS against locked only
    (3001, 0, 35, 1, True, False, True),     # CGS b=1
    (2500, 4, 60, 8, True, True, True),      # C5-like widths
    (2500, 0, 95, 8, False, False, True),
    (2000, 10, 100, 2, True, True, True),    # more columns than one launch: chunked path
    (255, 0, 7, 5, True, True, True),
    (1, 0, 1, 1, False, False, True),
    (0, 0, 4, 2, False, False, True),        # empty local part
])
def test_ortho_sweep(libs, n, q, mv, b, update, useY, xx):
    pad = 3
    if isinstance(n, complex):  # n + pad*1j: explicit padding of the leading dimension
        n, pad = int(n.real), int(n.imag)
    rng = np.random.default_rng(1234 + n + q + mv + b)
    ld = n + pad
    Q = rng.standard_normal((max(q, 1), ld))
    V = rng.standard_normal((mv + b, ld))  # X = V(:, mv:mv+b)
    Cm = rng.standard_normal((b, q + mv + 2)) * 0.1  # ldc = q+mv+2
    Y = rng.standard_normal((b, b)) + 2 * np.eye(b)
    res = []
    for lib in libs:
        d = Dev(lib)
        dQ, dV = d.up(Q), d.up(V)
        dX = off(dV, 8 * ld * mv)
        rows = q + mv + (b if xx else 0)
        P = np.full((b, rows + 1), np.nan)
        rc = lib.pb200_dortho_sweep(d.ctx, n, dQ if q else None, q, ld, dV, mv, ld, dX, b, ld,
                                    Cm.ctypes.data if update else None, q + mv + 2,
                                    Y.ctypes.data if (update and useY) else None, b, 1 if xx else 0,
                                    P.ctypes.data, rows + 1)
        assert rc == 0
        Vout = d.down(dV, mv + b, ld)
        res.append((P[:, :rows].copy(), Vout[mv:, :n].copy(), Vout[:mv].copy()))
        d.close()
    (Pg, Xg, Vg), (Po, Xo, Vo) = res
    scale = max(1.0, np.abs(Po).max()) if Po.size else 1.0
    # tolerance: n-term fp64 sums in different orders -> ~ sqrt(n) * eps * scale, use a loose 1e-11
    assert np.allclose(Pg, Po, rtol=0, atol=1e-11 * scale * max(1, n) ** 0.5)
    assert np.allclose(Xg, Xo, rtol=1e-12, atol=1e-12)
    assert np.array_equal(Vg, Vo)  # the basis itself is never written


@pytest.mark.parametrize("n,m,nh,case", [
    (3000, 28, 4, "cand"),      # candidates: X, R, Rnorms
    (3000, 28, 4, "norms"),     # norms only
    (5001, 40, 24, "restart"),  # in-place V,W <- V*h, W*h ; X,R block ; G,H
    (2000, 64, 36, "restart"),  # C5-like restart (restart size 32 + block 4)
    (1999, 16, 12, "lock"),     # restart + columns to evecs + extra norms
    (3000, 40, 40, "restart"),  # restart keeping 36 vectors (basis not full): larger Gram blocks
    (3000, 64, 60, "restart"),  # Gram blocks beyond the single-launch kernel: general path
    (700, 150, 140, "lock"),    # more than 64 columns of h (reference test_001 keeps 140)
    (6144, 28, 4, "cand"),      # aligned, full tiles only: TMA-staged kernel
    (8192 + 34, 36, 4, "norms"),
    (5120 + 2, 56, 8, "cand"),
    (66000, 16, 3, "cand"),
    (6144 + 70, 40, 24, "restart"),
    (4096, 64, 36, "restart"),
    (257, 9, 3, "cand"),
    (0, 8, 2, "cand"),
    (6144, 28, 4, "candp"),     # candidates + first Gram panel of the block ortho, P = [V R]'R
    (3000 + 6, 37, 3, "candp"),
    (5120 + 2, 64, 8, "candp"),
    (40000, 20, 4, "candp"),
    (4096 + 8, 40, 12, "restart"),  # two 8-column tiles of h
    (4096 + 8, 33, 29, "lock"),     # four tiles
    (130000, 40, 24, "restart"),    # several tiles per CTA: ring wrap-around
    (150000, 28, 4, "candp"),
    (5000 + 2, 64, 44, "restart"),  # six tiles of h: the column-split restart kernel <6,2>
    (4097, 48, 40, "lock"),         # five tiles, odd n (single-row tail store), columns to evecs
    (70000 + 6, 64, 40, "restart"), # C5 restart shape, several tiles per CTA
    (9000, 24, 16, "restart"),      # two tiles <2,4>
    (3001, 96, 28, "restart"),      # widest basis the staged kernels take (4 tiles)
])
def test_vwxr(libs, n, m, nh, case):
    rng = np.random.default_rng(99 + n + m + nh)
    ld = n + 5 if n % 2 else n + 16
    V = rng.standard_normal((m + 8, ld))
    W = rng.standard_normal((m + 8, ld))
    h = rng.standard_normal((nh, m + 1))  # ldh = m+1
    theta = rng.standard_normal(nh)
    res = []
    for lib in libs:
        d = Dev(lib)
        dV, dW = d.up(V), d.up(W)
        E = np.zeros((8, ld))
        dE = d.up(E)
        o = api.VwxrOut()
        nR = 0
        Rn = np.zeros(16)
        rn = np.zeros(200)
        G = np.zeros((160, 161))
        Hm = np.zeros((160, 163))
        Pp = np.zeros((8, m + 8 + 3))
        if case in ("cand", "candp"):
            o.X[0] = api.VwxrCols(off(dV, 8 * ld * m).value, ld, 0, nh)
            o.R = api.VwxrCols(off(dW, 8 * ld * m).value, ld, 0, nh)
            o.Rnorms_host = Rn.ctypes.data
            nR = nh
            if case == "candp":
                if not lib.pb200_dvwxr_can_fuse_gram(d.ctx, n, dV, dW, m, ld, nh, C.byref(o)):
                    pytest.skip("shape not covered by the fused kernel")
                o.P_host, o.ldP = Pp.ctypes.data, m + 8 + 3
        elif case == "norms":
            o.rb, o.re, o.rnorms_host = 0, nh, rn.ctypes.data
        else:
            rs = nh - 4  # restart size, last 4 columns feed the next block
            nconv = 2
            o.X[0] = api.VwxrCols(dV.value, ld, 0, rs)
            o.Wo = api.VwxrCols(dW.value, ld, 0, rs)
            o.X[1] = api.VwxrCols(off(dV, 8 * ld * rs).value, ld, nconv, nconv + 4)
            o.R = api.VwxrCols(off(dW, 8 * ld * rs).value, ld, nconv, nconv + 4)
            o.Rnorms_host = Rn.ctypes.data
            nR = 4
            o.nG, o.G_host, o.ldG = rs, G.ctypes.data, 161
            o.nH, o.H_host, o.ldH = rs, Hm.ctypes.data, 163
            if case == "lock":
                o.X[2] = api.VwxrCols(dE.value, ld, rs - 3, rs)
                o.rb, o.re, o.rnorms_host = rs - 3, rs, rn.ctypes.data
        rc = lib.pb200_dvwxr(d.ctx, n, dV, dW, m, ld, h.ctypes.data, m + 1, nh, theta.ctypes.data, C.byref(o))
        assert rc == 0
        res.append((d.down(dV, m + 8, ld)[:, :n], d.down(dW, m + 8, ld)[:, :n], d.down(dE, 8, ld)[:, :n],
                    Rn[:nR].copy(), rn.copy(), G.copy(), Hm.copy(), Pp.copy()))
        d.close()
    g, o_ = res
    sc = np.sqrt(m) * 3
    for a, b_ in zip(g[:3], o_[:3]):
        assert np.allclose(a, b_, rtol=1e-12, atol=1e-12 * sc)
    assert np.allclose(g[3], o_[3], rtol=1e-11)
    assert np.allclose(g[4], o_[4], rtol=1e-11)
    assert np.allclose(g[5], o_[5], rtol=0, atol=1e-11 * max(1.0, np.abs(o_[5]).max()))
    assert np.allclose(g[6], o_[6], rtol=0, atol=1e-11 * max(1.0, np.abs(o_[6]).max()))
    assert np.allclose(g[7], o_[7], rtol=0, atol=1e-11 * max(1.0, np.abs(o_[7]).max()))


def _csr_case(name):
    if name == "lap3d":
        return M.laplacian_nd((17, 13, 11))
    if name == "lap1d":
        return M.laplacian_1d(5000)
    if name == "powerlaw":
        return M.power_law_symmetric(20000, mean_degree=12.0, seed=3)
    if name == "longrow":
        # one dense row/column (longer than the staging capacity) plus a diagonal
        n = 9000
        rows = np.concatenate([np.zeros(n, dtype=np.int64), np.arange(1, n), np.arange(n)])
        cols = np.concatenate([np.arange(n), np.zeros(n - 1, dtype=np.int64), np.arange(n)])
        vals = np.concatenate([np.linspace(1, 2, n), np.linspace(1, 2, n)[1:], np.full(n, 3.0)])
        key = rows * n + cols
        u, st = np.unique(key, return_index=True)
        v = np.add.reduceat(vals[np.argsort(key, kind="stable")], st)
        return M._assemble(n, u // n, u % n, v)
    if name == "emptyrows":
        ip, ix, da = M.laplacian_1d(100)
        ip = np.concatenate([ip, np.full(50, ip[-1])])  # 50 trailing empty rows
        return ip, ix, da
    raise KeyError(name)


@pytest.mark.parametrize("v3", [1, 0])
@pytest.mark.parametrize("name", ["lap3d", "lap1d", "powerlaw", "longrow", "emptyrows"])
@pytest.mark.parametrize("b", [1, 3, 4, 8, 11])
def test_spmm(libs, name, b, v3, monkeypatch):
    """v3 = 1: row-major gather copy + 32-byte gathers (default for b >= 2); 0: column-major gathers"""
    monkeypatch.setenv("PB200_SPMM_V3", str(v3))
    ip, ix, da = _csr_case(name)
    nrows = len(ip) - 1
    ncols = max(int(ix.max()) + 1, nrows) if len(ix) else nrows
    rng = np.random.default_rng(5)
    X = rng.standard_normal((b, ncols + 2))
    ref = M.csr_matvec(ip, ix, da, X[:, :ncols].T)
    for lib in libs:
        d = Dev(lib)
        A = C.c_void_p()
        rp = np.ascontiguousarray(ip, dtype=np.int64)
        ci = np.ascontiguousarray(ix, dtype=np.int32)
        va = np.ascontiguousarray(da, dtype=np.float64)
        assert lib.pb200_csr_create(d.ctx, nrows, ncols, len(ci), rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0, 0, C.byref(A)) == 0
        dX = d.up(X)
        dY = d.up(np.zeros((b, nrows + 1)))
        assert lib.pb200_dspmm(d.ctx, A, dX, ncols + 2, dY, nrows + 1, b) == 0
        Y = d.down(dY, b, nrows + 1)[:, :nrows].T
        scale = np.abs(ref).max() + 1
        assert np.allclose(Y, ref, rtol=0, atol=1e-12 * scale * 50), (name, b)
        # transpose product (normal-equations operator)
        assert lib.pb200_csr_build_transpose(d.ctx, A) == 0
        dZ = d.up(np.zeros((b, ncols)))
        dYin = d.up(np.ascontiguousarray(ref.T))
        assert lib.pb200_dspmm_t(d.ctx, A, dYin, nrows, dZ, ncols, b) == 0
        Z = d.down(dZ, b, ncols).T
        # A^T y via numpy
        rows = np.repeat(np.arange(nrows), np.diff(ip))
        zt = np.stack([np.bincount(ix, weights=da * ref[rows, j], minlength=ncols) for j in range(b)], axis=1)
        assert np.allclose(Z, zt, rtol=0, atol=1e-11 * (np.abs(zt).max() + 1) * 50)
        lib.pb200_csr_destroy(d.ctx, A)
        d.close()


def _banded_random(n, offsets, seed):
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for o in offsets:
        r = np.arange(max(0, -o), min(n, n - o))
        keep = rng.random(len(r)) < 0.8
        rows.append(r[keep]), cols.append(r[keep] + o), vals.append(rng.standard_normal(int(keep.sum())))
    return M._assemble(n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))


@pytest.mark.parametrize("name,expect", [("lap3d_big", 3), ("lap1d", 3), ("banded", 3), ("emptyrows", 3),
                                         ("banded_wide", 1), ("powerlaw", 1), ("longrow", 1)])
@pytest.mark.parametrize("b", [2, 4, 5, 8])
def test_spmm_windowed(libs, name, b, expect, monkeypatch):
    """windowed layout (v4): the right-hand-side windows of a row block are staged in shared memory by bulk
    copies; forced with PB200_SPMM_V3=3, matrices that do not qualify (no column locality, long rows) must fall
    back to the column-major gathers -- same product either way, compared with the CPU restatement / numpy"""
    monkeypatch.setenv("PB200_SPMM_V3", "3")
    if name == "lap3d_big":
        ip, ix, da = M.laplacian_nd((41, 30, 23))
    elif name == "banded_wide":  # more distinct column segments per row block than a window holds
        ip, ix, da = _banded_random(30011, [-7000, -3001, -130, -1, 0, 1, 2, 64, 2999, 7000, 11000], 17)
    elif name == "banded":
        ip, ix, da = _banded_random(30011, [-3001, -130, -1, 0, 1, 2, 64, 2999], 17)
    else:
        ip, ix, da = _csr_case(name)
    nrows = len(ip) - 1
    ncols = max(int(ix.max()) + 1, nrows) if len(ix) else nrows
    ld = (ncols + 2) // 2 * 2 + 2
    rng = np.random.default_rng(7)
    X = rng.standard_normal((b, ld))
    ref = M.csr_matvec(ip, ix, da, X[:, :ncols].T)
    lib = libs[0]
    d = Dev(lib)
    A = C.c_void_p()
    rp = np.ascontiguousarray(ip, dtype=np.int64)
    ci = np.ascontiguousarray(ix, dtype=np.int32)
    va = np.ascontiguousarray(da, dtype=np.float64)
    assert lib.pb200_csr_create(d.ctx, nrows, ncols, len(ci), rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0, 0, C.byref(A)) == 0
    dX = d.up(X)
    dY = d.up(np.zeros((b, nrows + 1)))
    for _ in range(2):
        assert lib.pb200_dspmm(d.ctx, A, dX, ld, dY, nrows + 1, b) == 0
    lib.pb200_csr_layout.restype = C.c_int
    # the windows of 8 columns do not fit next to a 512-row slice of the matrix for every pattern: either layout
    assert lib.pb200_csr_layout(A, b) in ((expect,) if b <= 4 else (expect, 1)), (name, lib.pb200_csr_layout(A, b))
    Y = d.down(dY, b, nrows + 1)[:, :nrows].T
    scale = np.abs(ref).max() + 1
    assert np.allclose(Y, ref, rtol=0, atol=1e-12 * scale * 50), (name, b)
    lib.pb200_csr_destroy(d.ctx, A)
    d.close()


@pytest.mark.parametrize("name", ["lap3d", "powerlaw", "longrow", "emptyrows"])
@pytest.mark.parametrize("b", [1, 2, 3, 8, 9])
def test_zspmm(name, b):
    """complex Hermitian-patterned CSR x complex block (zprimme's matvec): values get a random imaginary
    part; the product is compared with numpy complex arithmetic"""
    lib = H.lib_product()
    ip, ix, da = _csr_case(name)
    nrows = len(ip) - 1
    ncols = max(int(ix.max()) + 1, nrows) if len(ix) else nrows
    rng = np.random.default_rng(11)
    vz = da + 1j * rng.standard_normal(len(da))
    X = rng.standard_normal((b, ncols + 2)) + 1j * rng.standard_normal((b, ncols + 2))
    rows = np.repeat(np.arange(nrows), np.diff(ip))
    ref = np.zeros((nrows, b), dtype=complex)
    for j in range(b):
        prod = vz * X[j, ix]
        ref[:, j] = np.bincount(rows, weights=prod.real, minlength=nrows) + 1j * np.bincount(rows, weights=prod.imag, minlength=nrows)
    d = Dev(lib)
    lib.pb200_zspmm.restype = C.c_int
    lib.pb200_zspmm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]
    A = C.c_void_p()
    rp = np.ascontiguousarray(ip, dtype=np.int64)
    ci = np.ascontiguousarray(ix, dtype=np.int32)
    va = np.ascontiguousarray(vz, dtype=np.complex128)
    assert lib.pb200_csr_create(d.ctx, nrows, ncols, len(ci), rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0, 1, C.byref(A)) == 0
    dX = d.up(np.ascontiguousarray(X).view(np.float64))              # (b, 2*(ncols+2)) doubles
    dY = d.up(np.zeros((b, 2 * (nrows + 1))))
    assert lib.pb200_zspmm(d.ctx, A, dX, ncols + 2, dY, nrows + 1, b) == 0
    Y = d.down(dY, b, 2 * (nrows + 1)).view(np.complex128)[:, :nrows].T
    scale = np.abs(ref).max() + 1
    assert np.allclose(Y, ref, rtol=0, atol=1e-12 * scale * 50), (name, b)
    lib.pb200_csr_destroy(d.ctx, A)
    d.close()


def test_utilities(libs):
    rng = np.random.default_rng(8)
    n, ld, k = 3333, 3340, 9
    X = rng.standard_normal((k, ld))
    Y = rng.standard_normal((k, ld))
    alpha = rng.standard_normal(k)
    perm = rng.permutation(k).astype(np.int32)
    diag = rng.uniform(0.5, 2.0, ld)
    outs = []
    for lib in libs:
        d = Dev(lib)
        dX, dY, dD = d.up(X), d.up(Y), d.up(diag)
        r = {}
        dots = np.zeros(k)
        assert lib.pb200_dcolumn_dots(d.ctx, n, dX, ld, dY, ld, k, dots.ctypes.data) == 0
        r["dots"] = dots
        assert lib.pb200_daxpy_columns(d.ctx, n, alpha.ctypes.data, dX, ld, dY, ld, k) == 0
        r["axpy"] = d.down(dY, k, ld)[:, :n]
        assert lib.pb200_dscale_columns(d.ctx, n, alpha.ctypes.data, dY, ld, k) == 0
        r["scale"] = d.down(dY, k, ld)[:, :n]
        assert lib.pb200_dpermute_columns(d.ctx, n, dY, ld, perm.ctypes.data, k) == 0
        r["perm"] = d.down(dY, k, ld)[:, :n]
        res = np.zeros(k)
        assert lib.pb200_dresidual_inplace(d.ctx, n, alpha.ctypes.data, dX, ld, dY, ld, k, res.ctypes.data) == 0
        r["res"], r["resW"] = res, d.down(dY, k, ld)[:, :n]
        xin = np.array([3, 1, 7], dtype=np.int32)
        yin = np.array([0, 8, 2], dtype=np.int32)
        assert lib.pb200_dcopy_columns(d.ctx, n, dX, ld, xin.ctypes.data, dY, ld, yin.ctypes.data, 3) == 0
        r["copy"] = d.down(dY, k, ld)[:, :n]
        shifts = alpha[:4] * 0.1
        assert lib.pb200_djacobi(d.ctx, n, dD, shifts.ctypes.data, 1e-3, dX, ld, dY, ld, 4) == 0
        r["jac"] = d.down(dY, k, ld)[:4, :n]
        outs.append(r)
        d.close()
    g, o = outs
    assert np.allclose(g["dots"], o["dots"], rtol=1e-11, atol=1e-11)
    assert np.allclose(g["res"], o["res"], rtol=1e-11)
    for key in ("axpy", "scale", "perm", "resW", "copy", "jac"):
        assert np.allclose(g[key], o[key], rtol=1e-14, atol=1e-14), key
    assert np.array_equal(g["perm"], g["scale"][perm])


def test_panel_reduction_is_reproducible(libs):
    """fixed-order two-stage reduction: the same call twice gives bitwise equal panels"""
    lib = libs[0]
    rng = np.random.default_rng(0)
    n, mv, b = 200000, 30, 4
    V = rng.standard_normal((mv + b, n))
    d = Dev(lib)
    dV = d.up(V)
    P1 = np.zeros((b, mv + b))
    P2 = np.zeros((b, mv + b))
    for P in (P1, P2):
        # consecutive sweeps walk the rows in opposite directions (L2 reuse); a solve always starts forward
        lib.pb200_ctx_begin_solve.argtypes = [C.c_void_p]
        assert lib.pb200_ctx_begin_solve(d.ctx) == 0
        assert lib.pb200_dortho_sweep(d.ctx, n, None, 0, n, dV, mv, n, off(dV, 8 * n * mv), b, n, None, 0, None, 0, 1,
                                      P.ctypes.data, mv + b) == 0
    d.close()
    assert np.array_equal(P1, P2)


@pytest.mark.parametrize("n,ncols,seed", [(1, 1, (0, 0, 0, 1)), (63, 1, (1, 2, 3, 5)), (64, 2, (4095, 4095, 4095, 4095)),
                                          (65, 3, (7, 0, 11, 13)), (100003, 4, (0, 1, 2, 3)), (4097, 8, (1234, 567, 89, 1011)),
                                          (300001, 1, (17, 4000, 2, 4093))])
def test_dlarnv_on_device_is_lapack_dlarnv(libs, n, ncols, seed):
    """pb200_dlarnv: the initial / replacement random vectors are drawn on the device; the values AND the seed
    left behind must be LAPACK's dlarnv(idist=2) bit for bit (the oracle twin calls dlarnv_ itself), otherwise
    every solve would start from a different vector than the reference's"""
    out = []
    ld = n + 3
    for lib in libs:
        d = Dev(lib)
        dX = d.up(np.zeros((ncols, ld)))
        iseed = (C.c_longlong * 4)(*seed)
        lib.pb200_dlarnv.restype = C.c_int
        lib.pb200_dlarnv.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int64, C.c_int, C.c_void_p, C.c_int64]
        assert lib.pb200_dlarnv(d.ctx, iseed, n, ncols, dX, ld) == 0
        out.append((d.down(dX, ncols, ld)[:, :n].copy(), tuple(iseed)))
        d.close()
    (xg, sg), (xo, so) = out
    assert sg == so
    assert np.array_equal(xg, xo)
    assert np.abs(xg).max() < 1.0 and xg.std() > 0.3 or n < 10
