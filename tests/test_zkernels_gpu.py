"""GPU parity of the complex kernels behind zprimme (include/primme_b200.h: pb200_zortho_sweep, pb200_zvwxr, the
multivector utilities) against the plain-C99 restatement oracle/kernels_ref_z.c on the same seeded inputs.  Complex
fp64: results differ only by summation order; tolerances as in tests/test_kernels_gpu.py."""
import ctypes as C

import numpy as np
import pytest

import harness as H
from primme_b200 import api

pytestmark = pytest.mark.gpu


class ZDev:
    """device-side mirror of column-major complex host arrays (numpy (cols, rows) complex128)"""

    def __init__(self, lib):
        self.lib = lib
        self.ctx = C.c_void_p()
        assert lib.pb200_ctx_create(C.byref(self.ctx), -1) == 0
        self.ptrs = []

    def up(self, a):
        a = np.ascontiguousarray(a, dtype=np.complex128)
        p = C.c_void_p()
        assert self.lib.pb200_malloc(self.ctx, max(a.nbytes, 16), C.byref(p)) == 0
        ld = a.shape[-1]
        cols = a.shape[0] if a.ndim == 2 else 1
        assert self.lib.pb200_copy_h2d(self.ctx, a.ctypes.data, ld, p, ld, ld, cols, 16) == 0
        self.ptrs.append(p)
        return p

    def down(self, p, cols, rows):
        out = np.zeros((cols, rows), dtype=np.complex128)
        assert self.lib.pb200_copy_d2h(self.ctx, p, rows, out.ctypes.data, rows, rows, cols, 16) == 0
        return out

    def close(self):
        for p in self.ptrs:
            self.lib.pb200_free(self.ctx, p)
        self.lib.pb200_ctx_destroy(self.ctx)


def declare(lib):
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    lib.pb200_zortho_sweep.restype = i32
    lib.pb200_zortho_sweep.argtypes = [vp, i64, vp, i32, i64, vp, i32, i64, vp, i32, i64, vp, i32, vp, i32, i32, vp, i32]
    lib.pb200_zvwxr.restype = i32
    lib.pb200_zvwxr.argtypes = [vp, i64, vp, vp, i32, i64, vp, i32, i32, vp, vp]
    lib.pb200_zpermute_columns.argtypes = [vp, i64, vp, i64, vp, i32]
    lib.pb200_zcopy_columns.argtypes = [vp, i64, vp, i64, vp, vp, i64, vp, i32]
    lib.pb200_zaxpy_columns.argtypes = [vp, i64, vp, vp, i64, vp, i64, i32]
    lib.pb200_zscale_columns.argtypes = [vp, i64, vp, vp, i64, i32]
    lib.pb200_zcolumn_dots.argtypes = [vp, i64, vp, i64, vp, i64, i32, vp]
    lib.pb200_zresidual_inplace.argtypes = [vp, i64, vp, vp, i64, vp, i64, i32, vp]
    lib.pb200_zjacobi.argtypes = [vp, i64, vp, vp, dbl, vp, i64, vp, i64, i32]


@pytest.fixture(scope="module")
def libs():
    a, b = H.lib_product(), H.lib_oracle_kernels()
    declare(a), declare(b)
    return a, b


def off(p, nbytes):
    return C.c_void_p(p.value + nbytes)


def crandn(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.mark.parametrize("n,q,mv,b,update,useY,xx", [
    (3001, 0, 35, 1, True, False, True),      # C3: CGS pass, update with the previous overlaps + new overlaps
    (3001, 8, 27, 1, False, False, True),     # C3 with locked vectors, Gram only
    (5000, 0, 28, 4, True, True, True),
    (5000, 0, 28, 4, False, False, False),    # projection-like (X = W block)
    (2500, 4, 60, 8, True, True, True),       # more columns than one Gram launch at b = 8: chunked
    (4097, 3, 20, 3, True, False, True),
    (2000, 10, 130, 2, True, True, True),     # more columns than one update launch
    (70000, 0, 36, 2, True, True, True),
    (255, 0, 7, 5, True, True, True),
    (1, 0, 1, 1, False, False, True),
    (0, 0, 4, 2, False, False, True),
])
def test_zortho_sweep(libs, n, q, mv, b, update, useY, xx):
    rng = np.random.default_rng(4321 + n + q + mv + b)
    ld = n + 3
    Q = crandn(rng, (max(q, 1), ld))
    V = crandn(rng, (mv + b, ld))
    Cm = crandn(rng, (b, q + mv + 2)) * 0.1
    Y = crandn(rng, (b, b)) + 2 * np.eye(b)
    res = []
    for lib in libs:
        d = ZDev(lib)
        dQ, dV = d.up(Q), d.up(V)
        dX = off(dV, 16 * ld * mv)
        rows = q + mv + (b if xx else 0)
        P = np.full((b, rows + 1), np.nan, dtype=np.complex128)
        rc = lib.pb200_zortho_sweep(d.ctx, n, dQ if q else None, q, ld, dV, mv, ld, dX, b, ld,
                                    Cm.ctypes.data if update else None, q + mv + 2,
                                    Y.ctypes.data if (update and useY) else None, b, 1 if xx else 0,
                                    P.ctypes.data, rows + 1)
        assert rc == 0
        Vout = d.down(dV, mv + b, ld)
        res.append((P[:, :rows].copy(), Vout[mv:, :n].copy(), Vout[:mv].copy()))
        d.close()
    (Pg, Xg, Vg), (Po, Xo, Vo) = res
    scale = max(1.0, np.abs(Po).max()) if Po.size else 1.0
    assert np.allclose(Pg, Po, rtol=0, atol=1e-11 * scale * max(1, n) ** 0.5)
    assert np.allclose(Xg, Xo, rtol=1e-12, atol=1e-12 * max(1.0, q + mv) ** 0.5)
    assert np.array_equal(Vg, Vo)


@pytest.mark.parametrize("n,m,nh,case", [
    (3000, 28, 1, "cand"),
    (3000, 35, 4, "cand"),
    (3000, 28, 4, "norms"),
    (5001, 35, 23, "restart"),   # C3: restart size 21 + 1 retained + block
    (1999, 16, 12, "lock"),
    (700, 150, 140, "lock"),     # reference test_101 keeps up to 140 vectors
    (66000, 16, 3, "cand"),
    (0, 8, 2, "cand"),
])
def test_zvwxr(libs, n, m, nh, case):
    rng = np.random.default_rng(77 + n + m + nh)
    ld = n + 5
    V = crandn(rng, (m + 8, ld))
    W = crandn(rng, (m + 8, ld))
    h = crandn(rng, (nh, m + 1))
    theta = rng.standard_normal(nh)
    res = []
    for lib in libs:
        d = ZDev(lib)
        dV, dW = d.up(V), d.up(W)
        dE = d.up(np.zeros((8, ld), dtype=complex))
        o = api.VwxrOut()
        nR = 0
        Rn, rn = np.zeros(16), np.zeros(200)
        G = np.zeros((160, 161), dtype=complex)
        Hm = np.zeros((160, 163), dtype=complex)
        if case == "cand":
            o.X[0] = api.VwxrCols(off(dV, 16 * ld * m).value, ld, 0, nh)
            o.R = api.VwxrCols(off(dW, 16 * ld * m).value, ld, 0, nh)
            o.Rnorms_host = Rn.ctypes.data
            nR = nh
        elif case == "norms":
            o.rb, o.re, o.rnorms_host = 0, nh, rn.ctypes.data
        else:
            rs = nh - 4
            nconv = 2
            o.X[0] = api.VwxrCols(dV.value, ld, 0, rs)       # in place: V <- V h
            o.Wo = api.VwxrCols(dW.value, ld, 0, rs)
            o.X[1] = api.VwxrCols(off(dV, 16 * ld * rs).value, ld, nconv, nconv + 4)
            o.R = api.VwxrCols(off(dW, 16 * ld * rs).value, ld, nconv, nconv + 4)
            o.Rnorms_host = Rn.ctypes.data
            nR = 4
            o.nG, o.G_host, o.ldG = rs, G.ctypes.data, 161
            o.nH, o.H_host, o.ldH = rs, Hm.ctypes.data, 163
            if case == "lock":
                o.X[2] = api.VwxrCols(dE.value, ld, rs - 3, rs)
                o.rb, o.re, o.rnorms_host = rs - 3, rs, rn.ctypes.data
        rc = lib.pb200_zvwxr(d.ctx, n, dV, dW, m, ld, h.ctypes.data, m + 1, nh, theta.ctypes.data, C.byref(o))
        assert rc == 0
        res.append((d.down(dV, m + 8, ld)[:, :n], d.down(dW, m + 8, ld)[:, :n], d.down(dE, 8, ld)[:, :n],
                    Rn[:nR].copy(), rn.copy(), G.copy(), Hm.copy()))
        d.close()
    g, o_ = res
    sc = np.sqrt(m) * 4
    for a, b_ in zip(g[:3], o_[:3]):
        assert np.allclose(a, b_, rtol=1e-12, atol=1e-12 * sc)
    assert np.allclose(g[3], o_[3], rtol=1e-11)
    assert np.allclose(g[4], o_[4], rtol=1e-11)
    assert np.allclose(g[5], o_[5], rtol=0, atol=1e-11 * max(1.0, np.abs(o_[5]).max()))
    assert np.allclose(g[6], o_[6], rtol=0, atol=1e-11 * max(1.0, np.abs(o_[6]).max()))


@pytest.mark.parametrize("n", [0, 1, 5000, 150001])
def test_zutilities(libs, n):
    rng = np.random.default_rng(5 + n)
    ld, nc = n + 2, 11
    X, Y = crandn(rng, (nc, ld)), crandn(rng, (nc, ld))
    alpha = crandn(rng, nc)
    theta = rng.standard_normal(nc)
    diag = rng.uniform(0.0, 1.0, max(n, 1))
    diag[: min(n, 3)] = 0.3  # hits the safeguard with shift 0.3
    shifts = np.full(nc, 0.3)
    perm = np.array([3, 0, 1, 2, 4, 10, 6, 7, 8, 9, 5], dtype=np.int32)
    xin = np.array([2, 5, 7], dtype=np.int32)
    yin = np.array([0, 1, 9], dtype=np.int32)
    res = []
    for lib in libs:
        d = ZDev(lib)
        dX, dY = d.up(X), d.up(Y)
        dots = np.zeros(nc, dtype=complex)
        assert lib.pb200_zcolumn_dots(d.ctx, n, dX, ld, dY, ld, nc, dots.ctypes.data) == 0
        assert lib.pb200_zaxpy_columns(d.ctx, n, alpha.ctypes.data, dX, ld, dY, ld, nc) == 0
        y1 = d.down(dY, nc, ld)[:, :n].copy()
        assert lib.pb200_zscale_columns(d.ctx, n, alpha.ctypes.data, dY, ld, nc) == 0
        y2 = d.down(dY, nc, ld)[:, :n].copy()
        rn = np.zeros(nc)
        assert lib.pb200_zresidual_inplace(d.ctx, n, theta.ctypes.data, dX, ld, dY, ld, nc, rn.ctypes.data) == 0
        y3 = d.down(dY, nc, ld)[:, :n].copy()
        assert lib.pb200_zpermute_columns(d.ctx, n, dY, ld, perm.ctypes.data, nc) == 0
        y4 = d.down(dY, nc, ld)[:, :n].copy()
        assert lib.pb200_zcopy_columns(d.ctx, n, dX, ld, xin.ctypes.data, dY, ld, yin.ctypes.data, 3) == 0
        y5 = d.down(dY, nc, ld)[:, :n].copy()
        dd = C.c_void_p()
        assert lib.pb200_malloc(d.ctx, 8 * max(n, 1), C.byref(dd)) == 0
        assert lib.pb200_copy_h2d(d.ctx, diag.ctypes.data, max(n, 1), dd, max(n, 1), max(n, 1), 1, 8) == 0
        assert lib.pb200_zjacobi(d.ctx, n, dd, shifts.ctypes.data, 1e-3, dX, ld, dY, ld, nc) == 0
        y6 = d.down(dY, nc, ld)[:, :n].copy()
        lib.pb200_free(d.ctx, dd)
        res.append((dots, y1, y2, rn, y3, y4, y5, y6))
        d.close()
    g, o_ = res
    assert np.allclose(g[0], o_[0], rtol=0, atol=1e-11 * max(1.0, n) ** 0.5 * 4)
    assert np.allclose(g[3], o_[3], rtol=1e-11, atol=1e-300)
    for i in (1, 2, 4, 5, 6, 7):
        assert np.allclose(g[i], o_[i], rtol=1e-13, atol=1e-13), i
    if n > 0:   # independent of the oracle: numpy
        assert np.allclose(g[0], np.einsum("cr,cr->c", X[:, :n].conj(), Y[:, :n]), rtol=0, atol=1e-9 * n ** 0.5)
